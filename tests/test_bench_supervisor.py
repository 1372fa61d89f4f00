"""bench.py's supervisor (CPU): every rank runs the benchmark in a child process and, when rank 0's child ends without the JSON line, all ranks
start a fresh child - at most three attempts - agreeing through their own small TCP store.  Without a GPU every child fails at once, which
exercises exactly that path: three attempts per rank, nothing on stdout, a non-zero exit code, no rank left waiting on the store."""
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_base():
    """a MASTER_PORT whose supervisor ports (base + 1717 .. + 1720) are free right now"""
    import socket
    for _ in range(50):
        with socket.socket() as a:
            a.bind(('127.0.0.1', 0))
            port = a.getsockname()[1]
        if port < 3000 or port > 65000:
            continue
        ok = True
        for k in range(4):
            with socket.socket() as b:
                try:
                    b.bind(('127.0.0.1', port + k))
                except OSError:
                    ok = False
                    break
        if ok:
            return port - 1717
    return 29871


def _run(env_extra, port):
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), **env_extra)
    env.pop('ESR_BENCH_WORKER', None)
    return subprocess.Popen([sys.executable, os.path.join(REPO, 'bench.py'), '--steps', '1', '--warmup', '1', '--no-extras'],
                            env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without a GPU: the children must fail')
def test_supervisor_restarts_all_ranks_and_gives_up_after_three_attempts():
    base = _free_base()
    procs = [_run({'RANK': str(r), 'LOCAL_RANK': str(r), 'WORLD_SIZE': '2'}, base) for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode != 0
        assert out.strip() == '', out                                         # stdout carries the JSON line or nothing
        assert err.count('ended without a result') == 3, err[-2000:]
        assert 'DistNetworkError' not in err and 'supervisor store' not in err, err[-2000:]


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without a GPU: the children must fail')
def test_supervisor_single_process():
    p = _run({}, _free_base())
    out, err = p.communicate(timeout=300)
    assert p.returncode != 0 and out.strip() == '' and err.count('ended without a result') == 3
