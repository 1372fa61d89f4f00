"""bench.py's reference arm (the CPU leg the driver runs next to the GPU arm) prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

from conftest import REPO


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=280, cwd=REPO)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['higher_is_better'] is True and d['value'] > 0
    assert d['unit'] == 'HR-MP/s' and 'RRDB' in d['metric'] and d['data'] == 'synthetic'
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']
