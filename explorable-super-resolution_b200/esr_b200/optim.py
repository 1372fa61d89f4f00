"""FlatAdam: torch.optim.Adam (the optimizer models/SRRaGAN_model.py:182,188 builds for G and D) whose CUDA parameters, gradients and
moments live in flat fp32 buffers.

  * `step()` is ONE launch of esr_adam_multi over the flat buffers (a generator has 702 parameter tensors: torch's foreach Adam
    issues a dozen multi-tensor launches and the single-tensor one ~5000 small ones);
  * the gradient buffer is registered: the engines (esr_b200.engine / esr_b200.disc) write every weight gradient straight into its
    view (`grad_view(p)`), so nothing is allocated, concatenated or copied per step, and the data-parallel all-reduce runs on the flat
    buffer itself (esr_b200.parallel.average_gradients);
  * state_dict() / load_state_dict() keep torch.optim.Adam's layout (per-parameter `step`, `exp_avg`, `exp_avg_sq`), so the
    reference's checkpoints (`optimizer_state_dict`, models/base_model.py:114-140) load and save unchanged.
Parameters on the CPU (the orchestration tests drive the model with stand-in networks there) take torch.optim.Adam's own step."""
import ctypes as C
import weakref

import torch

from . import lib as L

_GRAD_VIEWS = {}      # id(param) -> (weak reference to the parameter, its view into a flat gradient buffer)


def grad_view(p):
    """the registered flat-buffer view a gradient of `p` should be written into, or None"""
    ent = _GRAD_VIEWS.get(id(p))
    if ent is None:
        return None
    if ent[0]() is not p:
        del _GRAD_VIEWS[id(p)]
        return None
    return ent[1]


def _register_view(p, gv):
    key = id(p)
    _GRAD_VIEWS[key] = (weakref.ref(p, lambda _r, k=key: _GRAD_VIEWS.pop(k, None)), gv)


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        super(FlatAdam, self).__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._flat = {}       # group index -> dict(p, g, m, v, views, scratch, uploaded)

    # ---------------------------------------------------------------- flat buffers
    def _cuda_params(self, group):
        return [p for p in group['params'] if p.requires_grad and p.is_cuda and p.dtype == torch.float32]

    def _build(self, gi, group):
        params = self._cuda_params(group)
        dev = params[0].device
        # every tensor starts on a 16-byte boundary inside the flat buffers (float4 accesses)
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        flat = {k: torch.zeros(off, dtype=torch.float32, device=dev) for k in ('p', 'g', 'm', 'v')}
        views = []
        old_steps = {float(self.state[p]['step']) for p in params if 'step' in self.state[p]}
        # ONE step counter shared by every parameter of the group (each state entry references the same tensor, so state_dict()
        # keeps torch.optim.Adam's layout) - unless a loaded checkpoint holds unequal counts
        shared = torch.tensor(old_steps.pop() if len(old_steps) == 1 else 0.0, dtype=torch.float32) if len(old_steps) <= 1 else None
        with torch.no_grad():
            for p, o in zip(params, offs):
                n = p.numel()
                pv = flat['p'][o:o + n].view_as(p)
                pv.copy_(p.data)
                p.data = pv                                   # the parameter now LIVES in the flat buffer (same values, same shape)
                st = self.state[p]
                mv, vv = flat['m'][o:o + n].view_as(p), flat['v'][o:o + n].view_as(p)
                if 'exp_avg' in st:                           # resumed from a checkpoint
                    mv.copy_(st['exp_avg'])
                    vv.copy_(st['exp_avg_sq'])
                st['exp_avg'], st['exp_avg_sq'] = mv, vv
                if shared is not None:
                    st['step'] = shared
                else:
                    st.setdefault('step', torch.tensor(0.0, dtype=torch.float32))
                gv = flat['g'][o:o + n].view_as(p)
                if p.grad is not None and p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
                    p.grad = gv
                _register_view(p, gv)
                views.append((p, pv.data_ptr(), gv, mv, vv))
        flat.update(params=params, views=views, total=off, uploaded=False, step=shared,
                    scratch=torch.empty(max(int(L.load().esr_adam_scratch_bytes(1)), 64), dtype=torch.uint8, device=dev))
        self._flat[gi] = flat
        return flat

    def _flat_for(self, gi, group):
        flat = self._flat.get(gi)
        params = self._cuda_params(group)
        if not params:
            return None
        ok = flat is not None and len(flat['params']) == len(params)
        if ok:      # a .to(device) / optimizer.load_state_dict moves tensors out of the flat buffers (all of them at once: the first and
            for k in (0, len(params) - 1):      # the last parameter are checked every step)
                p, pptr, gv, mv, vv = flat['views'][k]
                st = self.state[p]
                if p is not params[k] or p.data_ptr() != pptr or st.get('exp_avg') is None or st['exp_avg'].data_ptr() != mv.data_ptr() \
                        or (flat['step'] is not None and st.get('step') is not flat['step']):
                    ok = False
                    break
        return flat if ok else self._build(gi, group)

    def state_dict(self):
        """torch.optim.Adam's layout.  Internally every parameter of a group references ONE step counter; a checkpoint must not carry that
        sharing (torch's own Adam increments `step` once per parameter), so each entry gets its own copy here."""
        sd = super(FlatAdam, self).state_dict()
        for st in sd['state'].values():
            if 'step' in st and torch.is_tensor(st['step']):
                st['step'] = st['step'].clone()
        return sd

    def flat_grad(self, gi=0):
        """the flat gradient buffer of parameter group gi (None until the first step / register()), for the all-reduce"""
        flat = self._flat.get(gi)
        return flat['g'] if flat is not None else None

    def register(self):
        """build the flat buffers now (otherwise on the first step), so that the first backward already writes into them"""
        for gi, group in enumerate(self.param_groups):
            self._flat_for(gi, group)
        return self

    def grads_in_place(self, gi=0):
        """True when every parameter's .grad IS its flat view (or is None): the flat buffer can be reduced / stepped directly"""
        flat = self._flat.get(gi)
        if flat is None:
            return False
        return all(p.grad is None or p.grad is gv or p.grad.data_ptr() == gv.data_ptr() for p, _, gv, _, _ in flat['views'])

    # ---------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        cpu_groups = []
        lib = L.load() if any(p.is_cuda for g in self.param_groups for p in g['params'][:1]) else None
        for gi, group in enumerate(self.param_groups):
            if group.get('amsgrad') or group.get('maximize'):
                raise NotImplementedError('FlatAdam: amsgrad / maximize are not built')
            params = self._cuda_params(group)
            if len(params) != len([p for p in group['params'] if p.requires_grad]):
                cpu_groups.append(group)      # CPU (or non-fp32) parameters: torch's own arithmetic
                continue
            if not params:
                continue
            flat = self._flat_for(gi, group)
            n_missing = 0
            for p, _, gv, _, _ in flat['views']:
                g = p.grad
                if g is None:
                    n_missing += 1
                elif g is not gv and g.data_ptr() != gv.data_ptr():
                    gv.copy_(g)               # a gradient produced elsewhere (autograd of a torch op): one copy into its view
                    p.grad = gv
            if n_missing == len(flat['views']):
                continue
            beta1, beta2 = group['betas']
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            hyper = (float(group['lr']), float(beta1), float(beta2), float(group['eps']), float(group['weight_decay']))
            if n_missing == 0 and flat['step'] is not None:
                t = int(flat['step']) + 1
                arg = None
                if not flat['uploaded']:
                    arg = (L.AdamTensor * 1)()
                    arg[0].p, arg[0].g, arg[0].m, arg[0].v, arg[0].n = (flat['p'].data_ptr(), flat['g'].data_ptr(), flat['m'].data_ptr(),
                                                                        flat['v'].data_ptr(), flat['total'])
                    flat['uploaded'] = True
                L.check(lib.esr_adam_multi(arg, 1, C.c_void_p(flat['scratch'].data_ptr()), flat['scratch'].numel(), *hyper, t, float(grad_scale), stream))
                flat['step'] += 1
                # the kernel wrote the parameters behind autograd's back: bump their version counters (the engines re-pack their
                # tensor-core weight images when a parameter's version changes, exactly as after torch.optim.Adam's in-place update)
                torch.autograd.graph.increment_version(flat['params'])
            else:
                # torch skips parameters without a gradient (a zero gradient would still decay their moments), and a loaded checkpoint
                # may hold unequal step counts: one table entry per tensor that has a gradient, grouped by step count
                if flat['step'] is not None:        # un-share the counter: the parameters' counts are about to differ
                    for p, _, _, _, _ in flat['views']:
                        self.state[p]['step'] = flat['step'].clone()
                    flat['step'] = None
                by_step = {}
                for p, _, gv, mv, vv in flat['views']:
                    if p.grad is not None:
                        by_step.setdefault(float(self.state[p]['step']), []).append((p, gv, mv, vv))
                for t0, ents in by_step.items():
                    tab = (L.AdamTensor * len(ents))()
                    for k, (p, gv, mv, vv) in enumerate(ents):
                        tab[k].p, tab[k].g, tab[k].m, tab[k].v, tab[k].n = p.data_ptr(), gv.data_ptr(), mv.data_ptr(), vv.data_ptr(), p.numel()
                    scratch = torch.empty(int(lib.esr_adam_scratch_bytes(len(ents))), dtype=torch.uint8, device=flat['p'].device)
                    L.check(lib.esr_adam_multi(tab, len(ents), C.c_void_p(scratch.data_ptr()), scratch.numel(), *hyper, int(t0) + 1, float(grad_scale), stream))
                    for p, _, _, _ in ents:
                        self.state[p]['step'] += 1
                    torch.autograd.graph.increment_version([p for p, _, _, _ in ents])
        if cpu_groups:
            if float(grad_scale) != 1.0:
                for group in cpu_groups:
                    for p in group['params']:
                        if p.grad is not None:
                            p.grad.mul_(grad_scale)
            saved = self.param_groups
            self.param_groups = cpu_groups
            try:
                super(FlatAdam, self).step()
            finally:
                self.param_groups = saved
        return loss
