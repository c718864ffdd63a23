"""Adam / RAdam + weight running average as ONE fused multi-tensor kernel (libb200lp `b200lp_adam_ema_multi`).

Drop-in `torch.optim.Optimizer` (same constructor arguments as `torch.optim.Adam` / the reference's vendored RAdam,
same `state_dict()` layout: per-parameter `step`, `exp_avg`, `exp_avg_sq`), used by runners/holycow.py when the
parameters live on a CUDA device.  The step counter and its derived scalars are kept on the device, so an entire
training step including both optimizer updates can be captured in a CUDA graph and replayed.

Replaces (reference): `torch.optim.Adam.step`, `utils/radam.py:29-95`, and the parameter loop of
`TrainingModule.update_running_average` (`runners/holycow.py:99-105`).
"""
import torch
from torch.optim.optimizer import Optimizer

CHUNK = 65536


class FusedAdamEMA(Optimizer):
    MODE = 0   # torch.optim.Adam semantics

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid optimizer hyper-parameters lr={lr} eps={eps} betas={betas}")
        if weight_decay != 0:
            raise NotImplementedError("the fused optimizer implements weight_decay=0 (the value the reference uses)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.degenerated_to_sgd = degenerated_to_sgd
        self._tables = None
        self._ema_of = {}
        self.ema_alpha = 1.0

    def _hyper(self):
        """The one set of hyper-parameters the kernel runs with.  Several param groups are fine as long as they agree
        (the reference builds a single group, runners/holycow.py:34-41); disagreeing groups are refused instead of
        being silently treated as group 0."""
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            for k in ('lr', 'betas', 'eps', 'weight_decay'):
                if g[k] != g0[k]:
                    raise NotImplementedError(f"fused optimizer: param groups disagree on {k} ({g[k]} vs {g0[k]})")
        if g0['weight_decay'] != 0:
            raise NotImplementedError("the fused optimizer implements weight_decay=0 (the value the reference uses)")
        return float(g0['lr']), float(g0['betas'][0]), float(g0['betas'][1]), float(g0['eps'])

    def sync_hyper(self):
        """Upload lr / ema_alpha to the device state vector when they changed (an LR schedule, the runner switching the
        running-average alpha).  The kernels read them from there, so a captured CUDA graph of the step follows the
        change; call between replays (never while capturing)."""
        t = self._tables
        if t is None:
            return
        want = (self._hyper()[0], float(self.ema_alpha))
        if t.get('hyper') != want:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("fused optimizer: lr / ema_alpha changed while a CUDA graph is being captured")
            t['state'][4:6].copy_(torch.tensor(want, dtype=torch.float32))
            t['hyper'] = want

    # ------------------------------------------------------------------ wiring
    def attach_ema(self, pairs, alpha):
        """pairs: iterable of (parameter, running-average tensor) — updated inside the same kernel."""
        self._ema_of = {id(p): e for p, e in pairs}
        self.ema_alpha = float(alpha)
        self._tables = None

    def _params(self):
        return [p for g in self.param_groups for p in g['params'] if p.requires_grad]

    def _build(self):
        params = self._params()
        if not params or not params[0].is_cuda:
            from b200lp.lib import B200lpError
            raise B200lpError("FusedAdamEMA.step needs CUDA parameters (the hot path has no CPU fallback)")
        dev = params[0].device
        total = sum(p.numel() for p in params)
        m_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        v_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        old_step = 0.0
        rows, chunk_t, chunk_o = [], [], []
        off = 0
        for i, p in enumerate(params):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            n = p.numel()
            m, v = m_flat[off:off + n].view_as(p), v_flat[off:off + n].view_as(p)
            st = self.state[p]
            if 'exp_avg' in st:          # restored from a checkpoint (or a previous table): carry the moments over
                m.copy_(st['exp_avg'])
                v.copy_(st['exp_avg_sq'])
                old_step = max(old_step, float(st['step']))
            ema = self._ema_of.get(id(p))
            rows.append([p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                         ema.data_ptr() if ema is not None else 0, n])
            for o in range(0, n, CHUNK):
                chunk_t.append(i)
                chunk_o.append(o)
            off += n
        hyper = (self._hyper()[0], float(self.ema_alpha))
        state_dev = torch.tensor([old_step, 0.0, 0.0, 1.0, hyper[0], hyper[1], 0.0, 0.0], dtype=torch.float32, device=dev)
        for p in params:
            self.state[p]['step'] = state_dev[0]     # in-memory alias of the ONE device counter; see state_dict()
        off = 0
        for p in params:
            n = p.numel()
            self.state[p]['exp_avg'] = m_flat[off:off + n].view_as(p)
            self.state[p]['exp_avg_sq'] = v_flat[off:off + n].view_as(p)
            off += n
        self._tables = dict(
            params=params, grad_ptrs=[p.grad.data_ptr() for p in params], m=m_flat, v=v_flat, state=state_dev,
            table=torch.tensor(rows, dtype=torch.int64, device=dev),
            chunk_t=torch.tensor(chunk_t, dtype=torch.int32, device=dev),
            chunk_o=torch.tensor(chunk_o, dtype=torch.int64, device=dev), n_chunks=len(chunk_t), hyper=hyper)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = None              # moments are re-packed into the flat buffers at the next step

    def state_dict(self):
        """Same layout as torch.optim.Adam / the reference's RAdam — and NOT aliased: in memory every parameter's `step`
        is a view of one device counter and the moments are views into two flat buffers; a checkpoint written that way
        and loaded into an optimizer that does `state['step'] += 1` per parameter would advance the shared counter
        len(params) times per step.  Here every parameter gets its own Python-number `step` and cloned moments."""
        sd = super().state_dict()
        step = None
        out_state = {}
        for idx, st in sd['state'].items():
            new = dict(st)
            if torch.is_tensor(new.get('step')):
                if step is None:
                    step = int(round(float(new['step'])))
                new['step'] = step
            for k in ('exp_avg', 'exp_avg_sq'):
                if torch.is_tensor(new.get(k)):
                    new[k] = new[k].detach().clone()
            out_state[idx] = new
        return {'state': out_state, 'param_groups': sd['param_groups']}

    def snapshot(self):
        """(step, exp_avg, exp_avg_sq flat copies) of the live tables — GraphedTrainStep restores this after its
        side-effect-free warm-up."""
        if self._tables is None:
            self._build()            # also carries over moments / step restored from a checkpoint
        t = self._tables
        return t['state'][:4].clone(), t['m'].clone(), t['v'].clone()

    def restore(self, snap, fresh_state=None):
        """Inverse of snapshot(): the tables may have been rebuilt since (new gradient buffers), the parameter order —
        and with it the flat layout of the moments — is the same."""
        t = self._tables
        if t is None or snap is None:
            return
        t['state'][:4].copy_(snap[0])
        t['m'].copy_(snap[1])
        t['v'].copy_(snap[2])

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        t = self._tables
        if t is None or any(p.grad is None or p.grad.data_ptr() != gp for p, gp in zip(t['params'], t['grad_ptrs'])):
            self._build()
            t = self._tables
        from b200lp import lib as L
        from b200lp import ops
        _, beta1, beta2, eps = self._hyper()
        self.sync_hyper()
        lib = L.load()
        L.check(lib.b200lp_adam_ema_multi(
            L.c_void_p(t['table'].data_ptr()), L.c_void_p(t['chunk_t'].data_ptr()), L.c_void_p(t['chunk_o'].data_ptr()),
            t['n_chunks'], CHUNK, L.ptr(t['state']), beta1, beta2, eps, self.MODE, int(self.degenerated_to_sgd),
            L.stream_ptr()), "adam_ema_multi")
        ops.bump_generation(t['params'])  # parameters changed behind torch's back: their packed copies are stale
        ops.repack_weights(t['params'], owner=id(self))   # ... and refreshed here by one multi-tensor launch
        return loss


class FusedRAdamEMA(FusedAdamEMA):
    MODE = 1   # utils/radam.py semantics (rectified step once N_sma >= 5, momentum-SGD step before)
