"""Training runner — drop-in for the reference's `runners/holycow.py` (get_args :18-31, get_optimizer :34-41,
TrainingModule :44-210, run_epoch :212-402): same module-level API, same step protocol

    forward E -> G -> D -> criteria;  loss_G.backward(retain_graph) ; [all-reduce] ; optimizer_G.step()
    optimizer_D.zero_grad() ; loss_D.backward() ; [all-reduce] ; optimizer_D.step() ; EMA of E/G weights

Differences that do not change results:
  * data parallelism: gradients live as views into two flat fp32 buckets (E+G, D); after each backward ONE
    NCCL all-reduce(AVG) over NVLink runs on the bucket that backward produced (the reference all-reduces all
    83.6 M parameters twice through apex.Reducer, including gradients it then discards — SURVEY.md §2a C1/C2);
  * the discriminator's fake-for-G pass runs on detached weights (`skip_discarded_wgrad`), because the reference
    zeroes those weight gradients before they are ever used (runners/holycow.py:247);
  * EMA and gradient zeroing are multi-tensor (_foreach) launches.
"""
import copy
import itertools
import logging
import os
import time

import torch
from torch import nn

from utils import utils
from utils.radam import RAdam
from utils.utils import Meter

torch.optim.RAdam = RAdam   # the reference installs its vendored RAdam the same way (:5-6)

logger = logging.getLogger('runner')


def get_args(parser):
    parser.add('--iteration', type=int, default=0, help="Optional iteration number to start from")
    parser.add('--log_frequency_loss', type=int, default=1)
    parser.add('--log_frequency_images', type=int, default=100)
    parser.add('--log_frequency_fixed_images', type=int, default=2500)
    parser.add('--detailed_metrics', action='store_bool', default=True)
    parser.add('--num_visuals_per_img', default=2, type=int)
    parser.add('--fixed_val_ids', action='append', type=int, default=[50, 100, 200, 250, 300])
    parser.add('--batch_size_inference', default=5, type=int)
    parser.add('--cuda_graph', action='store_bool', default=True,
               help="capture the optimisation step in a CUDA graph and replay it (static batch shapes)")
    return parser


def optimizer_class(name, device):
    """`torch.optim.<name>` as the reference resolves it (Adam, or RAdam = the vendored rule), or — for CUDA
    parameters — the fused multi-tensor kernel version with identical update rule and state layout."""
    if str(device).startswith('cuda'):
        from utils.fused_optim import FusedAdamEMA, FusedRAdamEMA
        fused = {'Adam': FusedAdamEMA, 'RAdam': FusedRAdamEMA}
        if name in fused:
            return fused[name]
    return torch.optim.__dict__[name]


def get_optimizer(embedder, generator, args):
    model_parameters = list(generator.parameters())
    if 'finetune' not in args or not args.finetune:
        model_parameters += list(embedder.parameters())
    Optimizer = optimizer_class(args.optimizer, args.device)
    return Optimizer(model_parameters, lr=args.lr_gen, betas=(args.beta1, 0.999), eps=1e-5)


class GradBucket:
    """All gradients of a parameter list as views into one flat fp32 buffer: zeroing is one memset, the data-parallel
    exchange is one in-place NCCL all-reduce(AVG) with no flatten / unflatten copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        device = self.params[0].device if self.params else 'cpu'
        dtype = self.params[0].dtype if self.params else torch.float32      # fp32 in production; fp64 in CPU tests
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def attach(self):
        """Re-point .grad at the bucket (an optimizer's zero_grad(set_to_none=True) or a fresh parameter drops it)."""
        off = 0
        for p in self.params:
            g = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
            off += p.numel()

    def zero(self):
        self.attach()
        self.flat.zero_()

    def sinks(self):
        """{parameter storage pointer: its gradient view} for b200lp.ops.direct_grads (CUDA buckets only)."""
        if not self.flat.is_cuda or os.environ.get('B200LP_NO_GRAD_SINKS'):   # env switch: A/B measurements only
            return {}
        self.attach()
        return {p.data_ptr(): p.grad for p in self.params}

    def all_reduce(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            self.attach()
            if self.flat.is_cuda:
                torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.AVG)
            else:   # gloo (CPU tests) has no AVG
                torch.distributed.all_reduce(self.flat, op=torch.distributed.ReduceOp.SUM)
                self.flat.div_(torch.distributed.get_world_size())


class TrainingModule(torch.nn.Module):
    def __init__(self, embedder, generator, discriminator, criterion_list, metric_list, running_averages={}):
        """`running_averages`: None (do not track) or {name: state_dict} initial values for 'embedder'/'generator'."""
        super().__init__()
        self.embedder = embedder
        self.generator = generator
        self.discriminator = discriminator
        self.criterion_list = nn.ModuleList(criterion_list)
        self.metric_list = nn.ModuleList(metric_list)
        self.compute_losses = True
        self.use_running_averages = False
        self.initialize_running_averages(running_averages)
        self._buckets = None

    def initialize_running_averages(self, initial_values={}):
        self.running_averages = {}
        if initial_values is not None:
            for name in 'embedder', 'generator':
                model = getattr(self, name)
                self.running_averages[name] = copy.deepcopy(model)
                if name in initial_values:
                    try:
                        self.running_averages[name].load_state_dict(initial_values[name])
                    except Exception:
                        logger.warning(f"Parameters mismatch in {name} and the initial value of weights' "
                                       f"running averages. Initializing by cloning")
                        self.running_averages[name].load_state_dict(model.state_dict())
                else:
                    logger.info(f"No initial value of weights' running averages provided for {name}. "
                                f"Initializing by cloning")
        for module in self.running_averages.values():
            module.eval()
            module.requires_grad_(False)

    def update_running_average(self, alpha=0.999):
        """p_avg = p_avg*alpha + p*(1-alpha) for E and G parameters; buffers copied (reference :99-109).
        Parameters whose running average is maintained inside the fused optimizer kernel (see grad_buckets) are
        skipped here — they were averaged in the same pass that updated them."""
        fused = getattr(self, '_ema_in_optimizer', set())
        with torch.no_grad():
            for name, avg in self.running_averages.items():
                cur = getattr(self, name)
                pairs = [(a, c) for a, c in zip(avg.parameters(), cur.parameters()) if id(c) not in fused]
                if pairs and not self._ema_kernel(name, pairs, alpha):
                    p_avg, p_cur = [a for a, _ in pairs], [c for _, c in pairs]
                    torch._foreach_mul_(p_avg, alpha)
                    torch._foreach_add_(p_avg, p_cur, alpha=1 - alpha)
                b_avg, b_cur = list(avg.buffers()), list(cur.buffers())
                if b_avg and b_avg[0].is_cuda:
                    self._copy_buffers(name, b_avg, b_cur)
                elif b_avg:
                    torch._foreach_copy_(b_avg, b_cur)

    def _ema_kernel(self, name, pairs, alpha):
        """Running average of the parameters the fused optimizer does not own (fine-tuning: the frozen embedder's 29 M) as
        ONE multi-tensor launch (12 B per parameter instead of two `_foreach` passes); False = not applicable here."""
        if not all(a.is_cuda and a.dtype == torch.float32 and a.is_contiguous() and c.is_contiguous() and a.shape == c.shape
                   and c.dtype == torch.float32 for a, c in pairs):
            return False
        from b200lp import kernels as K
        sig = tuple((c.data_ptr(), a.data_ptr(), a.numel()) for a, c in pairs)
        plans = self.__dict__.setdefault('_ema_plans', {})
        plan = plans.get(name)
        if plan is None or plan['sig'] != sig:
            if torch.cuda.is_current_stream_capturing():     # the table upload is not capturable
                return False
            plan = K.ema_plan(pairs)
            plans[name] = plan
        K.ema_multi(plan, alpha)
        return True

    def _copy_buffers(self, name, b_avg, b_cur):
        """All buffer copies of one module (BatchNorm statistics, spectral-norm vectors: ~370 tiny tensors for E + G) as
        ONE multi-tensor launch; the device-side table is cached while the buffers stay where they are."""
        from b200lp import kernels as K
        pairs = [(a, c) for a, c in zip(b_avg, b_cur)
                 if a.is_contiguous() and c.is_contiguous() and a.dtype == c.dtype and a.shape == c.shape]
        rest = [(a, c) for a, c in zip(b_avg, b_cur)
                if not (a.is_contiguous() and c.is_contiguous() and a.dtype == c.dtype and a.shape == c.shape)]
        for a, c in rest:
            a.copy_(c)
        if not pairs:
            return
        sig = tuple((a.data_ptr(), c.data_ptr(), a.numel() * a.element_size()) for a, c in pairs if a.numel())
        plans = self.__dict__.setdefault('_copy_plans', {})
        plan = plans.get(name)
        if plan is None or plan['sig'] != sig:
            if torch.cuda.is_current_stream_capturing():     # the table upload is not capturable
                torch._foreach_copy_([a for a, _ in pairs], [c for _, c in pairs])
                return
            plan = K.copy_plan(pairs)
            plans[name] = plan
        K.copy_multi(plan)

    class _Flag:
        def __init__(self, owner, attr, value):
            self.owner, self.attr, self.old = owner, attr, getattr(owner, attr)
            setattr(owner, attr, value)

        def __enter__(self):
            pass

        def __exit__(self, *args):
            setattr(self.owner, self.attr, self.old)

    def set_use_running_averages(self, use_running_averages=True):
        """Usable as a plain call or as a context manager (restores the old value on exit)."""
        return TrainingModule._Flag(self, 'use_running_averages', use_running_averages)

    def set_compute_losses(self, compute_losses=True):
        return TrainingModule._Flag(self, 'compute_losses', compute_losses)

    def forward(self, data_dict, target_dict):
        if self.running_averages and self.use_running_averages:
            embedder, generator = self.running_averages['embedder'], self.running_averages['generator']
        else:
            embedder, generator = self.embedder, self.generator

        data_dict = copy.copy(data_dict)
        embedder(data_dict)
        generator(data_dict)
        data_dict.update(target_dict)
        if self.compute_losses:
            self.discriminator(data_dict)

        losses_G_dict, losses_D_dict = {}, {}
        for criterion in self.criterion_list:
            try:
                crit_out = criterion(data_dict)
            except Exception:
                if self.compute_losses:
                    raise
                continue
            if isinstance(crit_out, tuple):
                if len(crit_out) != 2:
                    raise TypeError(f'Unexpected number of outputs in criterion {type(criterion)}: '
                                    f'expected 2, got {len(crit_out)}')
                losses_G_dict.update(crit_out[0])
                losses_D_dict.update(crit_out[1])
            elif isinstance(crit_out, dict):
                losses_G_dict.update(crit_out)
            else:
                raise TypeError(f'Unexpected type of {type(criterion)} output: '
                                f'expected dict or tuple of two dicts, got {type(crit_out)}')
        return data_dict, losses_G_dict, losses_D_dict

    def compute_metrics(self, data_dict):
        metrics_meter = Meter()
        for metric in self.metric_list:
            metric_out, num_errors = metric(data_dict)
            for name, value in metric_out.items():
                metrics_meter.add(name, value, num_errors[name])
        return metrics_meter

    # ---------------------------------------------------------------- data parallel plumbing
    def grad_buckets(self, optimizer_G, optimizer_D):
        """(bucket_G, bucket_D) over exactly the parameters each optimizer owns; built lazily, rebuilt when the
        optimizers are re-created (fine-tuning start)."""
        key = (id(optimizer_G), id(optimizer_D))
        if self._buckets is None or self._buckets[0] != key:
            params_G = [p for g in optimizer_G.param_groups for p in g['params']]
            params_D = [p for g in optimizer_D.param_groups for p in g['params']] if optimizer_D else []
            self._buckets = (key, GradBucket(params_G), GradBucket(params_D) if params_D else None)
            if hasattr(self.discriminator, 'skip_discarded_wgrad'):
                self.discriminator.skip_discarded_wgrad = True
            # running averages of the parameters optimizer_G owns ride along in its fused kernel
            self._ema_in_optimizer = set()
            if hasattr(optimizer_G, 'attach_ema') and self.running_averages:
                owned = {id(p) for p in params_G}
                pairs = []
                for name, avg in self.running_averages.items():
                    for a, c in zip(avg.parameters(), getattr(self, name).parameters()):
                        if id(c) in owned:
                            pairs.append((c, a))
                            self._ema_in_optimizer.add(id(c))
                optimizer_G.attach_ema(pairs, optimizer_G.ema_alpha)
        return self._buckets[1], self._buckets[2]

    def broadcast_parameters(self):
        """Rank 0's parameters and buffers to every rank (what apex.parallel.Reducer does at construction)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            for t in itertools.chain(self.parameters(), self.buffers()):
                torch.distributed.broadcast(t.data, src=0)


def train_step(training_module, data_dict, target_dict, optimizer_G, optimizer_D, finetune=False):
    """One optimisation step (reference run_epoch :230-257).  Returns (all_data_dict, losses_G, losses_D)."""
    bucket_G, bucket_D = training_module.grad_buckets(optimizer_G, optimizer_D)
    alpha = 0.972 if finetune else 0.999
    if hasattr(optimizer_G, 'ema_alpha'):
        optimizer_G.ema_alpha = alpha
    all_data_dict, losses_G_dict, losses_D_dict = training_module(data_dict, target_dict)
    loss_G = sum(v for v in losses_G_dict.values() if isinstance(v, torch.Tensor))
    loss_D = sum(v for v in losses_D_dict.values() if isinstance(v, torch.Tensor))

    from b200lp import ops
    bucket_G.zero()
    with ops.direct_grads(bucket_G.sinks()):      # conv weight / bias gradients are accumulated in place by the kernels
        loss_G.backward(retain_graph=True)
    # The generator-side exchange and update (NCCL all-reduce of the E+G bucket, fused Adam + EMA, weight re-packing) touch
    # only E / G parameters, their gradients and optimizer state; the discriminator's backward reads only D weights and
    # saved activations.  So the two run CONCURRENTLY: the first on a side stream forked here, the second on the main
    # stream; they join before the running-average buffers are copied.  (Captured as a fork / join inside the CUDA graph.)
    overlap = bool(losses_D_dict) and bucket_G.flat.is_cuda and not os.environ.get('B200LP_NO_OVERLAP')
    if overlap:
        main = torch.cuda.current_stream()
        side = _side_stream(bucket_G.flat.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            bucket_G.all_reduce()
            optimizer_G.step()
    else:
        bucket_G.all_reduce()
        optimizer_G.step()

    if losses_D_dict:
        bucket_D.zero()
        with ops.direct_grads(bucket_D.sinks()):
            loss_D.backward()
        bucket_D.all_reduce()
        optimizer_D.step()
    if overlap:
        main.wait_stream(side)

    training_module.update_running_average(alpha)
    return all_data_dict, losses_G_dict, losses_D_dict


_SIDE_STREAMS = {}


def _side_stream(device):
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class GraphedTrainStep:
    """The whole optimisation step (forward, both backward passes, gradient exchange, both optimizer updates, EMA)
    captured ONCE in a CUDA graph and replayed per batch: the step is ~7000 kernel launches, which is host-launch
    bound when issued one by one.  Everything in the step is capture-safe: libb200lp launches on the current stream
    and never allocates or synchronises, the fused optimizer keeps its step counter on the device, and the gradient
    all-reduce is a single in-place NCCL collective on a static buffer.

    Usage:  step = GraphedTrainStep(tm, opt_G, opt_D, finetune, example_data, example_target)
            all_data, losses_G, losses_D = step(data_dict, target_dict)      # tensors are static: read before next call
    Batches must keep the example's shapes (the reference's train loader uses drop_last=True)."""

    def __init__(self, training_module, optimizer_G, optimizer_D, finetune, data_dict, target_dict, warmup=3):
        self.static_data = {k: v.clone() for k, v in data_dict.items() if torch.is_tensor(v)}
        self.static_target = {k: v.clone() for k, v in target_dict.items() if torch.is_tensor(v)}
        self.optimizers = [o for o in (optimizer_G, optimizer_D) if o is not None]
        args = (training_module, self.static_data, self.static_target, optimizer_G, optimizer_D, finetune)
        # The warm-up steps exist only for lazy initialisation (gradient buckets, optimizer tables, packed weights, the
        # allocator's pools); they must not train: one batch = one update, like the reference (run_epoch :230-257).
        # Everything a step mutates is snapshotted here and put back after the capture — parameters and buffers of
        # E / G / D (BatchNorm statistics, spectral-norm vectors), the running averages, both optimizers' moments and
        # step counters, and the RNG streams (Dropout in the pose encoder).
        saved = self._snapshot(training_module)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):         # lazy initialisation (buckets, optimizer tables, packed VGG weights, ...)
                train_step(training_module, dict(self.static_data), dict(self.static_target), optimizer_G, optimizer_D,
                           finetune)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from b200lp import lib as b200lp_lib
        launched = b200lp_lib.load().b200lp_launch_count()
        from b200lp import kernels as b200lp_kernels
        self.graph = torch.cuda.CUDAGraph()
        b200lp_kernels.WORK = {}            # algorithmic FLOPs / bytes per kernel family of ONE step (host bookkeeping)
        try:
            with torch.cuda.graph(self.graph):
                self.outputs = train_step(training_module, dict(self.static_data), dict(self.static_target),
                                          optimizer_G, optimizer_D, finetune)
        finally:
            self.work, b200lp_kernels.WORK = b200lp_kernels.WORK, None
        # libb200lp kernels recorded in the graph = launched by every replay
        self.kernels_per_replay = int(b200lp_lib.load().b200lp_launch_count() - launched)
        self._restore(training_module, saved)

    def _snapshot(self, tm):
        mods = [tm] + list(tm.running_averages.values())
        tensors = []
        for m in mods:
            tensors += [t for t in itertools.chain(m.parameters(), m.buffers())]
        seen, uniq = set(), []
        for t in tensors:
            if t.data_ptr() not in seen and t.numel():
                seen.add(t.data_ptr())
                uniq.append(t)
        opt = []
        for o in self.optimizers:
            if hasattr(o, 'snapshot'):
                opt.append((o.snapshot(), 0.0, None))
            else:
                opt.append((None, 0.0, copy.deepcopy(o.state_dict())))
        rng = (torch.get_rng_state(), torch.cuda.get_rng_state())
        return [(t, t.detach().clone()) for t in uniq], opt, rng

    def _restore(self, tm, saved):
        from b200lp import ops
        tensors, opt, rng = saved
        with torch.no_grad():
            for t, v in tensors:
                t.copy_(v)
            for o, (snap, step0, sd) in zip(self.optimizers, opt):
                if hasattr(o, 'restore'):
                    o.restore(snap, fresh_state=step0)
                elif sd is not None:
                    o.load_state_dict(sd)
        torch.set_rng_state(rng[0])
        torch.cuda.set_rng_state(rng[1])
        ops.bump_generation()                 # weights went back to their pre-warm-up values: packed copies are stale
        for o in self.optimizers:             # refresh the packed copies whose addresses the graph holds
            t = getattr(o, '_tables', None)
            if t is not None:
                ops.repack_weights(t['params'], owner=id(o))
        torch.cuda.synchronize()

    def matches(self, data_dict, target_dict):
        """True when the batch has the captured shapes (a short last batch, a missing key ... -> eager step)."""
        for static, given in ((self.static_data, data_dict), (self.static_target, target_dict)):
            for k, v in static.items():
                g = given.get(k)
                if not torch.is_tensor(g) or g.shape != v.shape or g.dtype != v.dtype:
                    return False
        return True

    # ---------------------------------------------------------------- host -> device input staging
    def _stage(self, which):
        """Staging buffers (two sets, used alternately) for batches that arrive in HOST memory: the upload of batch i+1
        runs on a copy stream while the graph of batch i executes; at the start of step i+1 the graph's static inputs
        are refreshed from the staging set by a device-to-device copy (tens of microseconds)."""
        if getattr(self, '_staging', None) is None:
            self._staging = [None, None]
            self._consumed = [None, None]      # event: the compute stream finished reading that staging set
            self._copy_stream = torch.cuda.Stream()
            self._pending = None
        if self._staging[which] is None:
            self._staging[which] = ({k: torch.empty_like(v) for k, v in self.static_data.items()},
                                    {k: torch.empty_like(v) for k, v in self.static_target.items()})
        return self._staging[which]

    def prefetch(self, data_dict, target_dict):
        """Start the host -> device upload of a FUTURE batch on the copy stream (pinned host tensors make it
        asynchronous).  The next __call__ with these very dicts consumes it."""
        if not self.matches(data_dict, target_dict) or any(v.is_cuda for v in data_dict.values() if torch.is_tensor(v)):
            return
        which = 1 - getattr(self, '_last_stage', 1)
        sd, st = self._stage(which)
        with torch.cuda.stream(self._copy_stream):
            # the set being overwritten was last read by the device-to-device refresh of two steps ago: wait for THAT
            # copy only (not for the compute stream's current position — the replay in flight must keep overlapping)
            if self._consumed[which] is not None:
                self._copy_stream.wait_event(self._consumed[which])
            for k, v in sd.items():
                v.copy_(data_dict[k], non_blocking=True)
            for k, v in st.items():
                v.copy_(target_dict[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        # the dicts themselves (not their ids: a freed dict's id can be handed to the next batch's dict)
        self._pending = (data_dict, target_dict, which, ev)
        self._last_stage = which

    def release(self):
        """Destroy the captured graph and everything it keeps alive (its private memory pool, the captured NCCL
        all-reduces).  Must happen before torch.distributed.destroy_process_group()."""
        self.graph = None
        self.outputs = None
        self._staging = None
        self._consumed = [None, None]
        self._pending = None
        torch.cuda.synchronize()

    def __call__(self, data_dict, target_dict, prefetch_next=None):
        """Replay the step on a batch.  Device-resident batches are copied into the static inputs directly; host batches
        go through the prefetched staging set when `prefetch` was called for them (else an in-line upload).
        `prefetch_next` = (data_dict, target_dict) of the batch after this one: its upload overlaps this replay."""
        from b200lp import ops
        if self.graph is None:
            raise RuntimeError("GraphedTrainStep was released")
        if not self.matches(data_dict, target_dict):
            raise ValueError("GraphedTrainStep: batch does not have the captured shapes / keys "
                             f"({ {k: tuple(v.shape) for k, v in self.static_data.items()} })")
        for o in self.optimizers:            # lr / ema_alpha live in a device vector the captured kernels read
            if hasattr(o, 'sync_hyper'):
                o.sync_hyper()
        pend = getattr(self, '_pending', None)
        if pend is not None and pend[0] is data_dict and pend[1] is target_dict:
            sd, st = self._staging[pend[2]]
            torch.cuda.current_stream().wait_event(pend[3])
            for k, v in self.static_data.items():
                v.copy_(sd[k], non_blocking=True)
            for k, v in self.static_target.items():
                v.copy_(st[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._consumed[pend[2]] = ev
            self._pending = None
        else:
            for k, v in self.static_data.items():
                v.copy_(data_dict[k], non_blocking=True)
            for k, v in self.static_target.items():
                v.copy_(target_dict[k], non_blocking=True)
        self.graph.replay()
        if prefetch_next is not None:
            self.prefetch(*prefetch_next)
        ops.bump_generation()        # weights were rewritten by the replayed optimizer kernels
        return self.outputs


def _with_lookahead(iterable):
    """(item, next item or None) pairs: lets the epoch loop start the next batch's upload while this one computes."""
    it = iter(iterable)
    try:
        cur = next(it)
    except StopIteration:
        return
    for nxt in it:
        yield cur, nxt
        cur = nxt
    yield cur, None


def run_epoch(dataloader, training_module, optimizer_G, optimizer_D, epoch, args, phase, writer=None, saver=None):
    meter = Meter()
    end = time.time()
    use_graph = phase == 'train' and getattr(args, 'cuda_graph', False) and str(args.device).startswith('cuda')
    for it, ((data_dict, target_dict), upcoming) in enumerate(_with_lookahead(dataloader)):
        meter.add('Data_time', time.time() - end)

        if use_graph:
            key = (id(optimizer_G), id(optimizer_D))
            graphed = getattr(training_module, '_graphed_step', None)
            if graphed is None or graphed[0] != key:
                dev_data, dev_target = dict(data_dict), dict(target_dict)
                utils.dict_to_device(dev_data, args.device)
                utils.dict_to_device(dev_target, args.device)
                graphed = (key, GraphedTrainStep(training_module, optimizer_G, optimizer_D, args.finetune, dev_data,
                                                 dev_target))
                training_module._graphed_step = graphed
            if graphed[1].matches(data_dict, target_dict):
                # host batches: this batch's upload was prefetched during the previous step; start the next one's now
                all_data_dict, losses_G_dict, losses_D_dict = graphed[1](data_dict, target_dict, prefetch_next=upcoming)
            else:       # e.g. a dataset plugin's own loader without drop_last: run this batch eagerly
                utils.dict_to_device(data_dict, args.device)
                utils.dict_to_device(target_dict, args.device)
                all_data_dict, losses_G_dict, losses_D_dict = train_step(
                    training_module, data_dict, target_dict, optimizer_G, optimizer_D, finetune=args.finetune)
        elif phase == 'train':
            utils.dict_to_device(data_dict, args.device)
            utils.dict_to_device(target_dict, args.device)
            all_data_dict, losses_G_dict, losses_D_dict = train_step(
                training_module, data_dict, target_dict, optimizer_G, optimizer_D, finetune=args.finetune)
        else:
            utils.dict_to_device(data_dict, args.device)
            utils.dict_to_device(target_dict, args.device)
            all_data_dict, losses_G_dict, losses_D_dict = training_module(data_dict, target_dict)
            if saver is not None:
                saver.save(epoch=epoch, data=all_data_dict)

        if args.detailed_metrics:   # one device->host sync per loss, like the reference (:260-262)
            for loss_name, loss_ in itertools.chain(losses_G_dict.items(), losses_D_dict.items()):
                meter.add(f'Loss_{loss_name}', float(loss_))
        del losses_G_dict, losses_D_dict

        meter.add('Batch_time', time.time() - end)
        if writer is not None and phase == 'train':
            if args.iteration % args.log_frequency_loss == 0:
                for name in meter.keys():
                    writer.add_scalar(f'Metrics/{phase}/{name}', meter.get_last(name), args.iteration)
            if args.log_frequency_images > 0 and args.iteration % args.log_frequency_images == 0:
                try:
                    from utils.visualize import make_visual
                    writer.add_image(f'Images/{phase}/visual', make_visual(all_data_dict, args.num_visuals_per_img),
                                     args.iteration, dataformats='HWC')
                except Exception as err:
                    logger.warning(f"Could not log images: {err}")
            args.iteration += 1
        del all_data_dict
        end = time.time()

    logger.info(f"{phase.capitalize()} epoch {epoch}: " +
                ", ".join(f"{name}: {meter.get_average(name):.4g}" for name in meter.keys()))
    return meter
