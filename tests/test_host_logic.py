"""Host-side logic of the drop-in boundary (no GPU): argument precedence, plugin loading, checkpoint round trip,
optimizer update rule, meters, the synthetic dataset contract."""
import importlib
import math
import os
import sys
from argparse import Namespace

import pytest
import torch

from helpers import make_args
from oracle import synth


def test_store_bool_and_parser_add():
    from utils.argparse_utils import MyArgumentParser
    p = MyArgumentParser(conflict_handler='resolve')
    p.add('--flag', action='store_bool', default=True)
    assert p.parse_args([]).flag is True
    assert p.parse_args(['--no-flag']).flag is False
    assert p.parse_args(['--no-flag', '--flag']).flag is True


def test_meter_average_last_and_nan():
    from utils.utils import Meter
    m = Meter()
    m.add('a', 1.0); m.add('a', 3.0); m.add('a', float('nan'))
    assert m.get_average('a') == 2.0 and m.get_num_measurements('a') == 2 and math.isnan(m.get_last('a'))
    n = Meter(); n.add('a', 5.0, 2)
    m += n
    assert m.get_average('a') == (1 + 3 + 10) / 4


def test_radam_matches_reference_update_rule():
    """utils/radam.py (multi-tensor) against a per-tensor port of the reference's vendored RAdam (oracle)."""
    from oracle.cpu_step import RAdamPort
    from utils.radam import RAdam
    torch.manual_seed(0)
    a = [torch.randn(5, 3, requires_grad=True), torch.randn(7, requires_grad=True)]
    b = [t.detach().clone().requires_grad_(True) for t in a]
    oa = RAdam(a, lr=5e-4, betas=(0.0, 0.999), eps=1e-5)
    ob = RAdamPort(b, lr=5e-4, betas=(0.0, 0.999), eps=1e-5)
    for step in range(8):                       # crosses the N_sma >= 5 switch (step 6 for beta2 = 0.999)
        for x, y in zip(a, b):
            g = torch.randn_like(x)
            x.grad = g.clone(); y.grad = g.clone()
        oa.step(); ob.step()
        for x, y in zip(a, b):
            torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-7)
    assert set(oa.state[a[0]].keys()) == {'step', 'exp_avg', 'exp_avg_sq'}     # checkpoint-compatible state keys


def test_synthetic_dataset_contract():
    from dataloaders.synthetic import Dataset
    args = Namespace(synthetic_num_samples=8, synthetic_num_identities=4, n_frames_for_encoder=3, image_size=32,
                     finetune=False, random_seed=123)
    ds = Dataset.get_dataset(args, 'train')
    assert args.num_labels == 4 and len(ds) == 8
    d, t = ds[5]
    assert d['enc_rgbs'].shape == (3, 3, 32, 32) and d['pose_input_rgbs'].shape == (1, 3, 32, 32)
    assert d['target_rgbs'].shape == (1, 3, 32, 32) and t['real_segm'].shape == (1, 3, 32, 32)
    assert set(t['real_segm'].unique().tolist()) <= {0.0, 1.0} and t['label'] == 1
    assert float((d['target_rgbs'] * (1 - t['real_segm'])).abs().max()) == 0.0        # target = image * mask
    d2, _ = ds[5]
    assert torch.equal(d['enc_rgbs'], d2['enc_rgbs'])                                  # deterministic
    args.finetune = True
    ds = Dataset.get_dataset(args, 'train')
    assert args.num_labels == 1 and ds[3][0]['enc_rgbs'].shape[0] == 1 and ds[3][1]['label'] == 0


def test_argument_precedence_and_plugin_loading(tmp_path, monkeypatch):
    """defaults < checkpoint args < yaml < command line (reference README.md:61-66, utils/utils.py:48-53)."""
    import train
    from utils import utils as U
    monkeypatch.chdir(tmp_path)
    (tmp_path / 'configs').mkdir()
    (tmp_path / 'configs' / 'mini.yaml').write_text(
        "generator: vector_pose_unsupervised_segmentation_noBottleneck\n"
        "embedder: unsupervised_pose_separate_embResNeXt_segmentation\ndiscriminator: no_landmarks\n"
        "criterions: adversarial, dice\ndataloader: synthetic\nrunner: holycow\nlr_gen: 0.5\nperc_weight: 0.25\n"
        "dice_weight: 7\n")
    ckpt = {'args': Namespace(lr_gen=0.125, dice_weight=3.0, fm_weight=99.0, device='cpu')}
    torch.save(ckpt, tmp_path / 'c.pth')
    monkeypatch.setattr(U, 'CONFIG_DIRS', [tmp_path / 'configs'])
    monkeypatch.setattr(sys, 'argv', ['train.py', '--config', 'mini', '--checkpoint_path', str(tmp_path / 'c.pth'),
                                      '--dice_weight', '11', '--device', 'cpu'])
    args, default_args, m, ck = U.get_args_and_modules(train.build_parser())
    assert args.dice_weight == 11.0          # command line wins
    assert args.lr_gen == 0.5                # yaml beats checkpoint
    assert default_args.lr_gen == 5e-5 or default_args.lr_gen == 0.5   # parse_args([]) after set_defaults
    assert [w.__module__ for w in m['criterion_list']] == ['criterions.adversarial', 'criterions.dice']
    assert m['runner'].__name__ == 'runners.holycow' and ck is not None
    assert args.experiment_name == 'mini'


def test_checkpoint_round_trip_and_finetune_loading(tmp_path):
    """save_model -> load_model_from_checkpoint keeps weights; entering fine-tuning changes the module structure the
    way the reference does (identity_embedding parameter, 1-row discriminator embedding, fresh optimizers)."""
    from utils import utils as U
    cfg = synth.SMALL_CFG
    runner = importlib.import_module('runners.holycow')
    args = make_args(cfg, generator='vector_pose_unsupervised_segmentation_noBottleneck', discriminator='no_landmarks',
                     embedder='unsupervised_pose_separate_embResNeXt_segmentation', runner='holycow',
                     rank=0, iteration=7, experiment_dir=str(tmp_path), inference=False)
    G = importlib.import_module('generators.vector_pose_unsupervised_segmentation_noBottleneck').Wrapper.get_net(args)
    D = importlib.import_module('discriminators.no_landmarks').Wrapper.get_net(args)

    class TinyEmbedder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(3))
            self.finetuning = False

        def enable_finetuning(self, data_dict=None):
            self.finetuning = True

    mod = type(sys)('embedders.tiny_test_embedder')
    mod.Wrapper = type('Wrapper', (), {'get_args': staticmethod(lambda p: None),
                                       'get_net': staticmethod(lambda a: TinyEmbedder())})
    sys.modules['embedders.tiny_test_embedder'] = mod
    args.embedder = 'tiny_test_embedder'
    E = TinyEmbedder()
    G.load_state_dict(synth.generator_state_dict(cfg, seed=1))
    D.load_state_dict(synth.discriminator_state_dict(cfg, seed=2))
    tm = runner.TrainingModule(E, G, D, [], [], {})
    opt_G = runner.get_optimizer(E, G, args)
    opt_D = importlib.import_module('discriminators.no_landmarks').Wrapper.get_optimizer(D, args)
    path = U.save_model(tm, opt_G, opt_D, args)
    assert os.path.basename(path) == 'model_00000007.pth'
    assert os.path.basename(U.save_model(tm, opt_G, opt_D, args)) == 'model_00000007_0.pth'   # never overwrite
    ck = U.load_checkpoint_file(path)
    assert set(ck) == {'embedder', 'generator', 'discriminator', 'optimizer_G', 'optimizer_D', 'running_averages', 'args'}
    E2, G2, D2, ra, saved_args, oG, oD = U.load_model_from_checkpoint(ck, args)
    for (k, a), (_, b) in zip(G.state_dict().items(), G2.state_dict().items()):
        assert torch.equal(a, b), k
    assert set(ra) == {'embedder', 'generator'}
    ft_args = make_args(cfg, finetune=True, generator=args.generator, discriminator=args.discriminator,
                        embedder=args.embedder, runner='holycow', inference=False)
    E3, G3, D3, _, _, oG3, _ = U.load_model_from_checkpoint(U.load_checkpoint_file(path), ft_args)
    assert G3.finetuning and D3.finetuning and E3.finetuning
    assert 'identity_embedding' in G3.state_dict() and D3.embed.weight_orig.shape == (1, cfg['embed_channels'])
    assert torch.equal(G3.state_dict()['constant.constant'], G.state_dict()['constant.constant'])
    assert len(oG3.state) == 0                                                              # optimizer state not loaded
    inf_args = make_args(cfg, finetune=True, inference=True, generator=args.generator, discriminator=args.discriminator,
                         embedder=args.embedder, runner='holycow')
    *_, oG4, oD4 = U.load_model_from_checkpoint(U.load_checkpoint_file(path), inf_args)
    assert oG4 is None and oD4 is None


def test_training_module_ema_and_flags():
    runner = importlib.import_module('runners.holycow')
    E, G, D = torch.nn.Linear(2, 2), torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
    tm = runner.TrainingModule(E, G, D, [], [], {})
    assert not any(p.requires_grad for p in tm.running_averages['generator'].parameters())
    w0 = G.weight.detach().clone()
    with torch.no_grad():
        G.weight.add_(1.0)
    tm.update_running_average(0.9)
    torch.testing.assert_close(tm.running_averages['generator'].weight, w0 * 0.9 + (w0 + 1.0) * 0.1)
    with tm.set_use_running_averages():
        assert tm.use_running_averages
    assert not tm.use_running_averages
    tm.set_compute_losses(False)
    assert tm.compute_losses is False
    assert runner.TrainingModule(E, G, D, [], [], None).running_averages == {}


def test_grad_bucket_views():
    runner = importlib.import_module('runners.holycow')
    lin = torch.nn.Linear(3, 2)
    bucket = runner.GradBucket(lin.parameters())
    lin(torch.ones(1, 3)).sum().backward()
    assert bucket.flat.abs().sum() > 0 and lin.weight.grad.data_ptr() == bucket.flat.data_ptr()
    lin.weight.grad = None                       # e.g. optimizer.zero_grad(set_to_none=True)
    bucket.zero()
    assert lin.weight.grad is not None and float(bucket.flat.abs().sum()) == 0.0


def test_inv_sigma_edge_matches_autograd_of_torch_formula():
    """ops.InvSigmaFn (the autograd edge given to a batched-kernel 1/sigma) against autograd through torch's
    spectral-norm formula sigma = u^T W v with u, v constants (torch/nn/utils/spectral_norm.py compute_weight)."""
    from b200lp import ops
    torch.manual_seed(1)
    w = torch.randn(6, 4, 3, 3, dtype=torch.float64, requires_grad=True)
    u = torch.nn.functional.normalize(torch.randn(6, dtype=torch.float64), dim=0)
    v = torch.nn.functional.normalize(torch.randn(36, dtype=torch.float64), dim=0)
    c = torch.randn(1, dtype=torch.float64)
    sigma = torch.dot(u, torch.mv(w.reshape(6, -1), v))
    ((1.0 / sigma).reshape(1) * c).sum().backward()
    ref = w.grad.clone()
    w.grad = None
    s_const = (1.0 / sigma.detach()).reshape(1)
    (ops.inv_sigma_edge(w, s_const, u, v) * c).sum().backward()
    torch.testing.assert_close(w.grad, ref, rtol=1e-12, atol=1e-14)


def test_direct_grads_context_and_cpu_buckets():
    """Gradient sinks are active only inside the context manager, nest, and are never offered for CPU buckets (the
    in-place accumulating kernels are CUDA-only; CPU runs keep plain autograd accumulation)."""
    from b200lp import ops
    from runners.holycow import GradBucket
    p = torch.nn.Parameter(torch.randn(4, 3))
    bucket = GradBucket([p])
    assert bucket.sinks() == {}
    fake_sink = {p.data_ptr(): p.grad}
    assert ops._sink(p) is None
    with ops.direct_grads(fake_sink):
        assert ops._sink(p) is p.grad
        assert ops._sink(p.detach()) is p.grad            # aliases of the parameter share its storage pointer
        assert ops._sink(torch.randn(2, 2)) is None
        with ops.direct_grads({}):
            assert ops._sink(p) is p.grad
    assert ops._sink(p) is None


def test_native_pose_schedule_recognises_mobilenet_v2_only():
    import torchvision
    from embedders import mobilenet_native
    assert mobilenet_native.supported(torchvision.models.mobilenet_v2(num_classes=16))
    assert not mobilenet_native.supported(torchvision.models.resnet18(num_classes=16))
    # on CPU (or when a gradient is needed) the embedder keeps the torch module: no silent native path without CUDA
    from embedders.unsupervised_pose_separate_embResNeXt_segmentation import Embedder
    e = Embedder.__new__(Embedder)
    torch.nn.Module.__init__(e)
    e.pose_encoder = torchvision.models.mobilenet_v2(num_classes=16)
    assert e._native_pose_path(torch.zeros(1, 3, 32, 32)) is False


def test_split_affine_matches_plain_slicing():
    """ops.SplitAffineFn (all AdaIN (gamma, beta) blocks from one autograd node) against plain column slicing of the
    projector output (reference assign_affine_params, generator :108-125): same views, same gradient."""
    from b200lp import ops
    torch.manual_seed(2)
    sizes = [8, 8, 4, 2]
    a = torch.randn(3, 2 * sum(sizes), dtype=torch.float64, requires_grad=True)
    b = a.detach().clone().requires_grad_(True)
    coefs = [torch.randn(3, c, dtype=torch.float64) for c in sizes for _ in range(2)]
    pairs = ops.split_affine(a, sizes)
    loss, off, k = 0.0, 0, 0
    ref = 0.0
    for i, c in enumerate(sizes):
        gamma, beta = pairs[i]
        assert gamma.stride(1) == 1 and gamma.stride(0) == beta.stride(0) == a.shape[1]
        torch.testing.assert_close(beta, b[:, off:off + c].detach())
        torch.testing.assert_close(gamma, b[:, off + c:off + 2 * c].detach())
        if i != 2:                                   # one pair left unused: its gradient must come back as zeros
            loss = loss + (gamma * coefs[k]).sum() + (beta.sin() * coefs[k + 1]).sum()
            ref = ref + (b[:, off + c:off + 2 * c] * coefs[k]).sum() + (b[:, off:off + c].sin() * coefs[k + 1]).sum()
        off += 2 * c
        k += 2
    loss.backward()
    ref.backward()
    torch.testing.assert_close(a.grad, b.grad)


def test_fused_optimizer_state_dict_is_not_aliased():
    """In memory every parameter's `step` of the fused optimizer is a view of one device counter and the moments are
    views into two flat buffers; state_dict() must hand out independent values, so that a checkpoint loaded into
    torch.optim.Adam / the vendored RAdam (which do `state['step'] += 1` per parameter) advances by ONE per step."""
    from utils.fused_optim import FusedAdamEMA
    from utils.radam import RAdam
    ps = [torch.nn.Parameter(torch.randn(n)) for n in (5, 7, 3, 11, 2)]
    opt = FusedAdamEMA(ps, lr=1e-3, betas=(0.0, 0.999), eps=1e-5)
    counter = torch.tensor([7.0, 0.0, 0.0, 1.0])
    m_flat, v_flat = torch.randn(28), torch.rand(28)
    off = 0
    for p in ps:       # what FusedAdamEMA._build leaves in self.state (here on the CPU: no kernel is involved)
        opt.state[p] = {"step": counter[0], "exp_avg": m_flat[off:off + p.numel()].view_as(p),
                        "exp_avg_sq": v_flat[off:off + p.numel()].view_as(p)}
        off += p.numel()
    sd = opt.state_dict()
    steps = [st["step"] for st in sd["state"].values()]
    assert steps == [7] * 5 and all(isinstance(s, int) for s in steps)
    ptrs = {st["exp_avg"].untyped_storage().data_ptr() for st in sd["state"].values()}
    assert len(ptrs) == 5 and m_flat.untyped_storage().data_ptr() not in ptrs
    for cls in (torch.optim.Adam, RAdam):
        other = cls(ps, lr=1e-3, betas=(0.0, 0.999), eps=1e-5)
        other.load_state_dict(sd)
        for p in ps:
            p.grad = torch.randn_like(p)
        other.step()
        assert all(float(other.state[p]["step"]) == 8.0 for p in ps), [float(other.state[p]["step"]) for p in ps]
    assert float(counter[0]) == 7.0
    back = FusedAdamEMA(ps, lr=1e-3, betas=(0.0, 0.999), eps=1e-5)
    back.load_state_dict(other.state_dict())          # and back: plain per-parameter state is accepted
    assert float(back.state[ps[0]]["step"]) == 8.0


def test_fused_optimizer_refuses_disagreeing_param_groups():
    from utils.fused_optim import FusedAdamEMA
    a, b = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(3))
    opt = FusedAdamEMA([{"params": [a]}, {"params": [b], "lr": 1e-2}], lr=1e-3)
    with pytest.raises(NotImplementedError, match="param groups disagree"):
        opt._hyper()
    same = FusedAdamEMA([{"params": [a]}, {"params": [b]}], lr=1e-3)
    assert same._hyper()[0] == 1e-3


def test_wgrad_batch_pad_fills_the_32_pixel_k_step():
    """kernels.wgrad_batch_pad: planes below 32 pixels need a batch that is a multiple of 32 / (H*W) (conv_wgrad.cu
    plan_wgrad: pn); ragged batches get all-zero samples, which leave the sum over pixels unchanged."""
    from b200lp.kernels import wgrad_batch_pad
    for n, h, w, want in [(1, 4, 4, 2), (3, 4, 4, 4), (2, 4, 4, 2), (8, 4, 4, 8), (5, 2, 2, 8), (1, 8, 4, 1), (3, 16, 16, 3)]:
        x, dy = torch.randn(n, h, w, 32), torch.randn(n, h, w, 64)
        xp, dyp = wgrad_batch_pad(x, dy)
        assert xp.shape == (want, h, w, 32) and dyp.shape == (want, h, w, 64), (n, h, w, xp.shape)
        assert xp.is_contiguous() and dyp.is_contiguous()
        if want == n:
            assert xp is x and dyp is dy                     # no copy when the batch already fits
        else:
            assert torch.equal(xp[:n], x) and torch.equal(dyp[:n], dy)
            assert not xp[n:].any() and not dyp[n:].any()
        # the weight gradient is a sum over samples: zero samples add nothing
        ref = torch.einsum("nhwi,nhwo->oi", x.double(), dy.double())
        got = torch.einsum("nhwi,nhwo->oi", xp.double(), dyp.double())
        assert torch.equal(ref, got)
