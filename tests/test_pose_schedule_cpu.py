"""The pose-encoder kernel SCHEDULE (embedders/mobilenet_native.py) on the CPU: every libb200lp entry point it calls is
replaced here by a plain-torch emulation of that kernel's contract (include/b200lp.h), so the test checks the host
logic — which layer's BatchNorm is applied by which consumer, where ReLU6 sits, skip connections, statistics counts,
running-statistics updates — against the torchvision module itself.  (The kernels are checked against torch on the GPU:
tools/gpu_diag.py `pose_encoder`.)"""
import copy

import pytest
import torch
import torch.nn.functional as F


def _act(x, scale, shift, relu6):
    if scale is None:
        return x
    y = x * scale + shift
    return y.clamp(0, 6) if relu6 else y


def _stats(y2d):
    return torch.stack([y2d.sum(0), (y2d * y2d).sum(0)])[None]          # (1 part, 2, C)


class _Emu:
    """torch emulations of b200lp.kernels' pose-encoder wrappers (same signatures, same return conventions)."""

    @staticmethod
    def mbv2_stem(x_nchw, weight, want_stats=False):
        y = F.conv2d(x_nchw, weight, stride=2, padding=1).permute(0, 2, 3, 1).contiguous()
        return (y, _stats(y.reshape(-1, 32))) if want_stats else y

    @staticmethod
    def pw_conv(x2d, weight, in_scale=None, in_shift=None, in_relu6=False, bias=None, want_stats=False):
        y = _act(x2d, in_scale, in_shift, in_relu6) @ weight.t()
        part = _stats(y) if want_stats else None
        if bias is not None:
            y = y + bias
        return (y, part) if want_stats else y

    @staticmethod
    def dw_conv3x3(x, weight, in_scale, in_shift, stride, want_stats=False):
        a = _act(x, in_scale, in_shift, True).permute(0, 3, 1, 2)
        y = F.conv2d(a, weight, stride=stride, padding=1, groups=x.shape[-1]).permute(0, 2, 3, 1).contiguous()
        return (y, _stats(y.reshape(-1, y.shape[-1]))) if want_stats else y

    @staticmethod
    def bn_finalize(bn, part, count, training):
        if training:
            s1, s2 = part[:, 0].sum(0), part[:, 1].sum(0)
            mean = s1 / count
            var = (s2 / count - mean * mean).clamp_min(0)
            if bn.track_running_stats and bn.running_mean is not None:
                with torch.no_grad():
                    bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * mean)
                    bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * var * count / max(count - 1, 1))
                    bn.num_batches_tracked += 1
        else:
            mean, var = bn.running_mean, bn.running_var
        scale = bn.weight.detach() / torch.sqrt(var + bn.eps)
        return scale, bn.bias.detach() - mean * scale

    @staticmethod
    def bn_apply(x, scale, shift, residual=None, relu6=False):
        y = _act(x, scale, shift, relu6)
        return y + residual if residual is not None else y

    @staticmethod
    def bn_relu6_avgpool(x, scale, shift):
        return _act(x, scale, shift, True).mean((1, 2))


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_native_schedule_matches_torchvision(monkeypatch, mode):
    import torchvision
    from b200lp import kernels as K
    from embedders import mobilenet_native
    for name in ("mbv2_stem", "pw_conv", "dw_conv3x3", "bn_finalize", "bn_apply", "bn_relu6_avgpool"):
        monkeypatch.setattr(K, name, getattr(_Emu, name))
    torch.manual_seed(0)
    net = torchvision.models.mobilenet_v2(num_classes=24).double()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 2.0)
    net.classifier[0].p = 0.0                      # dropout is RNG-dependent (it stays a torch op in the schedule)
    assert mobilenet_native.supported(net)
    a, b = copy.deepcopy(net), copy.deepcopy(net)
    a.train(mode == "train"); b.train(mode == "train")
    x = torch.rand(3, 3, 64, 64, dtype=torch.float64)
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)     # keep float64 through the schedule's `.float()`
    with torch.no_grad():
        ya = mobilenet_native.forward(a, x)
        yb = b(x)
    torch.testing.assert_close(ya, yb, rtol=1e-9, atol=1e-10)
    for p, q in zip(a.modules(), b.modules()):
        if isinstance(p, torch.nn.BatchNorm2d):
            torch.testing.assert_close(p.running_mean, q.running_mean, rtol=1e-9, atol=1e-11)
            torch.testing.assert_close(p.running_var, q.running_var, rtol=1e-9, atol=1e-11)
            assert int(p.num_batches_tracked) == int(q.num_batches_tracked) == (1 if mode == "train" else 0)


def _reference_pose_golden():
    from conftest import GOLDEN
    return torch.load(GOLDEN / "pose.pt", map_location="cpu", weights_only=False)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_native_schedule_matches_reference_golden(monkeypatch, mode):
    """The schedule (over the kernel emulations, float64) against the UNMODIFIED reference embedder's pose path on the
    deterministic weights of oracle/synth.py (tests/golden/pose.pt, oracle/make_golden_pose.py)."""
    from b200lp import kernels as K
    from embedders import mobilenet_native
    from embedders.unsupervised_pose_separate_embResNeXt_segmentation import Embedder
    from oracle import synth
    gold = _reference_pose_golden()
    for name in ("mbv2_stem", "pw_conv", "dw_conv3x3", "bn_finalize", "bn_apply", "bn_relu6_avgpool"):
        monkeypatch.setattr(K, name, getattr(_Emu, name))
    emb = Embedder(16, gold["num_classes"], "sum")
    emb.pose_encoder.load_state_dict(synth.pose_encoder_state_dict(gold["num_classes"], seed=7), strict=True)
    net = emb.pose_encoder.double()
    net.train(mode == "train")
    net.classifier[0].p = 0.0
    x = synth.pose_inputs(batch=3, image_size=128, seed=8)[:, 0].double()
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)
    with torch.no_grad():
        y = mobilenet_native.forward(net, x)
    ref = gold[f"{mode}.pose_embedding"].double()
    assert float((y - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    if mode == "train":
        bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
        rm, rv = torch.cat([m.running_mean for m in bns]), torch.cat([m.running_var for m in bns])
        assert float((rm - gold["train.running_mean"]).abs().max()) <= 1e-5 * float(gold["train.running_mean"].abs().max())
        assert float((rv - gold["train.running_var"]).abs().max()) <= 1e-5 * float(gold["train.running_var"].abs().max())
        assert int(bns[0].num_batches_tracked) == gold["train.num_batches_tracked"] == 1


def test_embedder_plugin_torch_path_matches_reference_golden():
    """The plugin's differentiable path (torch modules, used when a gradient flows through the encoder) is the same
    torchvision network as the reference's: identical outputs on identical weights."""
    from embedders.unsupervised_pose_separate_embResNeXt_segmentation import Embedder
    from oracle import synth
    gold = _reference_pose_golden()
    emb = Embedder(16, gold["num_classes"], "sum").eval()
    emb.pose_encoder.load_state_dict(synth.pose_encoder_state_dict(gold["num_classes"], seed=7), strict=True)
    d = {"pose_input_rgbs": synth.pose_inputs(batch=3, image_size=128, seed=8)}
    with torch.no_grad():
        emb.get_pose_embedding(d)
    torch.testing.assert_close(d["pose_embedding"], gold["eval.pose_embedding"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("use_sinks", [False, True])
def test_native_backward_matches_autograd(monkeypatch, use_sinks):
    """The hand-scheduled BACKWARD of the pose encoder (BatchNorm backward with recomputed ReLU6 masks, depthwise / 1x1 /
    stem weight gradients, residual wiring, gradient sinks) == torch autograd through the torchvision module, float64."""
    import torchvision
    import encoder_emulators
    from b200lp import ops
    from embedders import mobilenet_native
    encoder_emulators.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)
    torch.manual_seed(5)
    net = torchvision.models.mobilenet_v2(num_classes=24).double()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    net.classifier[0].p = 0.0
    a, b = copy.deepcopy(net).train(), copy.deepcopy(net).train()
    x = torch.rand(3, 3, 64, 64, dtype=torch.float64)
    wgt = torch.randn(3, 24, dtype=torch.float64)
    yb = b(x)
    (yb * wgt).sum().backward()
    ya = mobilenet_native.apply(a, x)
    torch.testing.assert_close(ya, yb, rtol=1e-9, atol=1e-10)
    if use_sinks:
        base = {n_: torch.randn_like(p) for n_, p in a.named_parameters()}
        bufs = {n_: base[n_].clone() for n_ in base}
        with ops.direct_grads({p.data_ptr(): bufs[n_] for n_, p in a.named_parameters()}):
            (ya * wgt).sum().backward()
        got = {n_: bufs[n_] - base[n_] for n_ in base}
        assert all(p.grad is None for p in a.parameters())
    else:
        (ya * wgt).sum().backward()
        got = {n_: p.grad for n_, p in a.named_parameters()}
    for (n_, _), q in zip(a.named_parameters(), b.parameters()):
        scale = float(q.grad.abs().max()) + 1e-30
        assert float((got[n_] - q.grad).abs().max()) <= 1e-7 * scale + 1e-12, (n_, float((got[n_] - q.grad).abs().max()), scale)


def test_native_backward_with_dropout_matches_autograd(monkeypatch):
    """With Dropout active the schedule draws the mask from torch's generator exactly where the module does (one
    F.dropout call on the pooled features), so the same seed gives the same outputs and gradients."""
    import torchvision
    import encoder_emulators
    from embedders import mobilenet_native
    encoder_emulators.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)
    torch.manual_seed(6)
    net = torchvision.models.mobilenet_v2(num_classes=16).double()
    a, b = copy.deepcopy(net).train(), copy.deepcopy(net).train()
    x = torch.rand(2, 3, 64, 64, dtype=torch.float64)
    torch.manual_seed(77)
    yb = b(x)
    yb.square().sum().backward()
    torch.manual_seed(77)
    ya = mobilenet_native.apply(a, x)
    ya.square().sum().backward()
    torch.testing.assert_close(ya, yb, rtol=1e-9, atol=1e-10)
    for (n_, p), q in zip(a.named_parameters(), b.parameters()):
        scale = float(q.grad.abs().max()) + 1e-30
        assert float((p.grad - q.grad).abs().max()) <= 1e-7 * scale + 1e-12, n_
