"""The identity-encoder kernel SCHEDULE (embedders/resnext_native.py) on the CPU in float64: every libb200lp entry point
it calls is replaced by a plain-torch emulation of that kernel's contract (tests/encoder_emulators.py,
tests/kernel_emulators.py), so these tests check the host logic — which BatchNorm is applied by which consumer, stride-2
sub-sampling, residual / downsample wiring, statistics counts and running-statistics updates, and the whole hand-written
BACKWARD chain (BatchNorm backward with recomputed ReLU masks, grouped / 1x1 / stem weight gradients, max-pool gather,
gradient sinks) — against torchvision's ResNeXt50-32x4d and torch autograd.  The kernels themselves are compared with
the same emulations on the GPU (tools/gpu_diag.py `encoder_*`)."""
import copy

import pytest
import torch

import encoder_emulators
import kernel_emulators


def _net(num_classes=32, seed=0):
    import torchvision
    torch.manual_seed(seed)
    net = torchvision.models.resnext50_32x4d(num_classes=num_classes).double()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
                m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 2.0)
    return net


@pytest.fixture
def emulated(monkeypatch):
    kernel_emulators.install(monkeypatch)
    encoder_emulators.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "float", lambda self: self)     # keep float64 through the schedule's `.float()`
    return monkeypatch


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_forward_matches_torchvision(emulated, mode):
    from embedders import resnext_native
    net = _net()
    assert resnext_native.supported(net)
    a, b = copy.deepcopy(net), copy.deepcopy(net)
    a.train(mode == "train"); b.train(mode == "train")
    x = torch.rand(4, 3, 64, 64, dtype=torch.float64)
    with torch.no_grad():
        ya = resnext_native.apply(a, x)
        yb = b(x)
    torch.testing.assert_close(ya, yb, rtol=1e-9, atol=1e-10)
    for p, q in zip(a.modules(), b.modules()):
        if isinstance(p, torch.nn.BatchNorm2d):
            torch.testing.assert_close(p.running_mean, q.running_mean, rtol=1e-9, atol=1e-11)
            torch.testing.assert_close(p.running_var, q.running_var, rtol=1e-9, atol=1e-11)
            assert int(p.num_batches_tracked) == int(q.num_batches_tracked) == (1 if mode == "train" else 0)


@pytest.mark.parametrize("mode,use_sinks", [("train", False), ("train", True), ("eval", False)])
def test_backward_matches_autograd(emulated, mode, use_sinks):
    """Every parameter gradient of the hand-scheduled backward == torch autograd through the torchvision module
    (train-mode batch statistics, and eval-mode running statistics); with `direct_grads` sinks the gradients are
    ACCUMULATED into the given buffers and autograd receives None."""
    from b200lp import ops
    from embedders import resnext_native
    net = _net(seed=1)
    a, b = copy.deepcopy(net), copy.deepcopy(net)
    a.train(mode == "train"); b.train(mode == "train")
    x = torch.rand(4, 3, 64, 64, dtype=torch.float64)
    wgt = torch.randn(4, 32, dtype=torch.float64)
    yb = b(x)
    (yb * wgt).sum().backward()
    ya = resnext_native.apply(a, x)
    torch.testing.assert_close(ya, yb, rtol=1e-9, atol=1e-10)
    if use_sinks:
        base = {n_: torch.randn_like(p) for n_, p in a.named_parameters()}      # "gradient accumulated so far"
        bufs = {n_: base[n_].clone() for n_ in base}
        sinks = {p.data_ptr(): bufs[n_] for n_, p in a.named_parameters()}
        with ops.direct_grads(sinks):
            (ya * wgt).sum().backward()
        for (n_, p), q in zip(a.named_parameters(), b.parameters()):
            assert p.grad is None, n_
            got = bufs[n_] - base[n_]
            scale = float(q.grad.abs().max()) + 1e-30
            assert float((got - q.grad).abs().max()) <= 1e-6 * scale + 1e-12, (n_, float((got - q.grad).abs().max()), scale)
    else:
        (ya * wgt).sum().backward()
        for (n_, p), q in zip(a.named_parameters(), b.parameters()):
            assert p.grad is not None, n_
            scale = float(q.grad.abs().max()) + 1e-30
            assert float((p.grad - q.grad).abs().max()) <= 1e-6 * scale + 1e-12, (n_, float((p.grad - q.grad).abs().max()), scale)


def test_embedder_plugin_uses_the_schedule_and_matches_reference_layout(emulated, monkeypatch):
    """Embedder.get_identity_embedding through the native schedule: (B, K) frames -> per-frame embeddings -> mean."""
    from embedders import unsupervised_pose_separate_embResNeXt_segmentation as plug
    torch.manual_seed(3)
    emb = plug.Embedder(32, 16, "sum").double()
    ref = copy.deepcopy(emb)
    monkeypatch.setattr(plug.Embedder, "_native_identity_path", lambda self, x: True)
    d = {"enc_rgbs": torch.rand(2, 2, 3, 64, 64, dtype=torch.float64)}
    emb.get_identity_embedding(d)
    with torch.no_grad():
        per_frame = ref.identity_encoder(d["enc_rgbs"].reshape(-1, 3, 64, 64)).view(2, 2, -1)
    torch.testing.assert_close(d["embeds_elemwise"], per_frame, rtol=1e-8, atol=1e-10)
    torch.testing.assert_close(d["embeds"], per_frame.mean(1), rtol=1e-8, atol=1e-10)
