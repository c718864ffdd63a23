"""Host logic of the generator / discriminator plugins on the CPU: the kernel wrappers are replaced by plain-torch
emulations of their contracts (tests/kernel_emulators.py) and the plugins run in float64, so every difference from the
oracle (oracle/reference_model.py, float64) or from the reference's golden vectors is a scheduling / bookkeeping error —
wrong layout, residual mode, AdaIN slice, spectral-norm state, gradient routing — not arithmetic noise.
The same plugins with the real kernels are compared with the same oracles on the GPU (tests/test_parity_gpu.py)."""
import importlib

import pytest
import torch

import kernel_emulators as E
from helpers import make_args, max_abs, rel_err
from oracle import reference_model as R
from oracle import synth


@pytest.fixture()
def emu(monkeypatch):
    E.install(monkeypatch)
    return E


@pytest.fixture(scope="module")
def small():
    cfg = synth.SMALL_CFG
    data, target, emb = synth.make_inputs(cfg, batch=2, seed=4)
    return dict(cfg=cfg, g_sd=synth.generator_state_dict(cfg, seed=1), d_sd=synth.discriminator_state_dict(cfg, seed=2),
                data=data, target=target, emb=emb)


def _G(cfg, sd):
    args = make_args(cfg, device="cpu")
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
    G.load_state_dict(sd, strict=True)
    return G.double()


def _D(cfg, sd):
    args = make_args(cfg, device="cpu")
    D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
    D.load_state_dict(sd, strict=True)
    return D.double()


def _close(a, b, rtol, floor=1e-4):
    """max |a - b| <= rtol * max(|b|_max, floor): biases in front of an instance norm have analytically zero gradient
    (1e-15 of rounding noise on both sides), hence the floor."""
    return max_abs(a, b) <= rtol * max(float(b.abs().max()), floor)


def _layout_g(cfg):
    return R.generator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["image_size"])


def _layout_d(cfg):
    return R.discriminator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["embed_channels"],
                                  cfg["dis_num_blocks"], cfg["image_size"])


@pytest.mark.parametrize("precision", ["bf16x3", "tf32"])
def test_generator_schedule_forward_and_buffers(emu, small, golden_small, precision):
    cfg, emb = small["cfg"], small["emb"]
    G = _G(cfg, small["g_sd"])
    G.precision = precision                          # both schedules (fused AdaIN+conv nodes / separate nodes)
    sd64 = {k: v.double() for k, v in small["g_sd"].items()}
    with torch.no_grad():
        G.eval()
        dd = dict(embeds=emb["embeds"].double(), pose_embedding=emb["pose_embedding"].double())
        G(dd)
        rgb64, segm64, _ = R.generator_forward(sd64, emb["embeds"].double(), emb["pose_embedding"].double(),
                                               _layout_g(cfg), training=False)
        assert max_abs(dd["fake_rgbs"], rgb64) < 1e-10 and max_abs(dd["fake_segm"], segm64) < 1e-10
        assert max_abs(dd["fake_rgbs"], golden_small["g_eval.fake_rgbs"]) < 2e-5       # the reference itself (fp32)
        G.train()                                     # one power iteration per spectral-normalised weight
        dd = dict(embeds=emb["embeds"].double(), pose_embedding=emb["pose_embedding"].double())
        G(dd)
        assert max_abs(dd["fake_rgbs"], golden_small["g_train.fake_rgbs"]) < 2e-5
        sd = G.state_dict()
        assert rel_err(sd["decoder_blocks.0.block.3.weight_u"], golden_small["g_train.u_after.decoder_blocks.0.block.3"]) < 1e-5
        assert rel_err(sd["affine_params_projector.2.weight_v"], golden_small["g_train.v_after.affine_params_projector.2"]) < 1e-5


@pytest.mark.parametrize("precision", ["bf16x3", "tf32"])
def test_generator_schedule_backward_matches_oracle_autograd(emu, small, precision):
    """Every parameter gradient of the plugin (autograd nodes of b200lp.ops over emulated kernels) against float64
    autograd through the oracle's restatement of the reference generator."""
    cfg, emb = small["cfg"], small["emb"]
    G = _G(cfg, small["g_sd"]).eval()
    G.precision = precision
    torch.manual_seed(0)
    c_rgb, c_segm = torch.randn(2, 3, 32, 32, dtype=torch.float64), torch.randn(2, 1, 32, 32, dtype=torch.float64)
    dd = dict(embeds=emb["embeds"].double(), pose_embedding=emb["pose_embedding"].double())
    G(dd)
    ((dd["fake_rgbs"] * c_rgb).sum() + (dd["fake_segm"] * c_segm).sum()).backward()
    sd64 = {k: v.double().requires_grad_(v.is_floating_point() and ("weight_orig" in k or "bias" in k or "constant" in k))
            for k, v in small["g_sd"].items()}
    rgb, segm, _ = R.generator_forward(sd64, emb["embeds"].double(), emb["pose_embedding"].double(), _layout_g(cfg),
                                       training=False)
    ((rgb * c_rgb).sum() + (segm * c_segm).sum()).backward()
    checked = 0
    for name, p in G.named_parameters():
        ref = sd64[name].grad
        assert ref is not None and p.grad is not None, name
        assert _close(p.grad, ref, 1e-8), name
        checked += 1
    assert checked >= 20


def test_gradient_sinks_equal_autograd_accumulation(emu, small):
    """ops.direct_grads: conv weight / bias gradients accumulated in place by the (emulated) fused kernels equal what
    autograd + AccumulateGrad produce, including the spectral-norm rank-1 term (train mode: batched sigma path)."""
    from b200lp import ops
    cfg, emb = small["cfg"], small["emb"]
    torch.manual_seed(1)
    c_rgb = torch.randn(2, 3, 32, 32, dtype=torch.float64)
    grads = []
    for use_sinks in (False, True):
        G = _G(cfg, small["g_sd"]).train()
        dd = dict(embeds=emb["embeds"].double(), pose_embedding=emb["pose_embedding"].double())
        G(dd)
        loss = (dd["fake_rgbs"] * c_rgb).sum() + dd["fake_segm"].sum()
        params = list(G.parameters())
        for p in params:
            p.grad = torch.zeros_like(p)
        sinks = {p.data_ptr(): p.grad for p in params} if use_sinks else {}
        with ops.direct_grads(sinks):
            loss.backward()
        grads.append({n: p.grad.clone() for n, p in G.named_parameters()})
    for n in grads[0]:
        assert _close(grads[1][n], grads[0][n], 1e-10), n
    assert any("block" in n and float(g.abs().max()) > 0 for n, g in grads[1].items())


def test_discriminator_schedule_three_passes(emu, small, golden_small):
    cfg = small["cfg"]
    D = _D(cfg, small["d_sd"]).train()
    fake = golden_small["g_eval.fake_rgbs"].double()
    target = small["data"]["target_rgbs"].double()
    with torch.no_grad():
        dd = dict(fake_rgbs=fake, target_rgbs=target, label=small["target"]["label"])
        D(dd)
    for k in ("fake_score_G", "fake_score_D", "real_score"):
        assert rel_err(dd[k], golden_small["d_train." + k]) < 2e-5, k
    assert rel_err(dd["real_embedding"], golden_small["d_train.real_embedding"]) < 1e-5
    for i in range(7):
        assert rel_err(dd["fake_features"][i], golden_small[f"d_train.fake_features.{i}"]) < 2e-5, i
        assert rel_err(dd["real_features"][i], golden_small[f"d_train.real_features.{i}"]) < 2e-5, i
    assert all(float(dd["fake_features"][i].min()) >= 0 for i in range(6))       # in-place-ReLU aliasing kept
    assert float(dd["fake_features"][6].min()) < 0
    assert rel_err(D.state_dict()["blocks.0.block.2.weight_u"], golden_small["d_train.u_after.blocks.0.block.2"]) < 1e-5
    # float64 oracle, eval mode (no power iteration): exact agreement
    D2 = _D(cfg, small["d_sd"]).eval()
    d64 = {k: v.double() for k, v in small["d_sd"].items()}
    with torch.no_grad():
        dd = dict(fake_rgbs=fake, target_rgbs=target, label=small["target"]["label"])
        D2(dd)
        o = R.discriminator_forward(d64, fake, target[:, 0], small["target"]["label"], _layout_d(cfg), training=False)
    for k in ("fake_score_G", "fake_score_D", "real_score"):
        assert rel_err(dd[k], o[k]) < 1e-10, k


def test_discriminator_schedule_backward(emu, small):
    """Gradients of the D loss w.r.t. every discriminator parameter and of the G-side loss w.r.t. the fake image
    (the pass that runs on detached weights) against the oracle's autograd."""
    cfg = small["cfg"]
    D = _D(cfg, small["d_sd"]).eval()
    torch.manual_seed(3)
    fake = torch.rand(2, 3, 32, 32, dtype=torch.float64, requires_grad=True)
    target = small["data"]["target_rgbs"].double()
    dd = dict(fake_rgbs=fake, target_rgbs=target, label=small["target"]["label"])
    D(dd)
    loss_d = torch.relu(1 - dd["real_score"]).mean() + torch.relu(1 + dd["fake_score_D"]).mean()
    loss_g = -dd["fake_score_G"].mean() + sum((f - r.detach()).abs().mean()
                                              for f, r in zip(dd["fake_features"], dd["real_features"]))
    (g_fake,) = torch.autograd.grad(loss_g, [fake], retain_graph=True)
    loss_d.backward()
    d64 = {k: v.double().requires_grad_(v.is_floating_point() and ("weight_orig" in k or "bias" in k))
           for k, v in small["d_sd"].items()}
    fake2 = fake.detach().clone().requires_grad_(True)
    o = R.discriminator_forward(d64, fake2, target[:, 0], small["target"]["label"], _layout_d(cfg), training=False)
    ref_d = torch.relu(1 - o["real_score"]).mean() + torch.relu(1 + o["fake_score_D"]).mean()
    ref_g = -o["fake_score_G"].mean() + sum((f - r.detach()).abs().mean()
                                            for f, r in zip(o["fake_features"], o["real_features"]))
    (ref_g_fake,) = torch.autograd.grad(ref_g, [fake2], retain_graph=True)
    ref_d.backward()
    assert rel_err(g_fake, ref_g_fake) < 1e-6          # the Cin=3 data gradient goes through an fp32 transposed weight
    for name, p in D.named_parameters():
        assert p.grad is not None and d64[name].grad is not None, name
        assert _close(p.grad, d64[name].grad, 1e-8), name


def test_tensor_core_tail_node(emu):
    """ops.AdaINTailFn (final AdaIN + padded-weight conv + compose; taken when C % 64 == 0 and the plane is >= 32 x 16)
    against plain autograd: outputs and all six gradients."""
    import torch.nn.functional as F
    from b200lp import ops
    torch.manual_seed(5)
    N, H, C = 2, 32, 64
    x = torch.randn(N, H, H, C, dtype=torch.float64, requires_grad=True)
    aff = torch.randn(N, 2 * C, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(4, C, 3, 3, dtype=torch.float64) * 0.05).requires_grad_(True)
    s = torch.tensor([0.7], dtype=torch.float64, requires_grad=True)
    b = torch.randn(4, dtype=torch.float64, requires_grad=True)
    assert ops.tail_tensor_core_ok(x, w)
    g_r, g_s = torch.randn(N, 3, H, H, dtype=torch.float64), torch.randn(N, 1, H, H, dtype=torch.float64)
    rgbs, segm = ops.adain_tail(x, aff[:, C:], aff[:, :C], w, s, b)
    got = torch.autograd.grad([rgbs, segm], [x, aff, w, s, b], [g_r, g_s])
    xd, ad, wd, sd, bd = [t.detach().clone().requires_grad_(True) for t in (x, aff, w, s, b)]
    a = F.relu(F.instance_norm(xd.permute(0, 3, 1, 2), eps=1e-4) * ad[:, C:][:, :, None, None] + ad[:, :C][:, :, None, None])
    t = torch.tanh(F.conv2d(a, wd * sd, bd, padding=1))
    seg = t[:, 3:] * 0.5 + 0.5
    ref = torch.autograd.grad([(t[:, :3] * 0.75 + 0.5) * seg, seg], [xd, ad, wd, sd, bd], [g_r, g_s])
    assert max_abs(rgbs, (t[:, :3] * 0.75 + 0.5) * seg) < 1e-12
    for name, u, v in zip(["dx", "daffine", "dw", "ds", "db"], got, ref):
        assert rel_err(u, v) < 1e-8, name
