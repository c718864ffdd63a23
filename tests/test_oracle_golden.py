"""The oracle restatement (oracle/reference_model.py) against golden vectors produced by the UNMODIFIED reference
modules (oracle/make_golden.py).  CPU only; this is what pins the oracle."""
import copy

import pytest
import torch

from oracle import reference_model as R
from oracle import synth

TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.fixture(scope="module")
def small():
    cfg = synth.SMALL_CFG
    return dict(cfg=cfg, g_sd=synth.generator_state_dict(cfg, seed=1), d_sd=synth.discriminator_state_dict(cfg, seed=2),
                vgg=synth.vgg_state_dict("vgg19", seed=3), vggface=synth.vgg_state_dict("vgg16", seed=5),
                inputs=synth.make_inputs(cfg, batch=2, seed=4),
                g_layout=R.generator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["image_size"]),
                d_layout=R.discriminator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["embed_channels"],
                                                cfg["dis_num_blocks"], cfg["image_size"]))


def test_state_dict_layout_matches_reference(small, golden_small):
    assert set(small["g_sd"].keys()) == set(golden_small["g_state_keys"])
    assert set(small["d_sd"].keys()) == set(golden_small["d_state_keys"])
    for k, shape in golden_small["g_state_shapes"].items():
        assert tuple(small["g_sd"][k].shape) == shape, k
    for k, shape in golden_small["d_state_shapes"].items():
        assert tuple(small["d_sd"][k].shape) == shape, k


def test_generator_eval_and_train(small, golden_small):
    data, target, emb = small["inputs"]
    with torch.no_grad():
        sd = copy.deepcopy(small["g_sd"])
        rgb, segm, _ = R.generator_forward(sd, emb["embeds"], emb["pose_embedding"], small["g_layout"], training=False)
        torch.testing.assert_close(rgb, golden_small["g_eval.fake_rgbs"], **TOL)
        torch.testing.assert_close(segm, golden_small["g_eval.fake_segm"], **TOL)
        rgb, _, _ = R.generator_forward(sd, emb["embeds"], emb["pose_embedding"], small["g_layout"], training=True)
        torch.testing.assert_close(rgb, golden_small["g_train.fake_rgbs"], **TOL)
        torch.testing.assert_close(sd["decoder_blocks.0.block.3.weight_u"],
                                   golden_small["g_train.u_after.decoder_blocks.0.block.3"], **TOL)
        torch.testing.assert_close(sd["affine_params_projector.2.weight_v"],
                                   golden_small["g_train.v_after.affine_params_projector.2"], **TOL)


def test_discriminator_train_and_eval(small, golden_small):
    data, target, emb = small["inputs"]
    fake = golden_small["g_eval.fake_rgbs"]
    with torch.no_grad():
        sd = copy.deepcopy(small["d_sd"])
        out = R.discriminator_forward(sd, fake, data["target_rgbs"][:, 0], target["label"], small["d_layout"], training=True)
        for k in ("fake_score_G", "fake_score_D", "real_score", "real_embedding"):
            torch.testing.assert_close(out[k], golden_small["d_train." + k], **TOL)
        for i in range(7):
            torch.testing.assert_close(out["fake_features"][i], golden_small[f"d_train.fake_features.{i}"], **TOL)
            torch.testing.assert_close(out["real_features"][i], golden_small[f"d_train.real_features.{i}"], **TOL)
        torch.testing.assert_close(sd["blocks.0.block.2.weight_u"], golden_small["d_train.u_after.blocks.0.block.2"], **TOL)
        # the in-place-ReLU aliasing: features 0..5 are post-ReLU, the last one is not
        assert all(float(out["fake_features"][i].min()) >= 0 for i in range(6))
        sd = copy.deepcopy(small["d_sd"])
        out = R.discriminator_forward(sd, fake, data["target_rgbs"][:, 0], target["label"], small["d_layout"], training=False)
        for k in ("fake_score_G", "fake_score_D", "real_score"):
            torch.testing.assert_close(out[k], golden_small["d_eval." + k], **TOL)


def test_criteria(small, golden_small):
    data, target, emb = small["inputs"]
    cfg = small["cfg"]
    fake, segm = golden_small["g_eval.fake_rgbs"], golden_small["g_eval.fake_segm"]
    tgt = data["target_rgbs"][:, 0]
    with torch.no_grad():
        sd = copy.deepcopy(small["d_sd"])
        d = R.discriminator_forward(sd, fake, tgt, target["label"], small["d_layout"], training=False)
        tol = dict(rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(R.perceptual_loss(small["vgg"], fake, tgt, cfg["perc_weight"]), golden_small["crit.VGG"], **tol)
        torch.testing.assert_close(R.idt_embed_loss(small["vggface"], fake, tgt, cfg["idt_embed_weight"]),
                                   golden_small["crit.VGGFace"], **tol)
        lg, ld = R.adversarial_losses(d["fake_score_G"], d["fake_score_D"], d["real_score"])
        torch.testing.assert_close(lg, golden_small["crit.adversarial_G"], **tol)
        torch.testing.assert_close(ld, golden_small["crit.adversarial_D"], **tol)
        torch.testing.assert_close(R.featmat_loss(d["fake_features"], d["real_features"], cfg["fm_weight"]),
                                   golden_small["crit.feature_matching"], **tol)
        torch.testing.assert_close(R.dice_loss(segm, target["real_segm"][:, 0], cfg["dice_weight"]),
                                   golden_small["crit.segmentation_dice"], **tol)
        torch.testing.assert_close(R.dis_embed_loss(emb["embeds_elemwise"], d["real_embedding"], cfg["dis_embed_weight"]),
                                   golden_small["crit.embedding_matching"], **tol)


def test_training_step_gradients(small, golden_small):
    """forward_losses + autograd reproduces the reference runner's losses and gradients (runners/holycow.py:230-252)."""
    data, target, emb = small["inputs"]
    cfg = small["cfg"]
    g_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") or k == "constant.constant"
                else v.clone()) for k, v in small["g_sd"].items()}
    d_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") else v.clone())
            for k, v in small["d_sd"].items()}
    out, lg, ld = R.forward_losses(g_sd, d_sd, small["vgg"], small["vggface"], cfg, emb["embeds"], emb["pose_embedding"],
                                   data["target_rgbs"][:, 0], target["real_segm"][:, 0], target["label"], training=True,
                                   embeds_elemwise=emb["embeds_elemwise"],
                                   criteria=("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice"))
    tol = dict(rtol=2e-4, atol=1e-6)
    for k, v in {**lg, **ld}.items():
        torch.testing.assert_close(v.detach(), golden_small["step.loss." + k], **tol)
    g_params = {k: v for k, v in g_sd.items() if v.requires_grad}
    grads = torch.autograd.grad(sum(lg.values()), list(g_params.values()), retain_graph=True, allow_unused=True)
    for (k, _), g in zip(g_params.items(), grads):
        ref_norm = golden_small["step.gradG.norms"][k]
        assert abs(float(g.norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-7, k
        if "step.gradG." + k in golden_small:
            torch.testing.assert_close(g, golden_small["step.gradG." + k], rtol=2e-3, atol=1e-6 + 1e-4 * ref_norm)
    d_params = {k: v for k, v in d_sd.items() if v.requires_grad}
    grads = torch.autograd.grad(sum(ld.values()), list(d_params.values()), allow_unused=True)
    for (k, _), g in zip(d_params.items(), grads):
        ref_norm = golden_small["step.gradD.norms"][k]
        gn = 0.0 if g is None else float(g.norm())
        assert abs(gn - ref_norm) <= 2e-3 * ref_norm + 1e-7, k
        if "step.gradD." + k in golden_small and g is not None:
            torch.testing.assert_close(g, golden_small["step.gradD." + k], rtol=2e-3, atol=1e-6 + 1e-4 * ref_norm)


def test_finetune_mode(small, golden_small):
    data, target, emb = small["inputs"]
    with torch.no_grad():
        sd = copy.deepcopy(small["g_sd"])
        ident = emb["embeds"][:1].expand(2, -1)
        rgb, _, _ = R.generator_forward(sd, ident, emb["pose_embedding"], small["g_layout"], training=False)
        torch.testing.assert_close(rgb, golden_small["ft.g_eval.fake_rgbs"], **TOL)
        assert "identity_embedding" in golden_small["ft.g_state_keys"]
        assert golden_small["ft.d_state_shapes"]["embed.weight_orig"] == (1, small["cfg"]["embed_channels"])


def test_full_size_generator(golden_full):
    """Full-size (256x256, default channels) generator, batch 1, eval: restatement vs the reference (fp32 and fp64)."""
    cfg = golden_full["cfg"]
    g_sd = synth.generator_state_dict(cfg, seed=11)
    _, _, emb = synth.make_inputs(cfg, batch=1, seed=14)
    layout = R.generator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["image_size"])
    with torch.no_grad():
        rgb, segm, _ = R.generator_forward(g_sd, emb["embeds"], emb["pose_embedding"], layout, training=False)
    assert (rgb[:, :, ::4, ::4] - golden_full["g_eval.fake_rgbs.sub4"]).abs().max() < 5e-5
    assert (rgb[:, :, ::4, ::4].double() - golden_full["g_eval.fake_rgbs.sub4.fp64"]).abs().max() < 5e-5
    assert (segm[:, :, ::4, ::4] - golden_full["g_eval.fake_segm.sub4"]).abs().max() < 5e-5
    assert float(golden_full["g_eval.fp32_vs_fp64_maxabs"]) < 1e-4


def test_full_size_step_losses_and_gradients():
    """The restatement at the REAL size (256x256, 64..512 channels, batch 2, train mode) against the full runner step of
    the unmodified reference (tests/golden/full_step.pt, oracle/make_golden_full.py): every loss, every generator /
    discriminator gradient norm, the sub-sampled gradient tensors."""
    from conftest import GOLDEN
    gold = torch.load(GOLDEN / "full_step.pt", map_location="cpu", weights_only=False)
    cfg = gold["cfg"]
    g_sd0, d_sd0 = synth.generator_state_dict(cfg, seed=21), synth.discriminator_state_dict(cfg, seed=22)
    g_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") or k == "constant.constant"
                else v.clone()) for k, v in g_sd0.items()}
    d_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") else v.clone())
            for k, v in d_sd0.items()}
    data, target, emb = synth.make_inputs(cfg, batch=2, seed=24)
    out, lg, ld = R.forward_losses(g_sd, d_sd, synth.vgg_state_dict("vgg19", seed=3), synth.vgg_state_dict("vgg16", seed=5),
                                   cfg, emb["embeds"], emb["pose_embedding"], data["target_rgbs"][:, 0],
                                   target["real_segm"][:, 0], target["label"], training=True,
                                   embeds_elemwise=emb["embeds_elemwise"],
                                   criteria=("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice"))
    for k, v in {**lg, **ld}.items():
        torch.testing.assert_close(v.detach(), gold["step.loss." + k], rtol=3e-4, atol=1e-6)

    def sub(t):
        if t.dim() == 4:
            return t[::max(1, t.shape[0] // 16), ::max(1, t.shape[1] // 16), ::max(1, t.shape[2] // 16), ::max(1, t.shape[2] // 16)]
        if t.dim() == 2:
            return t[::max(1, t.shape[0] // 64), ::max(1, t.shape[1] // 64)]
        return t
    g_params = {k: v for k, v in g_sd.items() if v.requires_grad}
    grads = torch.autograd.grad(sum(lg.values()), list(g_params.values()), retain_graph=True, allow_unused=True)
    g_max = max(gold["step.gradG.norms"].values())
    for (k, _), g in zip(g_params.items(), grads):
        ref_norm = gold["step.gradG.norms"][k]
        assert abs(float(g.norm()) - ref_norm) <= 3e-3 * ref_norm + 1e-5 * g_max, k
        ref = gold["step.gradG.sub." + k]
        assert float((sub(g) - ref).abs().max()) <= 3e-3 * float(ref.abs().max()) + 1e-5 * g_max, k
    d_params = {k: v for k, v in d_sd.items() if v.requires_grad}
    grads = torch.autograd.grad(sum(ld.values()), list(d_params.values()), allow_unused=True)
    d_max = max(gold["step.gradD.norms"].values())
    for (k, _), g in zip(d_params.items(), grads):
        ref_norm = gold["step.gradD.norms"][k]
        gn = 0.0 if g is None else float(g.norm())
        assert abs(gn - ref_norm) <= 3e-3 * ref_norm + 1e-5 * d_max, k


def test_512_step_losses_and_gradient_norms():
    """The restatement at BASELINE configs[4]'s shapes (512x512: 19 AdaIN sites, 8 up-blocks; batch 1, train mode) against
    the full runner step of the unmodified reference (tests/golden/step512.pt, oracle/make_golden_step512.py): every loss,
    every generator / discriminator gradient norm, the stored sub-sampled gradient tensors."""
    from conftest import GOLDEN
    gold = torch.load(GOLDEN / "step512.pt", map_location="cpu", weights_only=False)
    cfg = gold["cfg"]
    assert cfg["image_size"] == 512
    g_sd0, d_sd0 = synth.generator_state_dict(cfg, seed=31), synth.discriminator_state_dict(cfg, seed=32)
    g_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") or k == "constant.constant"
                else v.clone()) for k, v in g_sd0.items()}
    d_sd = {k: (v.clone().requires_grad_(True) if "weight_orig" in k or k.endswith("bias") else v.clone())
            for k, v in d_sd0.items()}
    data, target, emb = synth.make_inputs(cfg, batch=1, seed=34)
    out, lg, ld = R.forward_losses(g_sd, d_sd, synth.vgg_state_dict("vgg19", seed=3), synth.vgg_state_dict("vgg16", seed=5),
                                   cfg, emb["embeds"], emb["pose_embedding"], data["target_rgbs"][:, 0],
                                   target["real_segm"][:, 0], target["label"], training=True,
                                   embeds_elemwise=emb["embeds_elemwise"],
                                   criteria=("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice"))
    for k, v in {**lg, **ld}.items():
        torch.testing.assert_close(v.detach(), gold["step.loss." + k], rtol=3e-4, atol=1e-6)

    def sub(t):
        if t.dim() == 4:
            return t[::max(1, t.shape[0] // 16), ::max(1, t.shape[1] // 16), ::max(1, t.shape[2] // 16), ::max(1, t.shape[2] // 16)]
        if t.dim() == 2:
            return t[::max(1, t.shape[0] // 64), ::max(1, t.shape[1] // 64)]
        return t
    for sd, losses, key in ((g_sd, lg, "G"), (d_sd, ld, "D")):
        params = {k: v for k, v in sd.items() if v.requires_grad}
        grads = torch.autograd.grad(sum(losses.values()), list(params.values()), retain_graph=(key == "G"), allow_unused=True)
        ref_norms = gold[f"step.grad{key}.norms"]
        top = max(ref_norms.values())
        assert set(params) == set(ref_norms), set(params) ^ set(ref_norms)
        for (k, _), g in zip(params.items(), grads):
            gn = 0.0 if g is None else float(g.norm())
            assert abs(gn - ref_norms[k]) <= 3e-3 * ref_norms[k] + 1e-5 * top, (k, gn, ref_norms[k])
            ref = gold.get(f"step.grad{key}.sub." + k)
            if ref is not None:
                assert float((sub(g) - ref).abs().max()) <= 5e-3 * float(ref.abs().max()) + 1e-5 * top, k
