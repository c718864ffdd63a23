"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference modules
(imported from /root/reference; nothing is copied) on the deterministic weights/inputs of oracle/synth.py.

    python oracle/make_golden.py            # needs /root/reference (build container only)

This process puts the reference tree first on sys.path (its top-level package names `generators`, `criterions`, ...
collide with this repo's plugin tree), so it must never import the product packages.
Shims: a `yamlenv` module, fabricated VGG weight files in the layout
criterions/common/perceptual_loss.py:33-41,55 expects.
"""
import os
import sys
import tempfile
import types
from argparse import Namespace
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("LPR_REFERENCE", "/root/reference"))


def setup_reference_imports():
    if not (REF / "generators").is_dir():
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, str(REF))
    sys.path.insert(1, str(REPO))          # for `oracle.*` only
    import yaml
    shim = types.ModuleType("yamlenv")
    shim.load = yaml.safe_load
    sys.modules.setdefault("yamlenv", shim)


def fabricate_vgg_files(vgg19_sd, vggface_sd, dirname):
    """Write vgg19-d01eb7cb.pth / vgg_face_weights.pth with our synthetic conv weights in the reference's layout."""
    import torchvision
    torch.manual_seed(0)
    m = torchvision.models.vgg19()
    full = m.state_dict()
    for k, v in vgg19_sd.items():
        full["features." + k] = v.clone()
    # loader renames classifier.6 -> classifier.7 and prepends Flatten (so saved keys are classifier.{1,4,6})
    ren = {"classifier.0": "classifier.1", "classifier.3": "classifier.4"}
    out = {}
    for k, v in full.items():
        for a, b in ren.items():
            if k.startswith(a + "."):
                k = b + k[len(a):]
        out[k] = v
    torch.save(out, os.path.join(dirname, "vgg19-d01eb7cb.pth"))
    torch.manual_seed(0)
    f = torchvision.models.vgg16().features.state_dict()
    for k, v in vggface_sd.items():
        f[k] = v.clone()
    torch.save(f, os.path.join(dirname, "vgg_face_weights.pth"))


def make_args(cfg, vgg_dir, finetune=False):
    return Namespace(
        gen_padding="zero", in_channels=3, out_channels=3, num_channels=cfg["num_channels"],
        max_num_channels=cfg["max_num_channels"], embed_channels=cfg["embed_channels"],
        pose_embedding_size=cfg["pose_embedding_size"], norm_layer="in", gen_constant_input_size=4,
        gen_num_residual_blocks=2, image_size=cfg["image_size"], device="cpu", average_function="sum",
        dis_padding="zero", dis_num_blocks=cfg["dis_num_blocks"], num_labels=cfg["num_labels"],
        gan_type=cfg["gan_type"], fm_weight=cfg["fm_weight"], dice_weight=cfg["dice_weight"],
        perc_weight=cfg["perc_weight"], idt_embed_weight=cfg["idt_embed_weight"],
        dis_embed_weight=cfg["dis_embed_weight"], vgg_weights_dir=vgg_dir, optimizer="Adam", lr_gen=5e-5,
        lr_dis=2e-4, beta1=0.0, finetune=finetune, num_gpus=1)


class StubEmbedder(torch.nn.Module):
    """Stands in for the (stock torchvision) embedder plugin: writes precomputed embeddings into data_dict, through
    a trainable scale so that the runner's optimizer has an embedder parameter and gradients reach it."""

    def __init__(self, emb):
        super().__init__()
        self.emb = emb
        self.scale = torch.nn.Parameter(torch.ones(()))
        self.finetuning = False

    def forward(self, data_dict):
        data_dict["embeds"] = self.emb["embeds"] * self.scale
        data_dict["embeds_elemwise"] = self.emb["embeds_elemwise"] * self.scale
        data_dict["pose_embedding"] = self.emb["pose_embedding"] * self.scale


def main():
    setup_reference_imports()
    import importlib
    from oracle import synth

    out_dir = REPO / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)

    with tempfile.TemporaryDirectory() as vgg_dir:
        vgg19_sd = synth.vgg_state_dict("vgg19", seed=3)
        vggface_sd = synth.vgg_state_dict("vgg16", seed=5)
        fabricate_vgg_files(vgg19_sd, vggface_sd, vgg_dir)

        # ------------------------------------------------------------------ small configuration, batch 2
        cfg = synth.SMALL_CFG
        args = make_args(cfg, vgg_dir)
        G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
        D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
        crit = {n: importlib.import_module(f"criterions.{n}").Wrapper.get_net(args)
                for n in ("perceptual", "idt_embed", "adversarial", "featmat", "dice", "dis_embed")}
        g_sd = synth.generator_state_dict(cfg, seed=1)
        d_sd = synth.discriminator_state_dict(cfg, seed=2)
        G.load_state_dict(g_sd, strict=True)
        D.load_state_dict(d_sd, strict=True)
        data, target, emb = synth.make_inputs(cfg, batch=2, seed=4)
        gold = {"cfg": dict(cfg), "g_param_order": [k for k, _ in G.named_parameters()],
                "d_param_order": [k for k, _ in D.named_parameters()],
                "g_state_keys": list(G.state_dict().keys()), "d_state_keys": list(D.state_dict().keys()),
                "g_state_shapes": {k: tuple(v.shape) for k, v in G.state_dict().items()},
                "d_state_shapes": {k: tuple(v.shape) for k, v in D.state_dict().items()}}

        with torch.no_grad():
            G.eval()
            dd = dict(embeds=emb["embeds"], pose_embedding=emb["pose_embedding"])
            G(dd)
            gold["g_eval.fake_rgbs"] = dd["fake_rgbs"].clone()
            gold["g_eval.fake_segm"] = dd["fake_segm"].clone()
            G.train()
            dd = dict(embeds=emb["embeds"], pose_embedding=emb["pose_embedding"])
            G(dd)
            gold["g_train.fake_rgbs"] = dd["fake_rgbs"].clone()
            gold["g_train.u_after.decoder_blocks.0.block.3"] = G.state_dict()["decoder_blocks.0.block.3.weight_u"].clone()
            gold["g_train.v_after.affine_params_projector.2"] = G.state_dict()["affine_params_projector.2.weight_v"].clone()
            G.load_state_dict(g_sd, strict=True)

            # discriminator, train mode (three passes, three power iterations)
            D.train()
            fake = gold["g_eval.fake_rgbs"]
            dd = dict(fake_rgbs=fake, target_rgbs=data["target_rgbs"], label=target["label"])
            D(dd)
            for k in ("fake_score_G", "fake_score_D", "real_score", "real_embedding"):
                gold["d_train." + k] = dd[k].clone()
            for i, f in enumerate(dd["fake_features"]):
                gold[f"d_train.fake_features.{i}"] = f.clone()
            for i, f in enumerate(dd["real_features"]):
                gold[f"d_train.real_features.{i}"] = f.clone()
            gold["d_train.u_after.blocks.0.block.2"] = D.state_dict()["blocks.0.block.2.weight_u"].clone()
            D.load_state_dict(d_sd, strict=True)
            D.eval()
            dd = dict(fake_rgbs=fake, target_rgbs=data["target_rgbs"], label=target["label"])
            D(dd)
            for k in ("fake_score_G", "fake_score_D", "real_score"):
                gold["d_eval." + k] = dd[k].clone()
            # criteria on the eval-mode discriminator outputs
            dd.update(fake_segm=gold["g_eval.fake_segm"], real_segm=target["real_segm"],
                      embeds_elemwise=emb["embeds_elemwise"])
            gold["crit.VGG"] = crit["perceptual"](dd)["VGG"].clone()
            gold["crit.VGGFace"] = crit["idt_embed"](dd)["VGGFace"].clone()
            lg, ld = crit["adversarial"](dd)
            gold["crit.adversarial_G"] = lg["adversarial_G"].clone()
            gold["crit.adversarial_D"] = ld["adversarial_D"].clone()
            gold["crit.feature_matching"] = crit["featmat"](dd)["feature_matching"].clone()
            gold["crit.segmentation_dice"] = crit["dice"](dd)["segmentation_dice"].clone()
            gold["crit.embedding_matching"] = crit["dis_embed"](dd)["embedding_matching"].clone()

        # ------------------------------------------------------------------ one full runner step (train mode)
        runner = importlib.import_module("runners.holycow")
        G.load_state_dict(g_sd, strict=True)
        D.load_state_dict(d_sd, strict=True)
        E = StubEmbedder(emb)
        crit_list = [crit[n] for n in ("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice")]
        tm = runner.TrainingModule(E, G, D, crit_list, [], {})
        tm.train()
        opt_G = runner.get_optimizer(E, G, args)
        opt_D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_optimizer(D, args)
        all_dd, lG, lD = tm(dict(data), dict(target))
        loss_G = sum(lG.values())
        loss_D = sum(lD.values())
        for k, v in list(lG.items()) + list(lD.items()):
            gold["step.loss." + k] = v.detach().clone()
        opt_G.zero_grad()
        loss_G.backward(retain_graph=True)
        gold["step.gradG.norms"] = {k: p.grad.norm().item() for k, p in G.named_parameters()}
        for k in ("constant.constant", "decoder_blocks.0.block.3.weight_orig", "decoder_blocks.3.block.4.weight_orig",
                  "decoder_blocks.4.skip.1.weight_orig", "decoder_blocks.4.skip.1.bias", "decoder_blocks.7.weight_orig",
                  "decoder_blocks.7.bias", "affine_params_projector.0.weight_orig", "affine_params_projector.2.bias"):
            gold["step.gradG." + k] = dict(G.named_parameters())[k].grad.clone()
        gold["step.gradE.scale"] = E.scale.grad.clone()
        opt_G.step()
        opt_D.zero_grad()
        loss_D.backward()
        gold["step.gradD.norms"] = {k: p.grad.norm().item() for k, p in D.named_parameters()}
        for k in ("down_block.0.weight_orig", "down_block.0.bias", "down_block.2.weight_orig", "skip.0.weight_orig",
                  "blocks.0.block.2.weight_orig", "blocks.0.block.5.bias", "blocks.1.skip.0.weight_orig",
                  "blocks.5.block.5.weight_orig", "linear.weight_orig", "linear.bias", "embed.weight_orig"):
            gold["step.gradD." + k] = dict(D.named_parameters())[k].grad.clone()
        opt_D.step()
        tm.update_running_average(0.999)
        gold["step.after.G.decoder_blocks.0.block.3.weight_orig"] = \
            dict(G.named_parameters())["decoder_blocks.0.block.3.weight_orig"].detach().clone()
        gold["step.after.D.blocks.0.block.2.weight_orig"] = \
            dict(D.named_parameters())["blocks.0.block.2.weight_orig"].detach().clone()
        gold["step.after.ema.G.decoder_blocks.0.block.3.weight_orig"] = \
            tm.running_averages["generator"].state_dict()["decoder_blocks.0.block.3.weight_orig"].clone()

        # ------------------------------------------------------------------ fine-tuning mode (configs/finetuning-base)
        with torch.no_grad():
            G.load_state_dict(g_sd, strict=True)
            D.load_state_dict(d_sd, strict=True)
            G.enable_finetuning({"embeds": emb["embeds"][:1].clone()})
            D.enable_finetuning({"embeds": emb["embeds"][:1].clone()})
            gold["ft.g_state_keys"] = list(G.state_dict().keys())
            gold["ft.d_state_shapes"] = {k: tuple(v.shape) for k, v in D.state_dict().items() if k.startswith("embed")}
            G.eval()
            dd = dict(pose_embedding=emb["pose_embedding"])
            G(dd)
            gold["ft.g_eval.fake_rgbs"] = dd["fake_rgbs"].clone()
            gold["ft.d_embed_u"] = D.state_dict()["embed.weight_u"].clone()
            gold["ft.d_embed_v"] = D.state_dict()["embed.weight_v"].clone()
            D.eval()
            dd = dict(fake_rgbs=dd["fake_rgbs"], target_rgbs=data["target_rgbs"], label=torch.zeros(2, dtype=torch.long))
            D(dd)
            gold["ft.d_eval.real_score"] = dd["real_score"].clone()
            gold["ft.d_eval.fake_score_G"] = dd["fake_score_G"].clone()

        torch.save(gold, out_dir / "small.pt")
        print("wrote", out_dir / "small.pt", f"{(out_dir / 'small.pt').stat().st_size / 1e6:.2f} MB")

        # ------------------------------------------------------------------ full-size generator, batch 1, eval
        cfg = synth.FULL_CFG
        args = make_args(cfg, vgg_dir)
        with torch.no_grad():
            G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
            g_sd = synth.generator_state_dict(cfg, seed=11)
            G.load_state_dict(g_sd, strict=True)
            G.eval()
            _, _, emb = synth.make_inputs(cfg, batch=1, seed=14)
            dd = dict(embeds=emb["embeds"], pose_embedding=emb["pose_embedding"])
            G(dd)
            full = {"cfg": dict(cfg), "g_eval.fake_rgbs.sub4": dd["fake_rgbs"][:, :, ::4, ::4].clone(),
                    "g_eval.fake_segm.sub4": dd["fake_segm"][:, :, ::4, ::4].clone(),
                    "g_eval.fake_rgbs.mean": dd["fake_rgbs"].mean().clone(),
                    "g_eval.fake_rgbs.std": dd["fake_rgbs"].std().clone()}
            # fp64 ground truth for the same weights (error floor of the fp32 reference itself)
            G64 = G.double()
            dd64 = dict(embeds=emb["embeds"].double(), pose_embedding=emb["pose_embedding"].double())
            G64(dd64)
            full["g_eval.fake_rgbs.sub4.fp64"] = dd64["fake_rgbs"][:, :, ::4, ::4].clone()
            full["g_eval.fp32_vs_fp64_maxabs"] = (dd["fake_rgbs"].double() - dd64["fake_rgbs"]).abs().max().clone()
        torch.save(full, out_dir / "full.pt")
        print("wrote", out_dir / "full.pt", f"{(out_dir / 'full.pt').stat().st_size / 1e6:.2f} MB")
        print("fp32 vs fp64 reference max-abs on fake_rgbs:", float(full["g_eval.fp32_vs_fp64_maxabs"]))


if __name__ == "__main__":
    main()
