"""ORACLE — TEST INFRASTRUCTURE ONLY.  Full-size golden vectors from the UNMODIFIED reference modules (imported from
/root/reference; nothing is copied), on the deterministic weights / inputs of oracle/synth.py:

    python oracle/make_golden_full.py       # needs /root/reference (build container only); ~3 min of CPU

  tests/golden/full_step.pt  FULL_CFG (256x256, 64..512 channels, 37.5 M + 19.5 M parameters), batch 2, train mode:
                             discriminator three-pass scores + the 14 feature maps (sub-sampled), every criterion value,
                             one full runners.holycow step — losses, gradient norms of ALL generator / discriminator
                             parameters, sub-sampled gradient tensors, post-Adam / post-EMA weights
  tests/golden/full512.pt    image_size 512 (BASELINE configs[4] shapes: 19 AdaIN sites, 8 up-blocks), batch 1:
                             generator eval forward, discriminator train-mode pass
  tests/golden/identity.pt   the reference Embedder's identity path (ResNeXt50-32x4d, train-mode BatchNorm over
                             B*K = 8 frames at 128x128): embeddings, running statistics, parameter-gradient norms

Tensors are stored sub-sampled (strided slices; the slicing is part of the fixture and repeated by the tests) so the
fixtures stay small.
"""
import importlib
import os
import sys
import tempfile
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from oracle.make_golden import StubEmbedder, fabricate_vgg_files, make_args, setup_reference_imports  # noqa: E402


def sub(t):
    """Strided sub-sample used for every stored activation / gradient tensor: <= ~16 K values."""
    if t.dim() == 4:        # activations (N, C, H, W) and conv weights (O, I, k, k)
        s0 = max(1, t.shape[0] // 16)
        cs = max(1, t.shape[1] // 16)
        ss = max(1, t.shape[2] // 16)
        return t[::s0, ::cs, ::ss, ::ss].clone()
    if t.dim() == 2:
        return t[::max(1, t.shape[0] // 64), ::max(1, t.shape[1] // 64)].clone()
    return t.clone()


def main():
    setup_reference_imports()
    from oracle import synth
    out_dir = REPO / "tests" / "golden"
    torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as vgg_dir:
        fabricate_vgg_files(synth.vgg_state_dict("vgg19", seed=3), synth.vgg_state_dict("vgg16", seed=5), vgg_dir)

        # ------------------------------------------------------------------ FULL_CFG, batch 2
        cfg = synth.FULL_CFG
        args = make_args(cfg, vgg_dir)
        G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
        D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
        crit = {n: importlib.import_module(f"criterions.{n}").Wrapper.get_net(args)
                for n in ("perceptual", "idt_embed", "adversarial", "featmat", "dice", "dis_embed")}
        g_sd = synth.generator_state_dict(cfg, seed=21)
        d_sd = synth.discriminator_state_dict(cfg, seed=22)
        G.load_state_dict(g_sd, strict=True)
        D.load_state_dict(d_sd, strict=True)
        data, target, emb = synth.make_inputs(cfg, batch=2, seed=24)
        gold = {"cfg": dict(cfg)}
        with torch.no_grad():
            G.eval()
            dd = dict(embeds=emb["embeds"], pose_embedding=emb["pose_embedding"])
            G(dd)
            fake, segm = dd["fake_rgbs"].clone(), dd["fake_segm"].clone()
            gold["g_eval.fake_rgbs.sub"] = sub(fake)
            gold["g_eval.fake_rgbs"] = fake.clone()          # 1.5 MB: the discriminator / criteria tests start from it
            gold["g_eval.fake_segm"] = segm.clone()
            D.train()
            dd = dict(fake_rgbs=fake, target_rgbs=data["target_rgbs"], label=target["label"])
            D(dd)
            for k in ("fake_score_G", "fake_score_D", "real_score", "real_embedding"):
                gold["d_train." + k] = dd[k].clone()
            for i, f in enumerate(dd["fake_features"]):
                gold[f"d_train.fake_features.{i}.sub"] = sub(f)
                gold[f"d_train.fake_features.{i}.absmean"] = f.abs().mean().clone()
            for i, f in enumerate(dd["real_features"]):
                gold[f"d_train.real_features.{i}.sub"] = sub(f)
                gold[f"d_train.real_features.{i}.absmean"] = f.abs().mean().clone()
            gold["d_train.u_after.blocks.0.block.2"] = D.state_dict()["blocks.0.block.2.weight_u"].clone()
            dd.update(fake_segm=segm, real_segm=target["real_segm"], embeds_elemwise=emb["embeds_elemwise"])
            gold["crit.VGG"] = crit["perceptual"](dd)["VGG"].clone()
            gold["crit.VGGFace"] = crit["idt_embed"](dd)["VGGFace"].clone()
            lg, ld = crit["adversarial"](dd)
            gold["crit.adversarial_G"], gold["crit.adversarial_D"] = lg["adversarial_G"].clone(), ld["adversarial_D"].clone()
            gold["crit.feature_matching"] = crit["featmat"](dd)["feature_matching"].clone()
            gold["crit.segmentation_dice"] = crit["dice"](dd)["segmentation_dice"].clone()
            gold["crit.embedding_matching"] = crit["dis_embed"](dd)["embedding_matching"].clone()
        print("forward pieces done")

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
        loss_G, loss_D = sum(lG.values()), sum(lD.values())
        for k, v in list(lG.items()) + list(lD.items()):
            gold["step.loss." + k] = v.detach().clone()
        gold["step.fake_rgbs.sub"] = sub(all_dd["fake_rgbs"].detach())
        opt_G.zero_grad()
        loss_G.backward(retain_graph=True)
        gold["step.gradG.norms"] = {k: p.grad.norm().item() for k, p in G.named_parameters()}
        for k, p in G.named_parameters():
            gold["step.gradG.sub." + k] = sub(p.grad)
        gold["step.gradE.scale"] = E.scale.grad.clone()
        opt_G.step()
        opt_D.zero_grad()
        loss_D.backward()
        gold["step.gradD.norms"] = {k: p.grad.norm().item() for k, p in D.named_parameters()}
        for k, p in D.named_parameters():
            gold["step.gradD.sub." + k] = sub(p.grad)
        opt_D.step()
        tm.update_running_average(0.999)
        for k in ("decoder_blocks.5.block.4.weight_orig", "decoder_blocks.7.block.8.weight_orig"):
            gold["step.after.G.sub." + k] = sub(dict(G.named_parameters())[k].detach())
            gold["step.after.ema.G.sub." + k] = sub(tm.running_averages["generator"].state_dict()[k])
        gold["step.after.D.sub.blocks.0.block.2.weight_orig"] = sub(dict(D.named_parameters())["blocks.0.block.2.weight_orig"].detach())
        torch.save(gold, out_dir / "full_step.pt")
        print("wrote", out_dir / "full_step.pt", f"{(out_dir / 'full_step.pt').stat().st_size / 1e6:.2f} MB")
        del tm, G, D, E, opt_G, opt_D, all_dd, loss_G, loss_D

        # ------------------------------------------------------------------ 512 x 512 (BASELINE configs[4] shapes)
        cfg5 = dict(synth.FULL_CFG, image_size=512)
        args = make_args(cfg5, vgg_dir)
        with torch.no_grad():
            G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
            D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
            g_sd = synth.generator_state_dict(cfg5, seed=31)
            d_sd = synth.discriminator_state_dict(cfg5, seed=32)
            G.load_state_dict(g_sd, strict=True)
            D.load_state_dict(d_sd, strict=True)
            data, target, emb = synth.make_inputs(cfg5, batch=1, seed=34)
            G.eval()
            dd = dict(embeds=emb["embeds"], pose_embedding=emb["pose_embedding"])
            G(dd)
            g5 = {"cfg": dict(cfg5), "n_adain": len(G.adains), "g_eval.fake_rgbs.sub8": dd["fake_rgbs"][:, :, ::8, ::8].clone(),
                  "g_eval.fake_segm.sub8": dd["fake_segm"][:, :, ::8, ::8].clone(),
                  "g_eval.fake_rgbs.mean": dd["fake_rgbs"].mean().clone()}
            fake = dd["fake_rgbs"].clone()
            D.train()
            dd = dict(fake_rgbs=fake, target_rgbs=data["target_rgbs"], label=target["label"])
            D(dd)
            for k in ("fake_score_G", "fake_score_D", "real_score"):
                g5["d_train." + k] = dd[k].clone()
            g5["d_train.n_features"] = len(dd["fake_features"])
            for i, f in enumerate(dd["fake_features"]):
                g5[f"d_train.fake_features.{i}.sub"] = sub(f)
                g5[f"d_train.fake_features.{i}.shape"] = tuple(f.shape)
        torch.save(g5, out_dir / "full512.pt")
        print("wrote", out_dir / "full512.pt", f"{(out_dir / 'full512.pt').stat().st_size / 1e6:.2f} MB")

    # ---------------------------------------------------------------------- identity encoder (reference Embedder)
    emb_mod = importlib.import_module("embedders.unsupervised_pose_separate_embResNeXt_segmentation")
    E = emb_mod.Embedder(512, 256, "sum")
    sd = synth.identity_encoder_state_dict(512, seed=9)
    E.identity_encoder.load_state_dict(sd, strict=True)
    x = synth.identity_inputs(batch=2, frames=4, image_size=128, seed=10)
    gi = {"num_classes": 512}
    E.train()
    d = {"enc_rgbs": x}
    E.get_identity_embedding(d)
    gi["train.embeds"] = d["embeds"].detach().clone()
    gi["train.embeds_elemwise"] = d["embeds_elemwise"].detach().clone()
    wgt = torch.randn(2, 4, 512, generator=torch.Generator().manual_seed(12))
    (d["embeds_elemwise"] * wgt).sum().backward()
    gi["train.grad_norms"] = {k: p.grad.norm().item() for k, p in E.identity_encoder.named_parameters()}
    for k in ("conv1.weight", "bn1.weight", "layer1.0.conv2.weight", "layer1.0.downsample.0.weight", "layer2.0.conv2.weight",
              "layer3.2.conv1.weight", "layer3.2.bn2.bias", "layer4.0.conv2.weight", "layer4.2.conv3.weight", "fc.weight",
              "fc.bias"):
        gi["train.grad.sub." + k] = sub(dict(E.identity_encoder.named_parameters())[k].grad)
    bns = [m for m in E.identity_encoder.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    gi["train.running_mean"] = torch.cat([m.running_mean for m in bns])[::7].clone()
    gi["train.running_var"] = torch.cat([m.running_var for m in bns])[::7].clone()
    gi["train.num_batches_tracked"] = int(bns[0].num_batches_tracked)
    E.identity_encoder.load_state_dict(sd, strict=True)
    E.eval()
    with torch.no_grad():
        d = {"enc_rgbs": x}
        E.get_identity_embedding(d)
    gi["eval.embeds"] = d["embeds"].clone()
    # float64 ground truth of the train-mode embeddings (error floor of the fp32 reference itself)
    E64 = emb_mod.Embedder(512, 256, "sum").double()
    E64.identity_encoder.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, strict=True)
    E64.train()
    with torch.no_grad():
        d = {"enc_rgbs": x.double()}
        E64.get_identity_embedding(d)
    gi["train.embeds.fp64"] = d["embeds"].clone()
    gi["train.fp32_vs_fp64_rel"] = float((gi["train.embeds"].double() - d["embeds"]).abs().max() / d["embeds"].abs().max())
    torch.save(gi, out_dir / "identity.pt")
    print("wrote", out_dir / "identity.pt", f"{(out_dir / 'identity.pt').stat().st_size / 1e6:.2f} MB",
          "fp32 vs fp64 rel:", gi["train.fp32_vs_fp64_rel"])


if __name__ == "__main__":
    main()
