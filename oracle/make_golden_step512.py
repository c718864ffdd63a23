"""ORACLE — TEST INFRASTRUCTURE ONLY.  One full training step of the UNMODIFIED reference at 512x512 (BASELINE.json
configs[4] shapes: 8 up-blocks / 19 AdaIN sites in the generator, 7 discriminator blocks on a 512x512 input, all six
criteria), batch 1, on the deterministic weights / inputs of oracle/synth.py:

    python oracle/make_golden_step512.py      # needs /root/reference (build container only); ~2 min of CPU

  tests/golden/step512.pt   losses, sub-sampled fake image, gradient norms of ALL generator / discriminator parameters,
                            sub-sampled gradient tensors of a handful of layers (first / middle / last of each network,
                            the 512x512-plane layers included), the embedder-scale gradient

The reference modules are imported from /root/reference (nothing is copied); weights and inputs are the ones
tests/golden/full512.pt was made with (seeds 31 / 32 / 34), so the forward test and this step share a configuration.
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
from oracle.make_golden_full import sub  # noqa: E402

# gradient tensors stored (sub-sampled): the layers that only exist / only reach these plane sizes at 512x512
G_SUB = ("decoder_blocks.0.block.3.weight_orig", "decoder_blocks.4.block.8.weight_orig", "decoder_blocks.7.block.4.weight_orig",
         "decoder_blocks.8.block.4.weight_orig", "decoder_blocks.8.block.8.weight_orig", "decoder_blocks.8.skip.1.weight_orig",
         "decoder_blocks.11.weight_orig", "affine_params_projector.2.weight_orig")
D_SUB = ("down_block.0.weight_orig", "down_block.2.weight_orig", "blocks.0.block.2.weight_orig", "blocks.1.block.5.weight_orig",
         "blocks.5.block.5.weight_orig", "linear.weight_orig")


def main():
    setup_reference_imports()
    from oracle import synth
    out_dir = REPO / "tests" / "golden"
    torch.set_num_threads(os.cpu_count() or 1)
    with tempfile.TemporaryDirectory() as vgg_dir:
        fabricate_vgg_files(synth.vgg_state_dict("vgg19", seed=3), synth.vgg_state_dict("vgg16", seed=5), vgg_dir)
        cfg = dict(synth.FULL_CFG, image_size=512)
        args = make_args(cfg, vgg_dir)
        G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
        D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
        crit = {n: importlib.import_module(f"criterions.{n}").Wrapper.get_net(args)
                for n in ("perceptual", "idt_embed", "adversarial", "featmat", "dice", "dis_embed")}
        G.load_state_dict(synth.generator_state_dict(cfg, seed=31), strict=True)
        D.load_state_dict(synth.discriminator_state_dict(cfg, seed=32), strict=True)
        data, target, emb = synth.make_inputs(cfg, batch=1, seed=34)
        runner = importlib.import_module("runners.holycow")
        E = StubEmbedder(emb)
        crit_list = [crit[n] for n in ("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice")]
        tm = runner.TrainingModule(E, G, D, crit_list, [], {})
        tm.train()
        opt_G = runner.get_optimizer(E, G, args)
        opt_D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_optimizer(D, args)
        all_dd, lG, lD = tm(dict(data), dict(target))
        loss_G, loss_D = sum(lG.values()), sum(lD.values())
        gold = {"cfg": dict(cfg), "batch": 1, "n_adain": len(G.adains)}
        for k, v in list(lG.items()) + list(lD.items()):
            gold["step.loss." + k] = v.detach().clone()
        gold["step.fake_rgbs.sub8"] = all_dd["fake_rgbs"].detach()[:, :, ::8, ::8].clone()
        opt_G.zero_grad()
        loss_G.backward(retain_graph=True)
        gp = dict(G.named_parameters())
        gold["step.gradG.norms"] = {k: p.grad.norm().item() for k, p in gp.items()}
        for k in G_SUB:
            gold["step.gradG.sub." + k] = sub(gp[k].grad)
        gold["step.gradE.scale"] = E.scale.grad.clone()
        opt_G.step()
        opt_D.zero_grad()
        loss_D.backward()
        dp = dict(D.named_parameters())
        gold["step.gradD.norms"] = {k: p.grad.norm().item() for k, p in dp.items()}
        for k in D_SUB:
            gold["step.gradD.sub." + k] = sub(dp[k].grad)
        torch.save(gold, out_dir / "step512.pt")
        print("wrote", out_dir / "step512.pt", f"{(out_dir / 'step512.pt').stat().st_size / 1e6:.2f} MB")
        for k, v in gold.items():
            if k.startswith("step.loss."):
                print(k, float(v))


if __name__ == "__main__":
    main()
