"""ORACLE — TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Times the UNMODIFIED reference's training step on the host CPU.

    python oracle/ref_bench.py --workload finetune|metatrain --batch 8 --steps K --warmup W [--budget-s S] [--threads T]

Imports the reference's own modules (generators / discriminators / embedders / criterions plugins and runners.holycow)
from a reference tree — `baseline/_ref/` inside this repo (a verbatim, git-ignored copy staged by
`__graft_entry__.build()` when /root/reference exists; it travels to the GPU box with the snapshot), else
`/root/reference` — builds the networks through `Wrapper.get_net`, and runs the step of
`runners/holycow.py:230-257` exactly as `run_epoch` does: `training_module(data, target)`, `loss_G.backward(retain_graph=
True)`, `optimizer_G.step()`, `loss_D.backward()`, `optimizer_D.step()`, `update_running_average`.  Prints one JSON
object.  Runs as its own process: the reference's top-level package names (`generators`, `criterions`, ...) collide
with this repo's plugin tree, so the two can never be imported together.

Shims (nothing in the reference is edited): a `yamlenv` module, fabricated VGG weight files in the
layout criterions/common/perceptual_loss.py expects (there is no network to download the real ones; timing does not
depend on the values), synthetic batches with the dataloader's output contract.
"""
import argparse
import json
import os
import sys
import tempfile
import time
import types
from argparse import Namespace
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

WORKLOADS = {   # mirrors bench.py WORKLOADS (BASELINE.json configs[1] / configs[2])
    "finetune": dict(finetune=True, criteria="adversarial, featmat, idt_embed, perceptual, dice", optimizer="RAdam",
                     lr_gen=5e-4, lr_dis=8e-4, k_frames=1, num_labels=1),
    "metatrain": dict(finetune=False, criteria="idt_embed, perceptual, adversarial, featmat, dis_embed, dice",
                      optimizer="Adam", lr_gen=5e-5, lr_dis=2e-4, k_frames=8, num_labels=16),
}
WORKLOADS["metatrain512"] = WORKLOADS["metatrain"]      # configs[4]: the caller passes --image-size 512 --batch 4
# configs[0] / configs[3]: drive.py's inner loop (drive.py:84-98) — fine-tuned generator + pose embedder, no discriminator
WORKLOADS["drive"] = dict(WORKLOADS["finetune"], drive=True)


def find_reference():
    for cand in (os.environ.get("LPR_REFERENCE"), REPO / "baseline" / "_ref", "/root/reference"):
        if cand and (Path(cand) / "generators" / "vector_pose_unsupervised_segmentation_noBottleneck.py").is_file():
            return Path(cand)
    return None


def fabricate_vgg_files(dirname, torch, seed=3):
    import torchvision
    g = torch.Generator().manual_seed(seed)
    m = torchvision.models.vgg19()
    full = m.state_dict()
    ren = {"classifier.0": "classifier.1", "classifier.3": "classifier.4"}
    out = {}
    for k, v in full.items():
        if k.startswith("features.") and v.dim() == 4:
            v = torch.randn(v.shape, generator=g) * (2.0 / (9 * v.shape[1])) ** 0.5
        for a, b in ren.items():
            if k.startswith(a + "."):
                k = b + k[len(a):]
        out[k] = v
    torch.save(out, os.path.join(dirname, "vgg19-d01eb7cb.pth"))
    f = torchvision.models.vgg16().features.state_dict()
    torch.save({k: (torch.randn(v.shape, generator=g) * (2.0 / (9 * v.shape[1])) ** 0.5 if v.dim() == 4 else v)
                for k, v in f.items()}, os.path.join(dirname, "vgg_face_weights.pth"))


def drive_loop(args, torch, E, G, cores, ref):
    """The reference's drive.py inner loop (drive.py:84-98) on synthetic driver frames: pose embedding -> generator in
    fine-tuned mode (drive.py:52,63-70; eval mode = `set_eval_mode_in_test`) -> permute / clamp_ / mul_ / byte of the
    result and of the driver frame, concatenated side by side.  One "step" = one batch of `--batch` frames (the
    reference hard-codes 1, drive.py:57).  Video encoding is left out on both arms."""
    import numpy as np
    S, B = args.image_size, args.batch
    G.enable_finetuning({"embeds": torch.randn(1, 512)})
    E.enable_finetuning()
    E.eval(); G.eval()
    g = torch.Generator().manual_seed(5)
    frames = [torch.rand(B, 1, 3, S, S, generator=g) for _ in range(2)]

    def to_u8(image):
        return image.permute(1, 2, 0).clamp_(0, 1).mul_(255).cpu().byte().numpy()

    def step(i):
        with torch.no_grad():
            d = {"pose_input_rgbs": frames[i % 2].clone()}
            E.get_pose_embedding(d)
            G(d)
            out = [np.concatenate((to_u8(d["pose_input_rgbs"][b, 0]), to_u8(d["fake_rgbs"][b])), axis=1) for b in range(B)]
        return out

    for i in range(args.warmup):
        step(i)
    times = []
    t_start = time.perf_counter()
    for i in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        step(i)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > args.budget_s:
            break
    dt = sum(times) / len(times)
    print(json.dumps({"frames_per_s": B / dt, "s_per_step": dt, "steps": len(times), "warmup": args.warmup, "batch": B,
                      "cores": cores, "reference_tree": str(ref), "kind": "reference",
                      "spread": (max(times) - min(times)) / dt if len(times) > 1 else 0.0,
                      "torch_threads": torch.get_num_threads()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="finetune", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--budget-s", type=float, default=240.0, help="stop timing after this many seconds (>= 1 step)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--image-size", type=int, default=256)
    args = ap.parse_args()

    ref = find_reference()
    if ref is None:
        print(json.dumps({"unavailable": "no reference tree (baseline/_ref or /root/reference)"}))
        return
    sys.path.insert(0, str(ref))
    import yaml
    shim = types.ModuleType("yamlenv")
    shim.load = yaml.safe_load
    sys.modules.setdefault("yamlenv", shim)
    import importlib

    import torch
    wl = WORKLOADS[args.workload]
    runner = importlib.import_module("runners.holycow")          # the reference's runner (sets OMP threads to 1 ...)
    cores = args.threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)        # ... overridden: all host cores (the reference pins 1, utils/utils.py:19)
    S, B, K = args.image_size, args.batch, wl["k_frames"]
    with tempfile.TemporaryDirectory() as vgg_dir:
        fabricate_vgg_files(vgg_dir, torch)
        ns = Namespace(
            gen_padding="zero", in_channels=3, out_channels=3, num_channels=64, max_num_channels=512, embed_channels=512,
            pose_embedding_size=256, norm_layer="in", gen_constant_input_size=4, gen_num_residual_blocks=2, image_size=S,
            device="cpu", average_function="sum", dis_padding="zero", dis_num_blocks=7, num_labels=wl["num_labels"],
            gan_type="gan", fm_weight=10.0, dice_weight=1.0, perc_weight=3e-2, idt_embed_weight=6e-3,
            dis_embed_weight=1e-2, vgg_weights_dir=vgg_dir, optimizer=wl["optimizer"], lr_gen=wl["lr_gen"],
            lr_dis=wl["lr_dis"], beta1=0.0, finetune=wl["finetune"], num_gpus=1, batch_size=B)
        torch.manual_seed(123)
        G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(ns)
        Dm = importlib.import_module("discriminators.no_landmarks")
        D = Dm.Wrapper.get_net(ns)
        E = importlib.import_module("embedders.unsupervised_pose_separate_embResNeXt_segmentation").Wrapper.get_net(ns)
        crits = [importlib.import_module(f"criterions.{c.strip()}").Wrapper.get_net(ns) for c in wl["criteria"].split(",")]
    if wl.get("drive"):
        return drive_loop(args, torch, E, G, cores, ref)
    if wl["finetune"]:           # train.py:240-279
        e = torch.randn(1, 512)
        G.enable_finetuning({"embeds": e.clone()})
        D.enable_finetuning({"embeds": e.clone()})
        E.enable_finetuning()
    tm = runner.TrainingModule(E, G, D, crits, [], {})
    tm.train()
    opt_G = runner.get_optimizer(E, G, ns)
    opt_D = Dm.Wrapper.get_optimizer(D, ns)

    g = torch.Generator().manual_seed(123)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    disc = (((yy - S / 2) ** 2 + (xx - S / 2) ** 2) <= (0.4 * S) ** 2).float()
    data = {"enc_rgbs": torch.rand(B, K, 3, S, S, generator=g), "pose_input_rgbs": torch.rand(B, 1, 3, S, S, generator=g),
            "target_rgbs": torch.rand(B, 1, 3, S, S, generator=g) * disc}
    target = {"real_segm": disc.expand(B, 1, 3, S, S).contiguous(),
              "label": torch.randint(0, max(wl["num_labels"], 1), (B,), generator=g)}
    alpha = 0.972 if wl["finetune"] else 0.999

    def step():                   # runners/holycow.py:230-257, num_gpus == 1 (no reducer)
        all_data, losses_G, losses_D = tm(data, target)
        loss_G = sum(losses_G.values())
        loss_D = sum(losses_D.values())
        opt_G.zero_grad()
        loss_G.backward(retain_graph=True)
        opt_G.step()
        opt_D.zero_grad()
        loss_D.backward()
        opt_D.step()
        tm.update_running_average(alpha)
        return float(loss_G.detach()) + float(loss_D.detach())

    for _ in range(args.warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > args.budget_s:
            break
    dt = sum(times) / len(times)
    print(json.dumps({"frames_per_s": B / dt, "s_per_step": dt, "steps": len(times), "warmup": args.warmup, "batch": B,
                      "cores": cores, "reference_tree": str(ref), "kind": "reference",
                      "spread": (max(times) - min(times)) / dt if len(times) > 1 else 0.0,
                      "torch_threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
