"""How much of the full-size step's gradient error is inherent to TF32 operands?  Runs the ORACLE restatement
(oracle/reference_model.py — test infrastructure) of the same step on the GPU with torch's own kernels, once in true fp32
(allow_tf32 = False) and once with cuDNN / cuBLAS TF32 (allow_tf32 = True, torch's default for convolutions), and prints
each variant's error against the CPU golden of the unmodified reference next to this repo's error
(gpurun_out/full_step_gradient_errors.json, written by tests/test_parity_full_gpu.py) -> gpurun_out/grad_error_study.json"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import reference_model as R  # noqa: E402
from oracle import synth  # noqa: E402


def sub(t):
    if t.dim() == 4:
        return t[::max(1, t.shape[0] // 16), ::max(1, t.shape[1] // 16), ::max(1, t.shape[2] // 16), ::max(1, t.shape[2] // 16)]
    if t.dim() == 2:
        return t[::max(1, t.shape[0] // 64), ::max(1, t.shape[1] // 64)]
    return t


def main():
    gold = torch.load(ROOT / "tests" / "golden" / "full_step.pt", map_location="cpu", weights_only=False)
    cfg = gold["cfg"]
    dev = "cuda"
    ours = json.loads((ROOT / "gpurun_out" / "full_step_gradient_errors.json").read_text()) \
        if (ROOT / "gpurun_out" / "full_step_gradient_errors.json").exists() else None
    res = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        g_sd = {k: v.to(dev) for k, v in synth.generator_state_dict(cfg, seed=21).items()}
        d_sd = {k: v.to(dev) for k, v in synth.discriminator_state_dict(cfg, seed=22).items()}
        for sd in (g_sd, d_sd):
            for k, v in sd.items():
                if "weight_orig" in k or k.endswith("bias") or k == "constant.constant":
                    v.requires_grad_(True)
        data, target, emb = synth.make_inputs(cfg, batch=2, seed=24)
        to = lambda t: t.to(dev)
        out, lg, ld = R.forward_losses(g_sd, d_sd, {k: to(v) for k, v in synth.vgg_state_dict("vgg19", seed=3).items()},
                                       {k: to(v) for k, v in synth.vgg_state_dict("vgg16", seed=5).items()}, cfg,
                                       to(emb["embeds"]), to(emb["pose_embedding"]), to(data["target_rgbs"][:, 0]),
                                       to(target["real_segm"][:, 0]), to(target["label"]), training=True,
                                       embeds_elemwise=to(emb["embeds_elemwise"]),
                                       criteria=("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice"))
        rec = {}
        for tag, sd, loss, retain in (("G", g_sd, sum(lg.values()), True), ("D", d_sd, sum(ld.values()), False)):
            params = {k: v for k, v in sd.items() if v.requires_grad}
            grads = torch.autograd.grad(loss, list(params.values()), retain_graph=retain, allow_unused=True)
            for (k, _), g in zip(params.items(), grads):
                if g is None:
                    continue
                ref = gold[f"step.grad{tag}.sub." + k]
                ref_norm = gold[f"step.grad{tag}.norms"][k]
                rec[k] = {"sub_rel_to_max": float((sub(g).cpu() - ref).abs().max() / (ref.abs().max() + 1e-30)),
                          "norm_rel": abs(float(g.norm()) - ref_norm) / (ref_norm + 1e-30)}
        res["torch_tf32" if tf32 else "torch_fp32"] = rec
    rows = []
    for k in res["torch_fp32"]:
        o = None
        if ours:
            o = (ours["generator"].get(k) or ours["discriminator"].get(k) or {}).get("sub_rel_to_max")
        rows.append((res["torch_tf32"][k]["sub_rel_to_max"], k, res["torch_fp32"][k]["sub_rel_to_max"], o))
    rows.sort(reverse=True)
    print(f"{'parameter':50s} {'torch fp32':>11s} {'torch tf32':>11s} {'b200lp':>11s}   (max |error| of the sub-sampled gradient / its max)")
    for t, k, f, o in rows[:40]:
        print(f"{k:50s} {f:11.2e} {t:11.2e} {o if o is None else format(o, '11.2e')}")
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "grad_error_study.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
