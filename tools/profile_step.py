"""torch.profiler view of one training step (all CUDA kernels: libb200lp + torch/cuDNN/NCCL) -> gpurun_out/step_profile.txt.
Shows where the step's GPU time and host time go (launch-bound vs kernel-bound)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "finetune"]
    device = "cuda:0"
    torch.cuda.set_device(0)
    runner, tm, opt_G, opt_D, ns = bench.build_training(wl, device, 8)
    host = bench.make_host_batches(2, 8, wl["k_frames"], wl["num_labels"])
    dev = [({k: v.to(device) for k, v in d.items()}, {k: v.to(device) for k, v in t.items()}) for d, t in host]

    def step(i):
        d, t = dev[i % 2]
        runner.train_step(tm, dict(d), dict(t), opt_G, opt_D, finetune=wl["finetune"])

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 5
    stacks = "--stacks" in sys.argv
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=stacks, record_shapes=stacks) as prof:
        step(0)
        torch.cuda.synchronize()
    if stacks:      # who issues the small torch glue ops (copy_, add_, mul, fill_)?
        rows = []
        for e in prof.key_averages(group_by_stack_n=12):
            if e.key in ("aten::copy_", "aten::add_", "aten::mul", "aten::fill_", "aten::add", "aten::zero_", "aten::clone"):
                user = [f for f in e.stack if "site-packages" not in f and "profile_step" not in f][:3]
                rows.append((e.count, e.key, " <- ".join(x.strip()[-90:] for x in user)))
        for e in prof.key_averages(group_by_input_shape=True):
            if e.key in ("aten::copy_", "aten::add_", "aten::mul", "aten::fill_", "aten::add", "aten::zero_", "aten::clone"):
                rows.append((e.count, e.key, "shapes " + str(e.input_shapes)[:120]))
        rows.sort(reverse=True)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "step_glue_stacks.txt").write_text("\n".join(f"{c:5d} {k:14s} {s}" for c, k, s in rows[:80]))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    ka = prof.key_averages()
    cuda_total = sum(e.self_device_time_total for e in ka) / 1e3
    n_kernels = sum(e.count for e in ka if e.self_device_time_total > 0)
    txt = [f"wall per step (no profiler): {wall * 1e3:.2f} ms", f"sum of device kernel time in one step: {cuda_total:.2f} ms",
           f"device-side launches in one step: {n_kernels}", "",
           ka.table(sort_by="self_cuda_time_total", row_limit=80, max_name_column_width=90)]
    (out / "step_profile.txt").write_text("\n".join(txt))
    print("\n".join(txt[:3]))


if __name__ == "__main__":
    main()
