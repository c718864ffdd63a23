"""Kernel-level view of ONE replay of the captured training step (CUPTI through torch.profiler): per-kernel-name
device time, launch counts and the share of libb200lp / torch (at::, cudnn, cublas, nccl) kernels.

    python tools/profile_replay.py [finetune|metatrain] [--batch B] -> gpurun_out/replay_profile_<workload>.txt
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "finetune"
    batch = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 8
    wl = bench.WORKLOADS[name]
    device = "cuda:0"
    torch.cuda.set_device(0)
    runner, tm, opt_G, opt_D, ns = bench.build_training(wl, device, batch)
    host = bench.make_host_batches(2, batch, wl["k_frames"], wl["num_labels"])
    dev = [({k: v.to(device) for k, v in d.items()}, {k: v.to(device) for k, v in t.items()}) for d, t in host]
    graphed = runner.GraphedTrainStep(tm, opt_G, opt_D, wl["finetune"], dev[0][0], dev[0][1])
    for i in range(3):
        graphed(*dev[i % 2])
    torch.cuda.synchronize()
    rows, total = bench.replay_kernel_times(lambda: graphed(*dev[0]))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    ours = sum(ms for n, (ms, c) in rows.items() if bench.is_own_kernel(n))
    txt = [f"workload {name} batch {batch}: one graph replay = {total:.3f} ms of kernel time in "
           f"{sum(c for _, c in rows.values())} launches; libb200lp share {ours / total:.3f}", ""]
    for n, (ms, c) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
        txt.append(f"{ms:9.3f} ms {100 * ms / total:6.2f}% {c:5d}x  {'own ' if bench.is_own_kernel(n) else 'lib '} {n[:150]}")
    (out / f"replay_profile_{name}.txt").write_text("\n".join(txt))
    print("\n".join(txt[:40]))
    print(json.dumps({"families": bench.family_table(rows)}))


if __name__ == "__main__":
    main()
