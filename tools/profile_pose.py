"""torch.profiler kernel table of the native MobileNetV2 forward (eval bs 64 = drive.py's call, train bs 8 = a
fine-tuning step's call) next to the torchvision module -> gpurun_out/pose_profile.txt"""
import copy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "latent-pose-reenactment_b200"))
import torch  # noqa: E402
import torchvision  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from embedders import mobilenet_native  # noqa: E402


def main():
    net = torchvision.models.mobilenet_v2(num_classes=256).cuda()
    out = []
    for (n, mode) in [(64, "eval"), (8, "train")]:
        x = torch.rand(n, 3, 256, 256, device="cuda")
        a = copy.deepcopy(net)
        a.train(mode == "train")
        for name, fn in [("native", lambda: mobilenet_native.forward(a, x)), ("torch", lambda: a(x))]:
            with torch.no_grad():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    fn()
                    torch.cuda.synchronize()
            ka = prof.key_averages()
            total = sum(e.self_device_time_total for e in ka)
            out.append(f"==== {name} {mode} N={n}: device time {total / 1e3:.3f} ms, {sum(e.count for e in ka)} launches")
            out.append(ka.table(sort_by="self_device_time_total", row_limit=14, max_name_column_width=70))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "pose_profile.txt").write_text("\n".join(out))
    print("\n".join(l for l in out if l.startswith("====")))


if __name__ == "__main__":
    main()
