"""Kernel-level device time of the identity encoder's forward + backward (64 x 256x256, train mode) through CUPTI
(torch.profiler): native schedule vs the torchvision module on cuDNN -> gpurun_out/identity_profile.txt"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    import torchvision
    from embedders import resnext_native
    dev = "cuda"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    torch.manual_seed(0)
    net = torchvision.models.resnext50_32x4d(num_classes=512).to(dev).train()
    x = torch.rand(n, 3, 256, 256, device=dev)
    wgt = torch.randn(n, 512, device=dev)

    def native():
        y = resnext_native.apply(net, x)
        (y * wgt).sum().backward()

    def stock():
        y = net(x)
        (y * wgt).sum().backward()

    txt = []
    for name, fn in (("native", native), ("torch+cudnn", stock)):
        for _ in range(2):
            fn()
        rows, total = bench.replay_kernel_times(fn)
        txt.append(f"== {name}: {total:.3f} ms of kernel time in {sum(c for _, c in rows.values())} launches")
        fam = {}
        for k, (ms, c) in rows.items():
            f = bench.kernel_family(k)
            fam[f] = (fam.get(f, (0, 0))[0] + ms, fam.get(f, (0, 0))[1] + c)
        for f, (ms, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
            txt.append(f"   family {f:24s} {ms:8.3f} ms {c:5d}x")
        for k, (ms, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
            txt.append(f"{ms:9.3f} ms {100 * ms / total:6.2f}% {c:5d}x  {k[:140]}")
        txt.append("")
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "identity_profile.txt").write_text("\n".join(txt))
    print("\n".join(txt))


if __name__ == "__main__":
    main()
