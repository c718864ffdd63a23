"""Short launch sequence of the dominant kernels at the step's dominant shapes, for `ncu --set full` (kept short: ncu
replays every kernel ~40 times).  Order of launches after the warm-up (each shape: 2 warm-up + 1 profiled launch):
    conv_igemm tf32  : 64->64 @256^2, 128->128 @128^2, 512->256 @64^2          (bs 8)
    conv_igemm bf16x3: 128->128 @128^2
    conv_wgrad tf32  : 64->64 @256^2, 128->128 @128^2
    adain_relu       : 64 ch @256^2 (no upsample), 128 ch @128^2 -> 256^2 (upsample)
    in_stats         : 64 ch @256^2
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "latent-pose-reenactment_b200"))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402
from b200lp import kernels as K  # noqa: E402
from gpu_diag import split_bf16  # noqa: E402


def main():
    dev = "cuda"
    torch.manual_seed(0)
    for (H, Cin, Cout) in [(256, 64, 64), (128, 128, 128), (64, 512, 256)]:
        x = torch.randn(8, H, H, Cin, device=dev)
        wp = K.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev))
        y = torch.empty(8, H, H, Cout, device=dev)
        for _ in range(2):
            K.conv_fwd(x, wp, 3, out=y)
    x = torch.randn(8, 128, 128, 128, device=dev)
    xs = split_bf16(x)
    wps = K.pack_conv_weight(torch.randn(128, 128, 3, 3, device=dev), precision=K.BF16X3)
    for _ in range(2):
        K.conv_fwd(xs, wps, 3)
    for (H, C) in [(256, 64), (128, 128)]:
        x = torch.randn(8, H, H, C, device=dev)
        dy = torch.randn(8, H, H, C, device=dev)
        for _ in range(2):
            K.conv_wgrad(x, dy, 3)
    aff = torch.randn(8, 256, device=dev)
    x = torch.randn(8, 256, 256, 64, device=dev)
    for _ in range(2):
        mean, rstd = K.in_stats(x, 1e-4)
        K.adain_relu(x, mean, rstd, aff[:, 64:128], aff[:, :64])
    x = torch.randn(8, 128, 128, 128, device=dev)
    mean, rstd = K.in_stats(x, 1e-4)
    for _ in range(2):
        K.adain_relu(x, mean, rstd, aff[:, 128:256], aff[:, :128], upsample2=True)
    # round-1 additions: fused weight-gradient path (tensor-core kernel -> tiled reduce + accumulate + <G,W> partials ->
    # rank-1 term), the generator tail on the tensor cores, the pose-encoder kernels
    import torch.nn.functional as F
    for (H, Cin, Cout) in [(32, 512, 512), (128, 128, 128)]:
        x = torch.randn(8, H, H, Cin, device=dev)
        dy = torch.randn(8, H, H, Cout, device=dev)
        w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.02
        grad = torch.zeros_like(w)
        u = F.normalize(torch.randn(Cout, device=dev), dim=0)
        v = F.normalize(torch.randn(Cin * 9, device=dev), dim=0)
        s_ = torch.ones(1, device=dev)
        for _ in range(2):
            K.conv_wgrad_sn_acc(x, dy, 3, grad, w, s_, u, v)
    from b200lp import ops
    x = torch.randn(8, 256, 256, 64, device=dev)
    w4 = torch.randn(4, 64, 3, 3, device=dev) * 0.05
    with torch.no_grad():
        for _ in range(2):
            ops.adain_tail(x, aff[:, 64:128], aff[:, :64], w4, torch.ones(1, device=dev), torch.zeros(4, device=dev))
    for (m, cin, cout) in [(131072, 16, 96), (2048, 576, 96), (512, 960, 320)]:
        xm = torch.randn(m, cin, device=dev)
        wm = torch.randn(cout, cin, device=dev)
        sc = torch.rand(cin, device=dev)
        sh = torch.randn(cin, device=dev)
        for _ in range(2):
            K.pw_conv(xm, wm, sc, sh, True, want_stats=True)
    for (h, c, stride) in [(128, 96, 2), (64, 144, 1)]:
        xd = torch.randn(8, h, h, c, device=dev)
        wd = torch.randn(c, 1, 3, 3, device=dev)
        sc = torch.rand(c, device=dev)
        sh = torch.randn(c, device=dev)
        for _ in range(2):
            K.dw_conv3x3(xd, wd, sc, sh, stride, want_stats=True)
    # round-2 additions: grouped 3x3 gradients on the tensor cores (ResNeXt layer1 / layer3 shapes, 64 identity frames),
    # the discriminator's fused ReLU-backward pass, the code-byte VGG tap, BatchNorm backward, tiled weight packing
    for (h, cpg) in [(64, 4), (16, 16)]:
        c = 32 * cpg
        xg = torch.randn(64, h, h, c, device=dev)
        dyg = torch.randn(64, h, h, c, device=dev)
        wg = torch.randn(c, cpg, 3, 3, device=dev) * 0.1
        wpt = K.pack_gconv_weight(wg, transpose=True)
        for _ in range(2):
            K.gconv3x3_dgrad(dyg, wg, (h, h), packed=wpt)
            K.gconv3x3_wgrad_tc(xg, dyg, cpg)
    yr = torch.randn(8, 128, 128, 64, device=dev).relu()
    dyr = torch.randn_like(yr)
    addr = torch.randn_like(yr)
    ba, bb = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    for _ in range(2):
        K.relu_bwd_fused(yr, dyr, add=addr, want_quarter=True, bias_a=ba, bias_b=bb, round_tf32=True)
    fa = torch.randn(8, 128, 128, 128, device=dev).relu()
    fb = torch.randn(8, 128, 128, 128, device=dev).relu()
    loss = torch.zeros(1, device=dev)
    gs = torch.ones(1, device=dev)
    for _ in range(2):
        code = K.l1_sum_code(fa, fb, loss, 1e-6)
        K.l1_code_bwd(code, tuple(fa.shape), gs, 1e-6, d_in=fb)
    xb = torch.randn(64 * 64 * 64, 256, device=dev)
    dyb = torch.randn_like(xb)
    mean, rstd = torch.zeros(256, device=dev), torch.ones(256, device=dev)
    gam = torch.ones(256, device=dev)
    for _ in range(2):
        K.bn_bwd(dyb, xb, mean, rstd, gam, gam, mean, mask_mode=2)
    ws_ = [torch.randn(512, 512, 3, 3, device=dev), torch.randn(256, 128, 3, 3, device=dev)]
    rows = []
    keep = []
    for w_ in ws_:
        for tr in (False, True):
            o = K.pack_conv_weight(w_, transpose=tr)
            keep.append(o)
            rows.append((w_.data_ptr(), o.data_ptr(), w_.shape[0], w_.shape[1], 9, int(tr), 0, w_.numel()))
    plan = K.pack_plan(rows, torch.device(dev))
    for _ in range(2):
        K.pack_conv_weight_multi(plan)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
