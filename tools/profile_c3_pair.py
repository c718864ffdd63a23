import sys
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent / "latent-pose-reenactment_b200")); sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent))
import torch
from b200lp import kernels as K
x = torch.randn(8, 3, 256, 256, device="cuda"); w = torch.randn(64, 3, 3, 3, device="cuda"); b = torch.randn(64, device="cuda")
for _ in range(3):
    K.conv3x3_c3_fwd(x, w, bias=b, relu=True, round_tf32=True)
for (H, Cin, Cout, v) in [(128, 128, 128, 12), (64, 512, 256, 11)]:
    xx = torch.randn(8, H, H, Cin, device="cuda"); wp = K.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device="cuda"))
    y = torch.empty(8, H, H, Cout, device="cuda")
    for _ in range(2):
        K.conv_fwd(xx, wp, 3, out=y, variant=v, splits=1)
torch.cuda.synchronize()
