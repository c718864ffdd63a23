"""Every C-ABI kernel against plain torch ops in float64 (the checks of tools/gpu_diag.py, run in-process).
Inputs are pre-rounded to TF32 where they feed the tensor cores, so the tolerance isolates kernel correctness
(accumulation order only: <= 2e-5 relative; 5e-5 for weight gradients summed over 524288 pixels)."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import gpu_diag  # noqa: E402

pytestmark = pytest.mark.gpu

NAMES = ["conv_basic", "conv_3x3", "conv_small_planes", "conv_epilogue", "conv_big", "conv_dgrad_pack", "wgrad_basic",
         "wgrad_3x3", "wgrad_big", "adain_fwd_bwd", "elementwise_misc", "direct_convs", "conv_bf16x3", "adain_split", "sn_kernels", "fused_optim", "conv_splitk", "conv_halo",
         "pack_multi", "conv_pair", "wgrad_sn_acc", "pose_encoder", "tail_tensor_core", "encoder_bn", "encoder_gconv",
         "encoder_misc", "identity_encoder", "losses_kernels", "pose_bwd_kernels", "gconv_tc", "relu_bwd_fused", "l1_code", "c3_tensor_core", "adain_fused", "wgrad_ragged_batch"]


@pytest.mark.parametrize("name", NAMES)
def test_kernel_group(name):
    results = gpu_diag.CHECKS[name]()
    bad = [r for r in results if not r.get("ok")]
    assert not bad, bad


def test_shape_validation_errors():
    """Unsupported shapes are rejected with an error code + message (no launch, no crash)."""
    import torch
    from b200lp import kernels as K
    from b200lp.lib import B200lpError
    x = torch.zeros(1, 16, 16, 48, device="cuda")          # Cin not a multiple of 32
    w = torch.zeros(64, 1, 48, device="cuda")
    with pytest.raises(B200lpError, match="multiples of 32"):
        K.conv_fwd(x, w, 1)
    x = torch.zeros(1, 12, 12, 32, device="cuda")          # H, W not powers of two
    with pytest.raises(B200lpError, match="powers of two"):
        K.conv_fwd(x, torch.zeros(32, 9, 32, device="cuda"), 3)
    with pytest.raises(B200lpError):
        K.conv_fwd(torch.zeros(1, 16, 16, 32), torch.zeros(32, 9, 32), 3)   # CPU tensors: no fallback
