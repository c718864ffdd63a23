"""bench.py's host-side bookkeeping, without a GPU: kernel-name -> family classification over the kernel names of the
committed replay profiles, the roofline object from stubbed kernel durations, the workload table."""
import re
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

TENSOR = {"conv_igemm_tf32", "conv_igemm_bf16x3", "conv_wgrad_tf32", "resnext_grouped"}


def _profile_names(fname):
    names = []
    for line in (ROOT / "profiles" / fname).read_text().splitlines():
        m = re.match(r"\s*[\d.]+ ms\s+[\d.]+%\s+\d+x\s+(own|lib)\s+(.*)$", line)
        if m:
            names.append((m.group(1), m.group(2).strip()))
    return names


@pytest.mark.parametrize("fname", ["r02_replay_profile_finetune.txt", "r02_replay_profile_metatrain.txt"])
def test_every_kernel_of_a_replayed_step_has_a_family(fname):
    names = _profile_names(fname)
    assert len(names) > 60
    fams = set()
    for own, name in names:
        fam = bench.kernel_family(name)
        fams.add(fam)
        assert bench.is_own_kernel(name) == (own == "own"), name
        if own == "lib":
            assert fam in ("torch/cudnn/cublas", "nccl", "memcpy/memset"), (name, fam)
        if "conv_halo2_kernel<256, 1, 0>" in name or "conv_igemm_kernel<128, 0>" in name:
            assert fam == "conv_igemm_tf32"
        if "conv_halo2_kernel<128, 1, 1>" in name:
            assert fam == "conv_igemm_bf16x3"
        if "conv_wgrad_tf32_kernel" in name:
            assert fam == "conv_wgrad_tf32"
        if "adain_bwd_" in name:
            assert fam == "adain_relu_bwd"
        if "in_stats_" in name:
            assert fam == "in_stats"
    assert TENSOR - {"resnext_grouped"} <= fams


def test_roofline_object_from_stubbed_kernel_times(monkeypatch):
    rows = {"void b200lp::conv_halo2_kernel<256, 1, 0>(b200lp::ConvMaps, b200lp::ConvParams)": (7.0, 183),
            "void b200lp::conv_wgrad_tf32_kernel<256>(CUtensorMap_st, CUtensorMap_st, b200lp::WgradParams)": (2.0, 60),
            "void b200lp::wgrad_reduce_acc_kernel<9>(float const*)": (1.0, 20),
            "void b200lp::adain_relu_kernel<true, true>(float const*)": (0.5, 17),
            "void at::native::vectorized_elementwise_kernel<4, at::native::FillFunctor<float> >(int)": (0.5, 20)}
    monkeypatch.setattr(bench, "replay_kernel_times", lambda fn: (rows, 11.0))
    work = {"conv_igemm_tf32": {"flops": 4.0e12}, "conv_wgrad_tf32": {"flops": 0.8e12}, "adain_relu": {"bytes": 2.0e9}}
    peaks = {"bf16_tflops_sustained": 1360.0, "hbm_gbs": 6536.4}
    out = bench.roofline_from_replay(lambda: None, work, peaks, workload="finetune")
    r = out["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == 680.0
    assert abs(r["achieved"] - 4.0e12 / 7.0e-3 / 1e12) < 1.0 and abs(r["frac"] - r["achieved"] / 680.0) < 1e-3
    assert r["launches_per_step"] == 183 and abs(r["avg_launch_ms"] - 7.0 / 183) < 1e-3
    assert isinstance(r["traffic"], int) and r["traffic"] > 1e6          # the fine-tune step's committed ncu capture
    assert abs(out["replay"]["libb200lp_share"] - 10.5 / 11.0) < 1e-3
    assert out["families"]["conv_wgrad_tf32"]["tflops"] == pytest.approx(400.0, rel=1e-3)
    # a workload without a committed ncu capture reports no DRAM traffic instead of another workload's
    assert bench.roofline_from_replay(lambda: None, work, peaks, workload="metatrain512")["roofline"]["traffic"] is None
    mt = bench.roofline_from_replay(lambda: None, work, peaks, workload="metatrain")["roofline"]
    assert mt["traffic"] != r["traffic"] and "metatrain" in mt["traffic_note"]


def test_workload_table_matches_baseline_configs():
    import json
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert len(base["configs"]) == 5
    w = bench.WORKLOADS
    assert w["finetune"]["finetune"] and w["finetune"]["optimizer"] == "RAdam" and "dis_embed" not in w["finetune"]["criteria"]
    assert not w["metatrain"]["finetune"] and w["metatrain"]["k_frames"] == 8 and w["metatrain"]["optimizer"] == "Adam"
    assert w["metatrain512"]["image_size"] == 512 and w["metatrain512"]["batch"] == 4
    ns = bench.make_namespace(w["metatrain512"], "cpu", "/nonexistent", 4)
    assert ns.image_size == 512 and ns.batch_size == 4 and ns.num_channels == 64 and ns.max_num_channels == 512
    assert bench.make_namespace(w["finetune"], "cpu", "/nonexistent", 8).image_size == 256
