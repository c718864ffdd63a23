"""Benchmark of the hot path (BASELINE.json metric: 256x256 reenactment frames/s, train step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload finetune|metatrain|drive] [--impl reference]

One "step" = one full optimisation step of the reference's runner (runners/holycow.py:224-257) on a batch of 8
synthetic 256x256 samples per GPU: embedder -> generator -> discriminator x3 -> criteria -> loss_G.backward ->
optimizer_G -> loss_D.backward -> optimizer_D -> EMA.  Default workload = BASELINE.json configs[1]
(`finetuning-base`: criteria adversarial+featmat+idt_embed+perceptual+dice, RAdam, EMA 0.972, identity encoder off),
per-GPU batch fixed as N grows (weak scaling, one NCCL all-reduce per backward).  Under torchrun (N > 1) the workload
is BASELINE.json configs[2] (meta-training, the only configuration the reference runs data-parallel); the N = 1 line
reports that workload too (`e2e.metatrain`).

Prints ONE JSON line (rank 0).  `value` = frames/s with the batches already resident in HBM; `e2e` = the same metric
through the public plugin API with pinned HOST batches (H2D inside the timed region, loss values read back).
`--impl reference` times the UNMODIFIED reference's step (its own modules, imported from baseline/_ref by
oracle/ref_bench.py in a separate process) on the host cores, same workload and batch.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "latent-pose-reenactment_b200"
for _p in (str(PKG), str(ROOT)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch

METRIC = "256x256 reenactment frames/sec (train step)"
UNIT = "frames/s"
FULL = dict(num_channels=64, max_num_channels=512, embed_channels=512, pose_embedding_size=256, image_size=256,
            dis_num_blocks=7)
WORKLOADS = {
    # BASELINE.json configs[1]
    "finetune": dict(desc="configs[1] finetuning-base: bs=8/GPU 256x256, adversarial+featmat+idt_embed+perceptual+dice, "
                          "RAdam, EMA 0.972, identity encoder off",
                     finetune=True, criteria="adversarial, featmat, idt_embed, perceptual, dice", optimizer="RAdam",
                     lr_gen=5e-4, lr_dis=8e-4, k_frames=1, num_labels=1),
    # BASELINE.json configs[2]
    "metatrain": dict(desc="configs[2] default: bs=8/GPU 256x256, K=8 identity frames, all six criteria, Adam, EMA 0.999",
                      finetune=False, criteria="idt_embed, perceptual, adversarial, featmat, dis_embed, dice",
                      optimizer="Adam", lr_gen=5e-5, lr_dis=2e-4, k_frames=8, num_labels=16),
    # BASELINE.json configs[4]: a stress configuration, not the headline — only run when asked for by name
    # (`--workload metatrain512`; per-GPU batch 4 unless --batch is given).  `value` is then 512x512 frames/s.
    "metatrain512": dict(desc="configs[4]: bs=4/GPU 512x512 (8 up-blocks, 19 AdaIN sites), K=8 identity frames, all six "
                              "criteria, Adam, EMA 0.999",
                         finetune=False, criteria="idt_embed, perceptual, adversarial, featmat, dis_embed, dice",
                         optimizer="Adam", lr_gen=5e-5, lr_dis=2e-4, k_frames=8, num_labels=16, image_size=512, batch=4),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS) + ["drive"],
                    help="default: finetune (BASELINE configs[1]) on one GPU, metatrain (configs[2]) under torchrun")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 8; 4 for metatrain512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--timed-only", action="store_true",
                    help="only the warm-up and the timed steps (no e2e / drive / roofline / cpu passes): for profilers")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
def make_namespace(wl, device, vgg_dir, batch):
    from argparse import Namespace
    return Namespace(
        gen_padding="zero", in_channels=3, out_channels=3, norm_layer="in", gen_constant_input_size=4,
        gen_num_residual_blocks=2, device=device, average_function="sum", dis_padding="zero",
        num_labels=wl["num_labels"], gan_type="gan", fm_weight=10.0, dice_weight=1.0, perc_weight=3e-2,
        idt_embed_weight=6e-3, dis_embed_weight=1e-2, vgg_weights_dir=vgg_dir, optimizer=wl["optimizer"],
        lr_gen=wl["lr_gen"], lr_dis=wl["lr_dis"], beta1=0.0, finetune=wl["finetune"], num_gpus=1, batch_size=batch,
        **dict(FULL, image_size=wl.get("image_size", FULL["image_size"])))


def fabricate_vgg_files(dirname, seed=3):
    """No network: VGG19 / VGG-Face weight files with He-initialised weights in the layout the criterion reads."""
    g = torch.Generator().manual_seed(seed)
    plans = {"vgg19": ((0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28), (64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512)),
             "vgg16": ((0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28), (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512))}
    for arch, (idxs, chans) in plans.items():
        sd, cin = {}, 3
        for i, c in zip(idxs, chans):
            w = torch.randn(c, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
            sd[f"{i}.weight"] = w / 64.0 if i == 0 else w
            sd[f"{i}.bias"] = torch.randn(c, generator=g) * 0.05
            cin = c
        if arch == "vgg19":
            torch.save({"features." + k: v for k, v in sd.items()}, os.path.join(dirname, "vgg19-d01eb7cb.pth"))
        else:
            torch.save(sd, os.path.join(dirname, "vgg_face_weights.pth"))


def make_host_batches(n, batch, k_frames, num_labels, s=256, seed=123, rank=0):
    """Pinned host batches with the dataloader's output contract (dataloaders/synthetic.py)."""
    from dataloaders.synthetic import Dataset
    ds = Dataset(n * batch * 64, s, k_frames, max(num_labels, 1), seed=seed)
    out = []
    for b in range(n):
        items = [ds[(rank * n + b) * batch + i] for i in range(batch)]
        data = {k: torch.stack([it[0][k] for it in items]) for k in items[0][0]}
        target = {"real_segm": torch.stack([it[1]["real_segm"] for it in items]),
                  "label": torch.tensor([it[1]["label"] for it in items], dtype=torch.long)}
        if torch.cuda.is_available():
            data = {k: v.pin_memory() for k, v in data.items()}
            target = {k: v.pin_memory() for k, v in target.items()}
        out.append((data, target))
    return out


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.thread = gpu_index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def timed(fn, steps, dist_on):
    """K calls of fn(i) between barrier+synchronize on both sides; CUDA-event time, max over ranks (ms)."""
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist_on:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    return float(ms.item())


# ----------------------------------------------------------------------------------------------------------------
def reference_cpu_run(workload, batch, steps, warmup, budget_s):
    """The UNMODIFIED reference's step on the host cores (oracle/ref_bench.py in its own process: the reference's package
    names collide with this repo's plugin tree).  Returns its JSON dict, or None when no reference tree travels with the
    repo (baseline/_ref, staged by __graft_entry__.build(), or /root/reference)."""
    # torch CPU convolutions stop scaling (and then thrash) well below the 128 hardware threads some hosts expose:
    # 132.8 s/step with 128 threads (profiles/r01_bench_first_run.json) — use at most 32 and say so
    cores = min(os.cpu_count() or 1, 32)
    cmd = [sys.executable, str(ROOT / "oracle" / "ref_bench.py"), "--workload", workload, "--batch", str(batch), "--steps",
           str(steps), "--warmup", str(warmup), "--budget-s", str(budget_s), "--threads", str(cores), "--image-size",
           str(WORKLOADS.get(workload, {}).get("image_size", FULL["image_size"]))]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s * 3 + 600, env=env).stdout
    except Exception:
        return None
    for line in reversed(out.strip().splitlines()):
        if line.startswith("{"):
            d = json.loads(line)
            return None if "unavailable" in d else d
    return None


def cpu_port_rate(workload, sample_batch, steps, warmup, budget_s=25.0):
    """Fallback when the reference tree is absent: frames/s of the CPU port of the step (oracle/cpu_step.py)."""
    from oracle.cpu_step import OracleTrainer
    from oracle import synth
    wl = WORKLOADS[workload]
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    cfg = dict(synth.FULL_CFG, image_size=wl.get("image_size", FULL["image_size"]))
    cfg["num_labels"] = max(wl["num_labels"], 1)
    crit = tuple(c.strip() for c in wl["criteria"].split(","))
    tr = OracleTrainer(cfg, finetune=wl["finetune"], criteria=crit, optimizer=wl["optimizer"], lr_gen=wl["lr_gen"],
                       lr_dis=wl["lr_dis"])
    data, target, _ = synth.make_inputs(cfg, batch=sample_batch, seed=4, n_identity_frames=wl["k_frames"])
    for _ in range(warmup):
        tr.step(data, target)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        tr.step(data, target)
        done += 1
        if time.perf_counter() - t0 > budget_s:      # bounded sample: stop after ~budget_s seconds of CPU work
            break
    dt = (time.perf_counter() - t0) / done
    return {"frames_per_s": sample_batch / dt, "s_per_step": dt, "steps": done, "warmup": warmup, "batch": sample_batch,
            "cores": cores, "kind": "port", "spread": 0.0}


def cpu_step_rate(workload, batch, steps, warmup, budget_s):
    r = reference_cpu_run(workload, batch, steps, warmup, budget_s)
    if r is None:
        r = cpu_port_rate(workload, batch, steps, warmup, budget_s)
        r["note"] = "reference tree absent (no baseline/_ref): CPU port of the step (oracle/cpu_step.py)"
    return r


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the step, SAME workload and per-GPU batch as this
    repo's arm, 1 warm-up + as many of the requested steps as fit the time budget (>= 3 when possible); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    wl_name = args.workload or ("finetune" if world == 1 else "metatrain")
    if wl_name == "drive":
        run_reference_drive(args)
        return
    warmup = 1 if args.warmup > 0 else 0
    args.batch = args.batch or WORKLOADS[wl_name].get("batch", 8)
    r = cpu_step_rate(wl_name, args.batch, max(args.steps, 3), warmup, budget_s=200.0)
    rate, dt = r["frames_per_s"], r["s_per_step"]
    sample = (f"{r['steps']} step(s) after {r['warmup']} warm-up at batch {r['batch']} = the per-GPU batch of the product arm, "
              f"fp32, {r['cores']} torch threads (reference pins 1, utils/utils.py:19: overridden), "
              f"run-to-run spread {100 * r.get('spread', 0.0):.1f} %")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl_name]["desc"], "global_batch": args.batch, "per_gpu_batch": args.batch,
                       "image_size": WORKLOADS[wl_name].get("image_size", FULL["image_size"]),
                       "note": ("the UNMODIFIED reference modules (runners.holycow step over the reference's plugins) on the "
                                "host CPU" if r["kind"] == "reference" else r.get("note", "CPU port")) +
                               f"; requested steps {args.steps} / warm-up {args.warmup} bounded to a ~200 s CPU budget"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


DRIVE_METRIC = "256x256 reenactment frames/sec (drive.py infer)"


def run_reference_drive(args):
    """`--impl reference --workload drive`: the reference's drive.py inner loop (drive.py:84-98, through its own embedder
    / generator modules) on the host cores.  Default batch 1 = BASELINE configs[0] (the reference hard-codes 1,
    drive.py:57); `--batch 64` = configs[3]'s batch on the CPU."""
    batch = args.batch or 1
    r = reference_cpu_run("drive", batch, max(args.steps, 3), 2 if args.warmup > 0 else 0, budget_s=120.0)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "no reference tree (baseline/_ref or /root/reference) for the "
                                                              "drive loop"}))
        return
    rate = r["frames_per_s"]
    sample = (f"{r['steps']} batch(es) of {r['batch']} frame(s) after {r['warmup']} warm-up, fp32, {r['cores']} torch threads "
              f"(reference pins 1, utils/utils.py:19: overridden), run-to-run spread {100 * r.get('spread', 0.0):.1f} %")
    print(json.dumps({
        "impl": "reference", "metric": DRIVE_METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": r["warmup"], "ms_per_step": r["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("configs[0]: drive.py single-frame inference, bs=1, CPU" if batch == 1 else
                                f"drive.py inference loop, bs={batch}, CPU"),
                   "per_gpu_batch": batch, "image_size": 256,
                   "note": "the UNMODIFIED reference modules (pose embedder + fine-tuned generator, drive.py:84-98) on the "
                           "host CPU; video encoding left out on both arms"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------------------------
def build_training(wl, device, batch):
    import importlib
    from b200lp import lib
    lib.require_device()
    runner = importlib.import_module("runners.holycow")
    vgg_dir = tempfile.mkdtemp(prefix="vgg_synth_")
    fabricate_vgg_files(vgg_dir)
    ns = make_namespace(wl, device, vgg_dir, batch)
    torch.manual_seed(123)
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(ns)
    D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(ns)
    E = importlib.import_module("embedders.unsupervised_pose_separate_embResNeXt_segmentation").Wrapper.get_net(ns)
    crits = [importlib.import_module(f"criterions.{c.strip()}").Wrapper.get_net(ns) for c in wl["criteria"].split(",")]
    if wl["finetune"]:
        e = torch.randn(1, FULL["embed_channels"], device=device)
        G.enable_finetuning({"embeds": e.clone()})
        D.enable_finetuning({"embeds": e.clone()})
        E.enable_finetuning()
        E.requires_grad_(False)      # not in optimizer_G when fine-tuning (runners/holycow.py:35-37)
    with torch.no_grad():            # a fresh net's spectral-norm vectors are random: converge them
        for net in (G, D):
            net.train()
            for m in net.modules():
                if hasattr(m, "inv_sigma"):
                    for _ in range(20):
                        m.inv_sigma()
        # AdaIN gains O(1) like a trained net
        proj = G.affine_params_projector.slot(2)
        off = 0
        for c in G.adain_sizes:
            proj.bias[off + c:off + 2 * c] += 1.0
            off += 2 * c
    tm = runner.TrainingModule(E, G, D, crits, [], {})
    tm.train()
    opt_G = runner.get_optimizer(E, G, ns)
    opt_D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_optimizer(D, ns)
    return runner, tm, opt_G, opt_D, ns


TRAFFIC_FILES = {"finetune": "r02_dram_traffic.json", "metatrain": "r02_dram_traffic_metatrain_final.json"}


def roofline_from_replay(replay_fn, work, peaks, workload="finetune"):
    """Per-family device time of ONE replay of the captured step (CUPTI kernel durations through torch.profiler: the
    kernels as they run back to back inside the graph, not event pairs around eager launches) against the algorithmic
    FLOPs / bytes the host booked while the step was captured (b200lp.kernels.WORK) -> the dominant
    family's achieved throughput.  Runs outside every timed region."""
    rows, total_ms = replay_kernel_times(replay_fn)
    table = family_table(rows, work)
    own_ms = sum(ms for n, (ms, c) in rows.items() if is_own_kernel(n))
    out = {"families": table, "replay": {"kernel_ms": round(total_ms, 3), "launches": sum(c for _, c in rows.values()),
                                         "libb200lp_share": round(own_ms / (total_ms or 1.0), 3)}}
    traffic = {}
    # ncu dram__bytes_{read,write}.sum per launch of one replayed step of THIS workload (tools/ncu_traffic.py);
    # a workload without a committed capture (metatrain512) reports traffic = null
    tj = ROOT / "profiles" / TRAFFIC_FILES.get(workload, "none")
    if tj.is_file():
        traffic = json.loads(tj.read_text())
    conv = table.get("conv_igemm_tf32")
    if conv and conv["tflops"]:
        tf32_peak = peaks.get("bf16_tflops_sustained", 1437.4) / 2.0
        ach = conv["tflops"]
        t = traffic.get("conv_igemm_tf32", {})
        out["roofline"] = {"kernel": "conv_igemm_tf32 family: conv_halo2 / conv_halo / conv_igemm <.., tf32> (tcgen05 fwd + dgrad)",
                           "bound": "tensor", "achieved": ach, "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                           "frac": round(ach / tf32_peak, 3), "traffic": t.get("dram_bytes_per_launch"),
                           "traffic_note": t.get("note"),
                           "peak_note": "TF32 dense = measured sustained bf16 cuBLAS TF/s / 2 (MEASURED_PEAKS.json); "
                                        "frac of the bf16 figure itself = %.3f" % (ach / (2 * tf32_peak)),
                           "timing": "CUPTI kernel durations of one graph replay", "avg_launch_ms":
                               round(conv["ms"] / conv["launches"], 4), "launches_per_step": conv["launches"],
                           "algorithmic_tflop_per_step": round(work.get("conv_igemm_tf32", {}).get("flops", 0.0) / 1e12, 3)}
        wg = table.get("conv_wgrad_tf32")
        if wg and wg["tflops"]:
            red = table.get("wgrad_reduce", {"ms": 0.0})
            out["roofline"]["wgrad"] = {"kernel": "conv_wgrad_tf32_kernel", "achieved": wg["tflops"],
                                        "frac": round(wg["tflops"] / tf32_peak, 3),
                                        "achieved_with_reduce": round(wg["tflops"] * wg["ms"] / (wg["ms"] + red["ms"]), 1),
                                        "unit": "TFLOP/s"}
        peak_h = peaks.get("hbm_gbs", 6580.3)
        hb = {}
        for fam in ("adain_relu", "in_stats", "adain_relu_bwd", "batchnorm", "l1", "elementwise"):
            f = table.get(fam)
            if f and f["gbs"]:
                hb[fam] = {"achieved": f["gbs"], "frac": round(f["gbs"] / peak_h, 3), "ms": f["ms"]}
        if hb:
            out["roofline"]["hbm"] = {"peak": peak_h, "unit": "GB/s", "bound": "hbm", "families": hb,
                                      "note": "algorithmic bytes (DESIGN.md §3) / CUPTI time of the family in one replay"}
    return out


def is_own_kernel(name):
    """Every kernel of libb200lp.so lives in namespace b200lp (csrc/*.cu); anything else is torch / cuDNN / cuBLAS / NCCL."""
    return "b200lp::" in name


def kernel_family(name):
    """Kernel name (CUPTI, demangled) -> family key used by the roofline table."""
    import re
    if not is_own_kernel(name):
        if "nccl" in name.lower():
            return "nccl"
        if name.startswith("Memcpy") or name.startswith("Memset"):
            return "memcpy/memset"
        return "torch/cudnn/cublas"
    base = name.split("b200lp::", 1)[1].split("(", 1)[0]
    m = re.match(r"(conv_igemm_kernel|conv_halo_kernel|conv_halo2_kernel)<(.*)>", base)
    if m:
        mode = m.group(2).split(",")[-1].replace("(int)", "").strip().rstrip(">")      # ncu prints `(int)1`
        return "conv_igemm_bf16x3" if mode.startswith("1") else "conv_igemm_tf32"
    base = base.split("<", 1)[0]
    table = (("splitk_epilogue", "conv_splitk_epilogue"), ("conv_wgrad_tf32", "conv_wgrad_tf32"), ("wgrad_", "wgrad_reduce"),
             ("sn_rank1", "wgrad_reduce"), ("sn_", "spectral_norm"), ("in_stats", "in_stats"), ("adain_bwd", "adain_relu_bwd"), ("adain_fused", "adain_relu"),
             ("adain_relu", "adain_relu"), ("l1_", "l1"), ("pack_conv_weight", "pack_conv_weight"),
             ("adam_ema", "optimizer"), ("ema_multi", "optimizer"), ("opt_tick", "optimizer"),
             ("gen_tail", "gen_tail"), ("conv3x3_c3", "c3_stem"), ("im2col3x3", "c3_stem"), ("col2im3x3", "c3_stem"),
             ("gconv", "resnext_grouped"), ("im2col7x7", "resnext"), ("maxpool3x3s2", "resnext"), ("subsample2", "resnext"),
             ("scatter_add2", "resnext"), ("avgpool_", "resnext"), ("sgemm_", "dense_small"), ("col_stats", "batchnorm"),
             ("bn_apply", "pose_encoder"), ("bn_relu6_avgpool", "pose_encoder"), ("bn_", "batchnorm"),
             ("pw_", "pose_encoder"), ("dw_", "pose_encoder"), ("mbv2", "pose_encoder"))
    for prefix, fam in table:
        if base.startswith(prefix):
            return fam
    return "elementwise"


def replay_kernel_times(fn):
    """{kernel name: (ms, launches)} of the device work fn() enqueues (one CUDA-graph replay of the step), measured by
    CUPTI through torch.profiler: per-kernel durations of the kernels as they run back to back inside the replay, not
    event pairs around eager launches.  Outside every timed region."""
    from torch.profiler import ProfilerActivity, profile
    multi = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if multi:
            # the profiler starts at a different moment on every rank: a barrier INSIDE the profiled region absorbs that
            # skew (its own NCCL kernel, dropped below), so the step's collectives do not spend their kernel time
            # waiting for a late peer
            torch.distributed.barrier()
            torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if multi:
        evs.sort(key=lambda e: e.time_range.start)
        first_nccl = next((k for k, e in enumerate(evs) if "nccl" in e.name.lower()), None)
        if first_nccl is not None:
            evs = evs[first_nccl + 1:]
    rows = {}
    for e in evs:
        ms, c = rows.get(e.name, (0.0, 0))
        rows[e.name] = (ms + e.device_time / 1e3, c + 1)
    return rows, sum(ms for ms, _ in rows.values())


WRAPPER_FAMILY = {"conv3x3_c3_fwd": "c3_stem", "c3_im2col": "c3_stem", "c3_col2im": "c3_stem", "conv3x3_c3_dgrad": "c3_stem",
                  "conv3x3_c3_wgrad": "c3_stem", "gen_tail_fwd": "gen_tail", "gen_tail_bwd_act": "gen_tail",
                  "gen_tail_bwd_data": "gen_tail"}


def family_table(rows, work=None):
    """Aggregate replay_kernel_times rows into families; `work` = {family: {flops, bytes}} algorithmic work of one step
    (b200lp.kernels.WORK, accumulated on the host while the step was captured)."""
    merged = {}
    for k, v in (work or {}).items():
        d = merged.setdefault(WRAPPER_FAMILY.get(k, k), {"flops": 0.0, "bytes": 0.0})
        d["flops"] += v.get("flops", 0.0)
        d["bytes"] += v.get("bytes", 0.0)
    work = merged
    fam = {}
    for name, (ms, c) in rows.items():
        d = fam.setdefault(kernel_family(name), {"ms": 0.0, "launches": 0})
        d["ms"] += ms
        d["launches"] += c
    total = sum(d["ms"] for d in fam.values()) or 1.0
    out = {}
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        w = (work or {}).get(k, {})
        out[k] = {"ms": round(v["ms"], 3), "share": round(v["ms"] / total, 3), "launches": v["launches"],
                  "tflops": round(w["flops"] / v["ms"] / 1e9, 1) if w.get("flops") else None,
                  "gbs": round(w["bytes"] / v["ms"] / 1e6, 1) if w.get("bytes") else None}
    return out


def drive_benchmark(device, batch=64, n_batches=6):
    """BASELINE configs[3]: drive.py inner loop (drive.py:84-98) batched — pinned host frames -> pose embedder ->
    generator forward (eval, fine-tuned) -> clamp/uint8 -> asynchronous D2H.  Returns frames/s (end to end)."""
    import importlib
    import drive
    wl = WORKLOADS["finetune"]
    ns = make_namespace(wl, device, "/nonexistent", batch)
    torch.manual_seed(7)
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(ns)
    E = importlib.import_module("embedders.unsupervised_pose_separate_embResNeXt_segmentation").Wrapper.get_net(ns)
    G.enable_finetuning({"embeds": torch.randn(1, FULL["embed_channels"], device=device)})
    E.enable_finetuning()
    with torch.no_grad():
        G.train()
        for m in G.modules():
            if hasattr(m, "inv_sigma"):
                for _ in range(20):
                    m.inv_sigma()
    G.eval(); E.eval()
    g = torch.Generator().manual_seed(5)
    frames = [torch.rand(batch, 3, 256, 256, generator=g) for _ in range(2)]
    with torch.no_grad():
        drive.render(E, G, (frames[i % 2] for i in range(2)), device, sink=None, with_driver=False)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = drive.render(E, G, (frames[i % 2] for i in range(n_batches)), device, sink=None, with_driver=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"value": round(n / dt, 1), "unit": "frames/s", "batch": batch, "frames": n,
            "what": "drive.py loop, bs=64: pinned H2D of driver frames, MobileNetV2 pose embed, generator forward "
                    "(bf16x3), clamp+uint8, async D2H of the frames"}


def run_drive_line(args):
    """`--workload drive`: BASELINE configs[3] as its own line (the default line carries the same number as
    `e2e.drive`).  End to end by construction: host frames in, host uint8 frames out, wall clock around the loop with a
    device synchronize on both sides; `--steps` = batches of 64 frames."""
    from b200lp import lib
    torch.cuda.set_device(0)
    l0 = lib.load().b200lp_launch_count()
    with torch.no_grad():
        d = drive_benchmark("cuda:0", batch=64, n_batches=max(args.steps, 2))
    launches = lib.load().b200lp_launch_count() - l0
    ms = 1e3 * d["batch"] / d["value"]
    frame_bytes = d["batch"] * 3 * 256 * 256
    cpu_baseline = None
    if not args.no_cpu_baseline:          # BASELINE configs[0]: the reference's loop at its own batch (1) on the host cores
        r = reference_cpu_run("drive", 1, 20, 2, budget_s=20.0)
        if r is not None:
            cpu_baseline = {"value": round(r["frames_per_s"], 3), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                            "sample": f"{r['steps']} frames at batch 1 (BASELINE configs[0]: drive.py's own batch) after "
                                      f"{r['warmup']} warm-up, fp32, {r['cores']} torch threads"}
    line = {
        "metric": DRIVE_METRIC, "value": d["value"], "unit": UNIT, "n_gpus": 1,
        "steps": max(args.steps, 2), "warmup": 2, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3 operands / f32 accumulate (f32 storage)", "data": "synthetic",
        "config": {"workload": "configs[3]: drive.py batched inference, bs=64, 256x256 driver frames, fine-tuned generator",
                   "per_gpu_batch": d["batch"], "image_size": 256, "what": d["what"],
                   "timing": "wall clock around the loop, device synchronised on both sides (host frames in, host frames out)"},
        "e2e": {"value": d["value"], "unit": UNIT, "h2d_bytes_per_step": 4 * frame_bytes, "d2h_bytes_per_step": frame_bytes},
        "gpu_launches": int(launches)}
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)


def shutdown_distributed(graphed_steps):
    """Leave a multi-rank run cleanly: the step's CUDA graph holds the communicator's captured all-reduces, and tearing
    the process group down UNDER a live graph hangs — so the graphs (and everything they keep alive) are destroyed first,
    the device is drained, and only then the process group is destroyed.  A watchdog thread ends the process if NCCL's
    teardown still does not return (everything has been printed by then)."""
    import gc
    torch.cuda.synchronize()
    torch.distributed.barrier()
    for g in graphed_steps:
        if g is not None:
            g.release()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(30.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    torch.distributed.destroy_process_group()
    timer.cancel()


class StepBench:
    """One workload set up for measurement: networks, optimizers, the captured step, host and device batches."""

    def __init__(self, name, device, batch, rank, use_graph=True):
        from utils import utils as U
        self.U, self.name, self.wl, self.B, self.device = U, name, WORKLOADS[name], batch, device
        wl = self.wl
        self.runner, self.tm, self.opt_G, self.opt_D, self.ns = build_training(wl, device, batch)
        self.tm.broadcast_parameters()
        self.n_batches = 4
        self.host = make_host_batches(self.n_batches, batch, wl["k_frames"], wl["num_labels"],
                                      s=wl.get("image_size", FULL["image_size"]), rank=rank)
        self.dev = [({k: v.to(device) for k, v in d.items()}, {k: v.to(device) for k, v in t.items()}) for d, t in self.host]
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host[0][0].values()) + \
            sum(v.numel() * v.element_size() for v in self.host[0][1].values())
        self.graphed, self.graph_note = None, "off (--no-graph)"
        if use_graph:
            try:
                self.graphed = self.runner.GraphedTrainStep(self.tm, self.opt_G, self.opt_D, wl["finetune"], self.dev[0][0],
                                                            self.dev[0][1])
                self.graph_note = "whole step captured once, replayed per batch"
            except Exception as err:      # keep measuring (eagerly) and say why
                self.graph_note = f"capture failed, eager launches: {type(err).__name__}: {str(err)[:160]}"
                torch.cuda.synchronize()
        self.n_losses = 0

    def step_eager(self, i):
        d, t = self.dev[i % self.n_batches]
        self.runner.train_step(self.tm, dict(d), dict(t), self.opt_G, self.opt_D, finetune=self.wl["finetune"])

    def step_resident(self, i):
        d, t = self.dev[i % self.n_batches]
        if self.graphed is not None:
            self.graphed(d, t)           # device->device copy into the graph's static inputs + replay
        else:
            self.runner.train_step(self.tm, dict(d), dict(t), self.opt_G, self.opt_D, finetune=self.wl["finetune"])

    def step_e2e(self, i):
        """The public API with HOST batches: pinned host -> device inside the timed region (the NEXT batch's upload is
        prefetched on the copy stream while this step's graph runs, like a prefetching data loader), replay, loss
        values read back to the host."""
        d, t = self.host[i % self.n_batches]
        if self.graphed is not None:
            nxt = self.host[(i + 1) % self.n_batches]
            _, lg, ld = self.graphed(d, t, prefetch_next=nxt)
        else:
            d, t = dict(d), dict(t)
            self.U.dict_to_device(d, self.device)
            self.U.dict_to_device(t, self.device)
            _, lg, ld = self.runner.train_step(self.tm, d, t, self.opt_G, self.opt_D, finetune=self.wl["finetune"])
        vals = torch.stack([v.detach().float().reshape(()) for v in list(lg.values()) + list(ld.values())])
        self.n_losses = vals.numel()
        return vals.cpu()           # device -> host read of the step's result (the runner's Meter does the same)

    def measure(self, steps, warmup, dist_on, with_e2e=True):
        for i in range(max(warmup, 3)):
            self.step_resident(i)
        ms = timed(self.step_resident, steps, dist_on)
        ms_e2e = None
        if with_e2e:
            for i in range(2):
                self.step_e2e(i)
            ms_e2e = timed(self.step_e2e, steps, dist_on)
        return ms, ms_e2e


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "drive":
        run_drive_line(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group(backend="nccl", init_method="env://", rank=rank, world_size=world)
    # N = 1: BASELINE configs[1] (the fine-tune step, the configuration the reference pins to ONE GPU, train.py:120-126);
    # N > 1: BASELINE configs[2] (meta-training: the only configuration the reference runs data-parallel,
    # train.py:98-109) — the N = 1 line carries the meta-training number too (`e2e.metatrain`) as the weak-scaling base.
    wl_name = args.workload or ("finetune" if world == 1 else "metatrain")
    wl = WORKLOADS[wl_name]
    B = args.batch or wl.get("batch", 8)
    S = wl.get("image_size", FULL["image_size"])
    from b200lp import lib
    sb = StepBench(wl_name, device, B, rank, use_graph=not args.no_graph)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for i in range(max(args.warmup, 3)):
        sb.step_resident(i)
    if sampler:
        sampler.start()
    l0 = lib.load().b200lp_launch_count()
    if args.timed_only:          # `ncu --profile-from-start off`: only the timed steps are profiled
        torch.cuda.profiler.start()
    ms = timed(sb.step_resident, args.steps, dist_on)
    if args.timed_only:
        torch.cuda.profiler.stop()
    launches = lib.load().b200lp_launch_count() - l0
    if sb.graphed is not None:      # replays do not pass through the library's host-side counter
        launches = sb.graphed.kernels_per_replay * args.steps
    clocks = sampler.stop() if sampler else None
    if args.timed_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": round(B * world * args.steps / (ms / 1e3), 2), "unit": UNIT,
                              "ms_per_step": round(ms / args.steps, 3), "gpu_launches": int(launches),
                              "note": "timed-only run (profiling aid, not a bench line)"}))
        if dist_on:
            shutdown_distributed([sb.graphed])
        return
    for i in range(2):
        sb.step_e2e(i)
    ms_e2e = timed(sb.step_e2e, args.steps, dist_on)

    frames = B * world * args.steps
    value = frames / (ms / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)

    extra = {}
    if not args.no_roofline and sb.graphed is not None:
        # every rank replays (the step contains the gradient all-reduces); rank 0 reports
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        prof = roofline_from_replay(lambda: sb.graphed(*sb.dev[0]), sb.graphed.work, peaks, workload=wl_name)
        if rank == 0:
            extra.update(prof)
            extra["peaks_source"] = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    also = {}
    if rank == 0 and world == 1:
        try:
            with torch.no_grad():
                also["drive"] = drive_benchmark(device)
        except Exception as err:
            also["drive"] = {"error": str(err)[:200]}
    if world == 1 and wl_name == "finetune" and args.workload is None:
        # the weak-scaling base of the N > 1 runs: the meta-training step on one GPU
        try:
            sb.graphed and sb.graphed.release()
            sb2 = StepBench("metatrain", device, B, rank, use_graph=not args.no_graph)
            k2 = max(5, args.steps // 2)
            ms2, ms2e = sb2.measure(k2, 3, False)
            also["metatrain"] = {"workload": WORKLOADS["metatrain"]["desc"], "value": round(B * k2 / (ms2 / 1e3), 2),
                                 "unit": UNIT, "ms_per_step": round(ms2 / k2, 3), "e2e_value": round(B * k2 / (ms2e / 1e3), 2),
                                 "e2e_ms_per_step": round(ms2e / k2, 3), "steps": k2, "h2d_bytes_per_step": sb2.h2d_bytes,
                                 "what": "BASELINE configs[2] at N = 1: the base of the N > 1 weak-scaling runs"}
            if not args.no_roofline and sb2.graphed is not None:
                rows, tot = replay_kernel_times(lambda: sb2.graphed(*sb2.dev[0]))
                own = sum(m for n_, (m, c) in rows.items() if is_own_kernel(n_))
                also["metatrain"]["libb200lp_share_of_kernel_time"] = round(own / (tot or 1.0), 3)
            sb2.graphed and sb2.graphed.release()
            try:    # context for the N > 1 lines of a back-to-back scaling run on this box (never an efficiency claim)
                (ROOT / ".bench_cache").mkdir(exist_ok=True)
                (ROOT / ".bench_cache" / "metatrain_n1.json").write_text(json.dumps(
                    {"value": also["metatrain"]["value"], "e2e_value": also["metatrain"]["e2e_value"],
                     "ms_per_step": also["metatrain"]["ms_per_step"], "unix_time": time.time()}))
            except OSError:
                pass
        except Exception as err:
            also["metatrain"] = {"error": f"{type(err).__name__}: {str(err)[:200]}"}
    if dist_on:
        torch.distributed.barrier()
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            # bounded sample: 1 warm-up (the first CPU step pays for thread pools and primitive caches) + up to 3 timed
            # steps at batch 2 within ~20 s of stepping
            r = cpu_step_rate(wl_name, 2, 3, 1, budget_s=20.0)
            cpu_baseline = {"value": round(r["frames_per_s"], 4), "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                            "sample": f"{r['steps']} step(s) after {r['warmup']} warm-up at batch {r['batch']} of the same workload "
                                      f"({r['s_per_step']:.1f} s/step), fp32, {r['cores']} torch threads; `--impl reference` runs the "
                                      f"full batch"}
        except Exception as err:   # the checker must never take the product number down with it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {err}"}

    if rank == 0:
        line = {"metric": METRIC if S == 256 else METRIC.replace("256x256", f"{S}x{S}"), "value": round(value, 2), "unit": UNIT,
                "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "tf32 operands / f32 accumulate (f32 storage)",
                "data": "synthetic",
                "config": {"workload": wl["desc"], "global_batch": B * world, "per_gpu_batch": B, "image_size": S,
                           "parallelism": f"dp{world}", "weights": "random init, spectral norm converged",
                           "l2": "per-step working set (activations, ~GBs) >> 126 MB L2; 4 distinct input batches cycled",
                           "cuda_graph": sb.graph_note,
                           "workload_by_n": "N = 1: configs[1] (reference pins fine-tuning to one GPU, train.py:120-126); "
                                            "N > 1: configs[2] meta-training (train.py:98-109); its N = 1 base is e2e.metatrain "
                                            "of the N = 1 line"},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": sb.h2d_bytes,
                        "d2h_bytes_per_step": 4 * sb.n_losses, "ms_per_step": round(ms_e2e / args.steps, 3),
                        "h2d": "pinned host batch -> device on a copy stream, prefetched one step ahead"},
                "gpu_launches": int(launches), "clocks": clocks}
        line["e2e"].update(also)
        if world > 1:
            try:
                base = json.loads((ROOT / ".bench_cache" / "metatrain_n1.json").read_text())
                if time.time() - base.get("unix_time", 0) < 6 * 3600:
                    line["config"]["n1_same_workload"] = {
                        "value": base["value"], "e2e_value": base["e2e_value"], "ms_per_step": base["ms_per_step"], "unit": UNIT,
                        "what": "configs[2] on ONE GPU of this box (written by the N = 1 run of this bench, e2e.metatrain): the "
                                "weak-scaling base of this line — the N = 1 line's own `value` is configs[1]"}
            except (OSError, ValueError, KeyError):
                pass
        line.update(extra)
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if dist_on:
        shutdown_distributed([sb.graphed])


if __name__ == "__main__":
    main()
