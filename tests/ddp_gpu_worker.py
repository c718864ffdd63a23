"""Worker of tests/test_ddp_gpu.py (one process per GPU under torch.distributed.run): the graph-replayed training step
on this rank's SHARD of a batch — gradients averaged by the in-graph NCCL all-reduce over the flat buckets — against the
eager single-shot step on the CONCATENATED batch, same initial weights.  Criteria = the batch-mean ones (adversarial,
featmat, perceptual, idt_embed, dis_embed): their global gradient is the mean of the per-rank gradients, which is what
the reference's apex Reducer computes (SUM / world, runners/holycow.py:241-250).  (dice is a ratio of batch SUMS — not
separable — and is left out here, as it would be in the reference's own data-parallel run vs a large-batch run.)"""
import importlib
import json
import os
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT / "latent-pose-reenactment_b200"), str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from helpers import StubEmbedder, make_args, write_vgg_files
    from oracle import synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    torch.distributed.init_process_group("nccl", init_method="env://", rank=rank, world_size=world)
    cfg = synth.SMALL_CFG
    per_rank = 2
    data, target, emb = synth.make_inputs(cfg, batch=per_rank * world, seed=4)
    runner = importlib.import_module("runners.holycow")
    names = ("adversarial", "featmat", "idt_embed", "perceptual", "dis_embed")
    results = {}
    for mode in ("full_eager", "shard_graph"):
        with tempfile.TemporaryDirectory() as vgg_dir:
            write_vgg_files(vgg_dir)
            args = make_args(cfg, device=dev, vgg_weights_dir=vgg_dir)
            crit = [importlib.import_module(f"criterions.{n}").Wrapper.get_net(args) for n in names]
        G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
        D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
        G.load_state_dict(synth.generator_state_dict(cfg, seed=1))
        D.load_state_dict(synth.discriminator_state_dict(cfg, seed=2))
        if mode == "full_eager":
            sl = slice(0, per_rank * world)
        else:
            sl = slice(rank * per_rank, (rank + 1) * per_rank)
        d = {k: v[sl].to(dev) for k, v in data.items()}
        t = {k: v[sl].to(dev) for k, v in target.items()}
        E = StubEmbedder({k: v[sl].to(dev) for k, v in emb.items()}).to(dev)
        tm = runner.TrainingModule(E, G, D, crit, [], {})
        tm.train()
        tm.broadcast_parameters()
        opt_G = runner.get_optimizer(E, G, args)
        opt_D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_optimizer(D, args)
        if mode == "full_eager":
            # every rank runs the identical full-batch step; the bucket all-reduce then averages identical gradients
            _, lg, ld = runner.train_step(tm, d, t, opt_G, opt_D, finetune=False)
            step = None
        else:
            step = runner.GraphedTrainStep(tm, opt_G, opt_D, False, d, t, warmup=2)
            _, lg, ld = step(d, t)
        torch.cuda.synchronize()
        bG, bD = tm.grad_buckets(opt_G, opt_D)
        results[mode] = dict(gG=bG.flat.detach().clone(), gD=bD.flat.detach().clone(),
                             wG=torch.cat([p.detach().flatten() for p in G.parameters()]),
                             wD=torch.cat([p.detach().flatten() for p in D.parameters()]),
                             losses={k: float(v) for k, v in {**lg, **ld}.items()})
        if step is not None:
            step.release()
    a, b = results["full_eager"], results["shard_graph"]
    rep = {"rank": rank}
    for k in ("gG", "gD", "wG", "wD"):
        ref = a[k]
        rep[k] = float((b[k] - ref).abs().max() / (ref.abs().max() + 1e-30))
    # the loss VALUES of a shard differ from the full batch's; their mean over ranks must match
    for k, v in b["losses"].items():
        tns = torch.tensor([v], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tns)
        rep["loss." + k] = abs(float(tns) / world - a["losses"][k]) / (abs(a["losses"][k]) + 1e-30)
    # and every rank must hold the same weights after the step
    w = torch.cat([b["wG"], b["wD"]])
    w0 = w.clone()
    torch.distributed.broadcast(w0, src=0)
    rep["rank_weight_divergence"] = float((w - w0).abs().max())
    out = Path(os.environ["DDP_TEST_OUT"])
    (out / f"rank{rank}.json").write_text(json.dumps(rep))
    torch.cuda.synchronize()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
