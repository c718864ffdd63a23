"""The N>1 host path on CPU: world_size-2 gloo processes, gradient buckets averaged across ranks
== single-process gradients on the concatenated batch (losses are batch means; the reference's apex Reducer computes
SUM / world, runners/holycow.py:241-250), and the rank-0 parameter broadcast."""
import os
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT / "latent-pose-reenactment_b200"))
    sys.path.insert(0, str(ROOT))
    import importlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    runner = importlib.import_module("runners.holycow")
    torch.manual_seed(100 + rank)                  # ranks start different: broadcast must fix that
    E, G, D = torch.nn.Linear(4, 4), torch.nn.Linear(4, 3), torch.nn.Linear(3, 1)
    tm = runner.TrainingModule(E, G, D, [], [], None)
    tm.broadcast_parameters()
    w_after_bcast = G.weight.detach().clone()
    opt_G = torch.optim.SGD(list(G.parameters()) + list(E.parameters()), lr=0.0)
    opt_D = torch.optim.SGD(D.parameters(), lr=0.0)
    bG, bD = tm.grad_buckets(opt_G, opt_D)
    x = torch.arange(16, dtype=torch.float32).view(4, 4) / 10.0
    xs = x[rank * 2:(rank + 1) * 2]                # shard of the global batch
    bG.zero(); bD.zero()
    D(G(E(xs))).pow(2).mean().backward()
    bG.all_reduce()
    bD.all_reduce()
    if rank == 0:
        torch.save({"gG": G.weight.grad.clone(), "gD": D.weight.grad.clone(), "w": w_after_bcast,
                    "state": {"E": E.state_dict(), "G": G.state_dict(), "D": D.state_dict()}}, out)
    else:
        torch.save({"w": w_after_bcast}, out + ".r1")
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29400 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0 = torch.load(out)
    r1 = torch.load(out + ".r1")
    assert torch.equal(r0["w"], r1["w"])                                   # broadcast from rank 0
    E, G, D = torch.nn.Linear(4, 4), torch.nn.Linear(4, 3), torch.nn.Linear(3, 1)
    E.load_state_dict(r0["state"]["E"]); G.load_state_dict(r0["state"]["G"]); D.load_state_dict(r0["state"]["D"])
    x = torch.arange(16, dtype=torch.float32).view(4, 4) / 10.0
    D(G(E(x))).pow(2).mean().backward()
    torch.testing.assert_close(r0["gG"], G.weight.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(r0["gD"], D.weight.grad, rtol=1e-5, atol=1e-7)
