"""world_size-2 gloo test of the N>1 host logic (CPU): contiguous batch sharding, per-rank planning of independent
requests, gather == unsharded result, and the max-over-ranks timing reduction bench.py uses.  The per-rank planner is the
CPU oracle here (no GPU in this container); on the GPU box the same sharding wraps DiffusionPlanner.plan."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import autonomous_driving_with_diffusion_model_b200 as P
from oracle import plan as OP
from oracle import weights as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B, T = 5, 2   # ragged split: ranks get 3 and 2 trajectories
    inp = W.synth_inputs(B, T, seed=9)
    sd = W.make_state_dict("FREE_GUIDANCE", with_perception=False)
    sh = lambda t, d=0: P.shard(t, rank, world, d)  # noqa: E731
    out = OP.plan(sd, "FREE_GUIDANCE", "guidance_ddpm", sh(inp["x"]), sh(inp["feat"]), T, target=sh(inp["target"]), noise=sh(inp["noise"], 1))
    sizes = [hi - lo for lo, hi in P.shard_bounds(B, world)]
    padded = torch.zeros(max(sizes), 16, 7)
    padded[: out.shape[0]] = out
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)            # result gather on the host side (not part of the sampling path)
    bufs = [g[:n] for g, n in zip(gathered, sizes)]
    t = torch.tensor([1.0 + rank])          # pretend per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = OP.plan(sd, "FREE_GUIDANCE", "guidance_ddpm", inp["x"], inp["feat"], T, target=inp["target"], noise=inp["noise"])
        q.put((float((torch.cat(bufs, 0) - full).abs().max()), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_plan_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-3     # metres on x,y (x23.315); no cross-sample operation anywhere (CPU BLAS blocking varies with batch; the CUDA path is bitwise, see GPU tests)
    assert tmax == 2.0     # max over ranks
