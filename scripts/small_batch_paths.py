"""Developer (GPU): DDIM-100 plan latency at small batch on the GEMV program (exact fp32) vs the tiled tensor-core kernels (bf16x3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
cfg = P.load_cfg(B200=dict(PRECISION="bf16x3"), EVAL=dict(SAMPLE_STEPS=100))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
for B in (1, 2, 3, 4, 6, 8):
    x = W.synth_inputs(B, 0, 1); xd, fd = x["x"].to(dev), x["feat"].to(dev)
    row = []
    for limit in (8, 0):
        m.set_small_batch_max(limit)
        for _ in range(2): pl.plan(xd, fd)
        torch.cuda.synchronize(); ts = []
        for _ in range(7):
            t0 = time.perf_counter(); pl.plan(xd, fd); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        row.append(sorted(ts)[3] * 1e3)
    print(f"B={B}: GEMV program {row[0]:.2f} ms, tensor-core tiles {row[1]:.2f} ms per DDIM-100 plan (p50 of 7)", flush=True)
