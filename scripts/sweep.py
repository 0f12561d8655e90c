"""Batch sweep: trajectories/s and achieved TFLOP/s vs batch for each precision (one GPU)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from oracle import weights as W
dev = "cuda:0"
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16x3", "bf16"]
batches = [int(b) for b in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 16, 64, 256, 1024, 4096]
cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
s = P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)); pl = P.DiffusionPlanner(m, s, cfg)
for prec in precs:
    m.set_precision(prec)
    for B in batches:
        x = W.synth_inputs(B, 0, 1); xd, fd = x["x"].to(dev), x["feat"].to(dev)
        for _ in range(2): pl.plan(xd, fd)
        torch.cuda.synchronize(); n = 5 if B <= 1024 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): pl.plan(xd, fd)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 78.874624e6 * B * T / (ms * 1e-3) / 1e12
        print(json.dumps(dict(precision=prec, B=B, T=T, ms_per_plan=round(ms, 3), us_per_step=round(ms * 1e3 / T, 1), traj_per_s=round(B / ms * 1e3), tflops_nominal=round(tf, 1))), flush=True)
