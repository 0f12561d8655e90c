"""Batch sweep on one GPU: trajectories/s, us per denoising iteration and achieved fraction of the
bf16 tensor peak for B in 1..4096 x {DDIM T in 2/10/100, DDPM T=100}, one JSON line per point (device-resident inputs,
CUDA events around whole plans, CUDA-graph replay).

    python scripts/sweep.py [--precisions bf16x3,bf16] [--batches 1,2,4,...] [--cases ddim:10,ddpm:100] [--budget-s 90] [--out file]

Stops cleanly when the time budget is spent (the points already measured are kept).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import autonomous_driving_with_diffusion_model_b200 as P  # noqa: E402
from autonomous_driving_with_diffusion_model_b200 import synthetic as W  # noqa: E402  (deterministic synthetic weights / inputs only)

FLOPS_PER_EVAL = 78_874_624  # SURVEY.md 8d, nominal 2*MAC per trajectory per denoiser evaluation (NO_GUIDANCE)

ap = argparse.ArgumentParser()
ap.add_argument("--precisions", default="bf16x3")
ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048,4096")
ap.add_argument("--cases", default="ddim:10,ddim:2,ddim:100,ddpm:100")
ap.add_argument("--budget-s", type=float, default=90.0)
ap.add_argument("--out", default=None)
a = ap.parse_args()

dev = torch.device("cuda:0")
peak = 1417.3
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["bf16_tflops_sustained"]
t_start = time.perf_counter()
sink = open(a.out, "w") if a.out else None


def emit(d):
    line = json.dumps(d)
    print(line, flush=True)
    if sink:
        sink.write(line + "\n")
        sink.flush()


cfg0 = P.load_cfg()
model = P.build_model(cfg0)
model.load_state_dict(W.make_state_dict("NO_GUIDANCE"))
model = model.to(dev).eval()
stream = torch.cuda.current_stream()
done = False
for case in a.cases.split(","):
    name, T = case.split(":")
    T = int(T)
    cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T))
    cls = P.GuidanceDDIMScheduler if name == "ddim" else P.GuidanceDDPMScheduler
    planner = P.DiffusionPlanner(model, cls(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
    for prec in a.precisions.split(","):
        model.set_precision(prec)
        for B in (int(b) for b in a.batches.split(",")):
            if time.perf_counter() - t_start > a.budget_s:
                done = True
                break
            inp = W.synth_inputs(B, 0, 1)
            x, f = inp["x"].to(dev), inp["feat"].to(dev)
            nz = torch.randn((T, B, 16, 7), device=dev) if name == "ddpm" else None   # timing only: any Gaussian draw
            run = lambda: planner.plan(x, f, noise=nz)  # noqa: E731
            run()
            run()
            torch.cuda.synchronize()
            n = 5 if B * T <= 25600 else 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                run()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            tf = FLOPS_PER_EVAL * B * T / (ms * 1e-3) / 1e12
            emit(dict(sched=name, T=T, precision=prec if B > 4 else "fp32 (small-batch GEMV path)", B=B, ms_per_plan=round(ms, 4),
                      us_per_iteration=round(ms * 1e3 / T, 2), traj_per_s=round(B / ms * 1e3, 1), tflops_nominal=round(tf, 2),
                      frac_of_bf16_sustained_peak=round(tf / peak, 5), launches_per_plan=planner.last_launch_count(), plans_timed=n))
        if done:
            break
    if done:
        emit(dict(note=f"time budget of {a.budget_s} s spent; remaining points not measured"))
        break
