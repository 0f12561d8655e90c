"""Developer (GPU): does ONE device run S concurrent sub-batch plans (S handles, S streams) faster than one plan of the whole batch?
An iteration at B = 64..256 is a latency-bound chain of launches that leaves SMs idle; independent chains interleave on them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = P.load_cfg(B200=dict(PRECISION="bf16x3"), EVAL=dict(SAMPLE_STEPS=100))
sd = W.make_state_dict("NO_GUIDANCE")
inp = W.synth_inputs(B, 0, 1)
xs, fs = inp["x"].to(dev), inp["feat"].to(dev)
ref = None
for S in (1, 2, 4):
    models, planners, streams = [], [], []
    for s in range(S):
        m = P.build_model(cfg); m.load_state_dict(sd); m = m.to(dev).eval()
        models.append(m); planners.append(P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)); streams.append(torch.cuda.Stream())
    n = B // S
    def run():
        outs = []
        for s in range(S):
            with torch.cuda.stream(streams[s]):
                outs.append(planners[s].plan(xs[s * n:(s + 1) * n], fs[s * n:(s + 1) * n]))
        torch.cuda.synchronize()
        return torch.cat(outs)
    for _ in range(2): y = run()
    t0 = time.perf_counter()
    for _ in range(3): y = run()
    dt = (time.perf_counter() - t0) / 3
    if ref is None: ref = y
    print(f"B={B} as {S} concurrent plan(s) of {n}: {dt * 1e3:.2f} ms -> {B / dt:.0f} traj/s; bitwise equal to the single plan: {bool(torch.equal(y, ref))} (max-abs difference {float((y - ref).abs().max()):.3e})", flush=True)
    del models, planners
