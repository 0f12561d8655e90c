"""Developer helper for ncu: short plans (eager launches, no graph) of the headline workload shape: B trajectories, T iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
B = int(sys.argv[1]); prec = sys.argv[2]; T = int(sys.argv[3]) if len(sys.argv) > 3 else 3; n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = "cuda:0"
cfg = P.load_cfg(B200=dict(PRECISION=prec), EVAL=dict(SAMPLE_STEPS=T))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg, use_graph=False)
x = W.synth_inputs(B, 0, 1); xd, fd = x["x"].to(dev), x["feat"].to(dev)
for _ in range(n): y = pl.plan(xd, fd)
torch.cuda.synchronize(); print("ok", float(y.abs().mean()), "launches per plan", pl.last_launch_count())
