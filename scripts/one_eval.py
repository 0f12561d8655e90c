"""Developer helper for ncu: a few denoiser evaluations at a given batch/precision (eager launches, no graph)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
B = int(sys.argv[1]); prec = sys.argv[2]; n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda:0"
cfg = P.load_cfg(B200=dict(PRECISION=prec))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
x = W.synth_inputs(B, 0, 1); xd, fd = x["x"].to(dev), x["feat"].to(dev)
t = torch.full((B,), 50, dtype=torch.long, device=dev)
for _ in range(n): y = m(xd, fd, t)
torch.cuda.synchronize(); print("ok", float(y.abs().mean()))
