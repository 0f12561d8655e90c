"""Developer (GPU): large-batch regime (B=4096, DDIM-10) throughput and fraction of the bf16 tensor peak; env switches apply."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = 10
cfg = P.load_cfg(B200=dict(PRECISION=prec), EVAL=dict(SAMPLE_STEPS=T))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
g = torch.Generator(device=dev).manual_seed(5)
x, f = torch.randn(B, 16, 7, device=dev, generator=g), torch.randn(B, 64, device=dev, generator=g)
for _ in range(2): pl.plan(x, f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): pl.plan(x, f)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
tf = 78_874_624 * B * T / (ms * 1e-3) / 1e12
env = {k: v for k, v in os.environ.items() if k.startswith("B2P_")}
print(json.dumps({"prec": prec, "B": B, "ms_per_plan": round(ms, 3), "us_per_iteration": round(ms * 100, 1), "traj_per_s": round(B / ms * 1e3), "nominal_tflops": round(tf, 1),
                  "frac_of_1417": round(tf / 1417.3, 4), "env": env}), flush=True)
