import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev="cuda:0"; prec=sys.argv[1] if len(sys.argv)>1 else "bf16x3"; B=int(sys.argv[2]) if len(sys.argv)>2 else 256
mode="NO_GUIDANCE"
cfg=P.load_cfg(B200=dict(PRECISION=prec), EVAL=dict(SAMPLE_STEPS=100))
m=P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode)); m=m.to(dev).eval()
s=P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)); pl=P.DiffusionPlanner(m,s,cfg)
x=W.synth_inputs(B,0,1); xd,fd=x["x"].to(dev),x["feat"].to(dev)
for _ in range(2): pl.plan(xd,fd)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(3): pl.plan(xd,fd)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/3
print(f"{prec} B={B} dbg={os.environ.get('B2P_TC_DBG','0')}: {dt*1e3:.2f} ms/plan -> {dt*1e4:.1f} us/step, {B/dt:.0f} traj/s", flush=True)
