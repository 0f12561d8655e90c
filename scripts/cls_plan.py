"""Developer helper for ncu: one classifier-guidance plan (config 4a, T=2) at a given batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
B = int(sys.argv[1]); prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
dev = "cuda:0"; mode = "CLASSIFIER_GUIDANCE"
cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=2), B200=dict(PRECISION=prec),
                 GUIDANCE=dict(USE_COND=mode, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]]))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode)); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg, use_graph=False)
inp = W.synth_inputs(B, 2, 3)
for _ in range(2): y = pl.plan(inp["x"].to(dev), inp["feat"].to(dev), target=inp["target"].to(dev))
torch.cuda.synchronize(); print("ok", float(y.abs().mean()))
