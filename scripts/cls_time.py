"""Developer (GPU): classifier-guidance plan (config 4a, DDIM-2, scale 15) timing at a given batch, graph replay."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"; mode = "CLASSIFIER_GUIDANCE"
cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=2), B200=dict(PRECISION="bf16x3"),
                 GUIDANCE=dict(USE_COND=mode, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]]))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode)); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
for B in (1, 16, 256, 1024):
    inp = W.synth_inputs(B, 2, 3)
    x, f, tg = inp["x"].to(dev), inp["feat"].to(dev), inp["target"].to(dev)
    for _ in range(3): pl.plan(x, f, target=tg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pl.plan(x, f, target=tg)
    e1.record(); torch.cuda.synchronize()
    g = e0.elapsed_time(e1) / 20
    e0.record()
    for _ in range(20): pl.plan(x, f, target=None)
    e1.record(); torch.cuda.synchronize()
    n = e0.elapsed_time(e1) / 20
    print(f"classifier DDIM-2 B={B}: guided {g:.3f} ms/plan, unguided (state_pred forward only) {n:.3f} ms/plan", flush=True)
