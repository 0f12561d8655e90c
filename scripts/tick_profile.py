import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"; mode, T = "CLASSIFIER_GUIDANCE", 2
cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION="bf16x3"),
                 GUIDANCE=dict(USE_COND=mode, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]]))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode)); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
frames = [torch.randn(1, 3, 256, 900, device=dev) for _ in range(8)]
tg = torch.tensor([[0.1, 0.3]], device=dev)
for i in range(5): pl.generate_traj(frames[i % 8], tg)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(200):
    pl.generate_traj(frames[i % 8], tg); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
