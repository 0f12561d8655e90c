"""Developer helper (GPU): closed-loop tick latency at batch 1 — a NEW camera frame every tick, so the encoder runs once
per plan (interact.py:170-176 -> generate_traj)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
for mode, T in (("NO_GUIDANCE", 100), ("FREE_GUIDANCE", 10), ("CLASSIFIER_GUIDANCE", 2)):
    cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION="bf16x3"),
                     GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
    m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode)); m = m.to(dev).eval()
    pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
    frames = [torch.randn(1, 3, 256, 900, device=dev) for _ in range(8)]
    tg = torch.tensor([[0.1, 0.3]], device=dev) if mode != "NO_GUIDANCE" else None
    for i in range(5): pl.generate_traj(frames[i % 8], tg)
    torch.cuda.synchronize()
    tick, enc = [], []
    for i in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pl.generate_traj(frames[i % 8], tg)
        torch.cuda.synchronize(); tick.append((time.perf_counter() - t0) * 1e3)
    for i in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad(): m.perception(frames[i % 8])
        torch.cuda.synchronize(); enc.append((time.perf_counter() - t0) * 1e3)
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"{mode} T={T}: tick p50 {med(tick):.3f} ms (encoder alone {med(enc):.3f} ms)", flush=True)
