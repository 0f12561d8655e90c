"""Developer helper (GPU): host-side overhead of DiffusionPlanner.plan at batch 1 (wall clock vs CUDA events, cProfile)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
for mode, T in (("NO_GUIDANCE", 2), ("NO_GUIDANCE", 10), ("CLASSIFIER_GUIDANCE", 2)):
    cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION="bf16x3"),
                     GUIDANCE=dict(USE_COND=mode, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
    m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode, with_perception=False), strict=False); m = m.to(dev).eval()
    pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
    inp = W.synth_inputs(1, T, 3)
    x, f = inp["x"].to(dev), inp["feat"].to(dev)
    tg = inp["target"].to(dev) if mode != "NO_GUIDANCE" else None
    for _ in range(5): pl.plan(x, f, target=tg)
    torch.cuda.synchronize()
    wall, evt, host = [], [], []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e0.record(); pl.plan(x, f, target=tg); e1.record()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        wall.append((t2 - t0) * 1e6); host.append((t1 - t0) * 1e6); evt.append(e0.elapsed_time(e1) * 1e3)
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"{mode} T={T} B=1: wall {med(wall):.0f} us, host-side call {med(host):.0f} us, CUDA events {med(evt):.0f} us, launches {pl.last_launch_count()}")
    if T == 2 and mode == "NO_GUIDANCE":
        pr = cProfile.Profile(); pr.enable()
        for _ in range(200): pl.plan(x, f, target=tg)
        pr.disable(); torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
