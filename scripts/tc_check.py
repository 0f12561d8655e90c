"""Developer check (GPU): denoiser forward and short plans in every precision mode vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from oracle import unet as U, weights as W, plan as OP

dev = "cuda:0"
modes = sys.argv[1:] or ["NO_GUIDANCE"]
for mode in modes:
    sd = W.make_state_dict(mode, seed=0)
    for prec in ("fp32", "bf16x3", "bf16"):
        cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), B200=dict(PRECISION=prec), EVAL=dict(SAMPLE_STEPS=10),
                         GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
        model = P.build_model(cfg); model.load_state_dict(sd); model = model.to(dev).eval()
        for B in (1, 3, 37):
            inp = W.synth_inputs(B, 0, 100 + B)
            t = torch.tensor([(17 * i + 3) % 100 for i in range(B)])
            cond = inp["target"] if mode == "FREE_GUIDANCE" else None
            ref = U.unet_forward(sd, inp["x"], inp["feat"], t, cond, mode)
            out = model(inp["x"].to(dev), inp["feat"].to(dev), t.to(dev), cond=None if cond is None else cond.to(dev))
            torch.cuda.synchronize()
            d = (out.cpu() - ref).abs()
            print(f"{mode} {prec} forward B={B}: max {float(d.max()):.3e} mean {float(d.mean()):.3e} finite={bool(torch.isfinite(out).all())}", flush=True)
        sched = P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg))
        planner = P.DiffusionPlanner(model, sched, cfg)
        B, T = 8, 10
        inp = W.synth_inputs(B, T, 3)
        tg = inp["target"] if mode != "NO_GUIDANCE" else None
        ref = OP.plan(sd, mode, "guidance_ddim", inp["x"], inp["feat"], T, target=tg, postprocess=False)
        out = planner.plan(inp["x"].to(dev), inp["feat"].to(dev), target=None if tg is None else tg.to(dev), postprocess=False)
        d = (out.cpu() - ref).abs()
        print(f"{mode} {prec} plan DDIM-{T} B={B}: max {float(d.max()):.3e} mean {float(d.mean()):.3e}", flush=True)
        if mode == "NO_GUIDANCE":
            x = W.synth_inputs(256, 0, 1)
            xd, fd = x["x"].to(dev), x["feat"].to(dev)
            p100 = P.DiffusionPlanner(model, sched, cfg, num_inference_steps=100)
            for _ in range(2): p100.plan(xd, fd)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): p100.plan(xd, fd)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            print(f"   B=256 DDIM-100: {dt*1e3:.1f} ms/plan  {256/dt:.0f} traj/s  launches {p100.last_launch_count()}", flush=True)
