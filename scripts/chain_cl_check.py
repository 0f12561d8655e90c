"""Developer (GPU): plan outputs of a few batch sizes / modes as one digest per case.  Run once with B2P_CHAIN_CL=1 and once with B2P_CHAIN_CL=2:
the two-CTA form of the chain kernel must give bit-identical trajectories."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W

dev = "cuda:0"
for prec in ("bf16x3", "bf16"):
    for mode, sched in (("NO_GUIDANCE", "ddim"), ("NO_GUIDANCE", "ddpm"), ("FREE_GUIDANCE", "ddim"), ("CLASSIFIER_GUIDANCE", "ddim")):
        for B in (5, 8, 37, 256, 300):
            cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=6), B200=dict(PRECISION=prec),
                             GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
            m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode, seed=1)); m = m.to(dev).eval()
            S = P.GuidanceDDIMScheduler if sched == "ddim" else P.GuidanceDDPMScheduler
            pl = P.DiffusionPlanner(m, S(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
            x = W.synth_inputs(B, 0, 1)
            kw = {}
            if mode != "NO_GUIDANCE":
                kw["target"] = x["target"].to(dev)
            if sched == "ddpm":
                kw["noise"] = W.synth_inputs(B, 6, 1)["noise"].to(dev)
            y = pl.plan(x["x"].to(dev), x["feat"].to(dev), **kw)
            torch.cuda.synchronize()
            y = y.float().cpu().contiguous()
            print(prec, mode, sched, B, hashlib.sha1(y.numpy().tobytes()).hexdigest()[:16], f"{float(y.abs().mean()):.6f}", bool(torch.isfinite(y).all()), flush=True)
