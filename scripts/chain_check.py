"""Developer check (GPU): the chain kernels (csrc/chain64.cu) against the per-layer path and the CPU oracle, then timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
from oracle import plan as OP, unet as U

dev = "cuda:0"
def cfg(mode, T, prec):
    return P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=prec),
                      GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
quick = "--quick" in sys.argv
for mode in ("NO_GUIDANCE", "FREE_GUIDANCE", "CLASSIFIER_GUIDANCE"):
    sd = W.make_state_dict(mode, seed=0)
    for prec in ("bf16x3", "bf16"):
        m = P.build_model(cfg(mode, 10, prec)); m.load_state_dict(sd); m = m.to(dev).eval()
        for B in (5, 8, 13, 40):
            inp = W.synth_inputs(B, 0, 300 + B)
            t = torch.full((B,), 41, dtype=torch.long)
            cond = inp["target"] if mode == "FREE_GUIDANCE" else None
            ref = U.unet_forward(sd, inp["x"], inp["feat"], t, cond, mode, **({"return_action_and_time_only": True} if mode == "CLASSIFIER_GUIDANCE" else {}))
            ref = ref[0] if isinstance(ref, tuple) else ref
            res = {}
            for ch in (True, False):
                m.set_chain(ch)
                out = m(inp["x"].to(dev), inp["feat"].to(dev), t.to(dev), cond=None if cond is None else cond.to(dev),
                        **({"return_action_and_time_only": True} if mode == "CLASSIFIER_GUIDANCE" else {}))
                out = out[0] if isinstance(out, tuple) else out
                torch.cuda.synchronize()
                res[ch] = float((out.cpu() - ref).abs().max())
            print(f"fwd {mode} {prec} B={B}: err chain {res[True]:.3e}  per-layer {res[False]:.3e}  launches {m.last_launch_count()}", flush=True)
        if mode == "NO_GUIDANCE":
            for kind, T in (("guidance_ddim", 6), ("guidance_ddpm", 5)):
                cls = {"guidance_ddim": P.GuidanceDDIMScheduler, "guidance_ddpm": P.GuidanceDDPMScheduler}[kind]
                c = cfg(mode, T, prec)
                B = 21
                inp = W.synth_inputs(B, T, 7)
                nz = inp["noise"] if kind.endswith("ddpm") else None
                ref = OP.plan(sd, mode, kind, inp["x"], inp["feat"], T, noise=nz, postprocess=False)
                for ch in (True, False):
                    m.set_chain(ch)
                    for g in (True, False):
                        pl = P.DiffusionPlanner(m, cls(cfg=c, **P.scheduler_kwargs(c)), c, use_graph=g)
                        out = pl.plan(inp["x"].to(dev), inp["feat"].to(dev), noise=None if nz is None else nz.to(dev), postprocess=False)
                        torch.cuda.synchronize()
                        print(f"plan {kind} T={T} {prec} chain={ch} graph={g}: err {float((out.cpu() - ref).abs().max()):.3e} launches {pl.last_launch_count()}", flush=True)
        del m
if not quick:
    mode, prec = "NO_GUIDANCE", "bf16x3"
    sd = W.make_state_dict(mode, seed=0)
    m = P.build_model(cfg(mode, 100, prec)); m.load_state_dict(sd); m = m.to(dev).eval()
    c = cfg(mode, 100, prec)
    for B in (8, 64, 256, 1024):
        inp = W.synth_inputs(B, 0, 1)
        x, f = inp["x"].to(dev), inp["feat"].to(dev)
        for ch in (True, False):
            m.set_chain(ch)
            pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=c, **P.scheduler_kwargs(c)), c)
            for _ in range(3): pl.plan(x, f)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): pl.plan(x, f)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"time B={B} chain={ch}: {ms:.3f} ms/plan = {ms * 10:.1f} us/iteration, {B / ms * 1e3:.0f} traj/s, launches {pl.last_launch_count()}", flush=True)
