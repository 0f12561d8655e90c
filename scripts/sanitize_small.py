"""Developer helper for compute-sanitizer: tiny eager plans that touch every kernel family once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for mode, prec, B, T in (("CLASSIFIER_GUIDANCE", "bf16x3", 1, 1), ("CLASSIFIER_GUIDANCE", "bf16x3", 5, 1), ("FREE_GUIDANCE", "fp32", 3, 1), ("NO_GUIDANCE", "bf16", 130, 1)):
    if which != "all" and which != mode:
        continue
    cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T * 50), B200=dict(PRECISION=prec),
                     GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0, LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
    m = P.build_model(cfg); m.load_state_dict(W.make_state_dict(mode, with_perception=False), strict=False); m = m.to(dev).eval()
    pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg, num_inference_steps=2, use_graph=False)
    inp = W.synth_inputs(B, 2, 3)
    y = pl.plan(inp["x"].to(dev), inp["feat"].to(dev), target=inp["target"].to(dev) if mode != "NO_GUIDANCE" else None)
    torch.cuda.synchronize()
    print(mode, prec, B, "ok", float(y.abs().mean()), flush=True)
