"""Developer check of the EXPERIMENTAL one-cluster-per-trajectory denoiser kernel (csrc/unet_cluster.cu, B2P_CLUSTER_EVAL=1):
one forward and one DDIM plan at batch 1 against the default small-batch (GEMV) path, plus plan latency of both.
Prints one JSON line.  Usage: python scripts/cluster_eval_check.py [T] [B]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import autonomous_driving_with_diffusion_model_b200 as P  # noqa: E402
from oracle import weights as W  # noqa: E402

MODE = os.environ.get("B2P_CLUSTER_CHECK_MODE", "1")   # "17": the (untested) st.async exchange variant
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
out = {"T": T, "B": B}
sd = W.make_state_dict("NO_GUIDANCE")
inp = W.synth_inputs(B, 0, 1)
x, f = inp["x"].to(dev), inp["feat"].to(dev)
t = torch.full((B,), 50, dtype=torch.long, device=dev)


def make(env):
    os.environ["B2P_CLUSTER_EVAL"] = env      # read by b2p_create
    cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T))
    m = P.build_model(cfg)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    return m, P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)


def latency(pl):
    for _ in range(3):
        pl.plan(x, f)
    ts = []
    for _ in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pl.plan(x, f)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


if "--bisect-only" in sys.argv:   # timing bisect alone (results are wrong in these modes), one JSON line per mode
    for mode in [a for a in sys.argv[3:] if a.isdigit()]:
        _, pm = make(mode)
        print(json.dumps({"mode": int(mode), "T": T, "B": B, "plan_p50_ms": latency(pm)}), flush=True)
    sys.exit(0)

m0, p0 = make("0")
y0 = m0(x, f, t)
plan0 = p0.plan(x, f, postprocess=False)
torch.cuda.synchronize()
out["default_launches_per_eval"] = m0.last_launch_count()
out["default_plan_p50_ms"] = latency(p0)
try:
    m1, p1 = make(MODE)
    y1 = m1(x, f, t)
    torch.cuda.synchronize()
    out["cluster_launches_per_eval"] = m1.last_launch_count()
    out["forward_max_abs_diff"] = float((y1 - y0).abs().max())
    plan1 = p1.plan(x, f, postprocess=False)
    torch.cuda.synchronize()
    out["plan_max_abs_diff"] = float((plan1 - plan0).abs().max())
    out["cluster_plan_p50_ms"] = latency(p1)
    for mode in [a for a in sys.argv[3:] if a.isdigit()]:   # timing bisect (results are wrong in these modes): 3 no dot, 5 no exchange, 9 no weight stream, 15 skeleton
        _, pm = make(mode)
        out[f"bisect_{mode}_plan_p50_ms"] = latency(pm)
except Exception as exc:  # report, do not hide
    out["error"] = repr(exc)
print(json.dumps(out), flush=True)
