"""Developer helper (GPU): per-layer stage clocks of the EXPERIMENTAL cluster evaluation kernel (csrc/unet_cluster.cu).
Needs the trace build:  B2P_TRACE_BUILD=1 python -m autonomous_driving_with_diffusion_model_b200.build
Usage: B2P_TRACE_BUILD=1 python scripts/cluster_trace.py
Stages of CTA 0 (SM clocks): layer start -> row tables + constants issued -> dot products (weight ring) -> K slices combined ->
DSMEM exchange + cluster barrier -> GroupNorm statistics -> epilogue."""
import ctypes as C
import os
import sys

os.environ["B2P_TRACE_BUILD"] = "1"
os.environ["B2P_CLUSTER_EVAL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import autonomous_driving_with_diffusion_model_b200 as P  # noqa: E402
from autonomous_driving_with_diffusion_model_b200 import _lib  # noqa: E402
from oracle import weights as W  # noqa: E402

dev = "cuda:0"
cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=10))
m = P.build_model(cfg)
m.load_state_dict(W.make_state_dict("NO_GUIDANCE"))
m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
inp = W.synth_inputs(1, 0, 1)
x, f = inp["x"].to(dev), inp["feat"].to(dev)
for _ in range(4):
    pl.plan(x, f)
torch.cuda.synchronize()
lib = _lib.load()
lib.b2p_debug_uc_trace.restype = C.c_int
MAXOPS = 48
buf = np.zeros((MAXOPS + 1) * 8, dtype=np.uint64)
assert lib.b2p_debug_uc_trace(buf.ctypes.data_as(C.c_void_p)) == 0
tr = buf.reshape(MAXOPS + 1, 8).astype(np.int64)
names = ["tables", "dot", "combine", "exchange", "stats", "epilogue"]
print(f"prologue {tr[MAXOPS, 1] - tr[MAXOPS, 0]} cycles")
print("layer " + " ".join(f"{n:>9s}" for n in names) + "     total")
tot = np.zeros(6, dtype=np.int64)
n_ops = int((tr[:MAXOPS, 0] > 0).sum())
for i in range(n_ops):
    d = [tr[i, k + 1] - tr[i, k] for k in range(6)]
    tot += np.array(d)
    print(f"{i:5d} " + " ".join(f"{v:9d}" for v in d) + f" {tr[i, 6] - tr[i, 0]:9d}")
print("sum   " + " ".join(f"{v:9d}" for v in tot) + f" {int(tot.sum()):9d}")
print("us    " + " ".join(f"{v / 1965:9.1f}" for v in tot) + f" {tot.sum() / 1965:9.1f}   (at 1.965 GHz; whole kernel {(tr[n_ops - 1, 6] - tr[MAXOPS, 0]) / 1965:.1f} us)")
