"""Developer helper (GPU): per-stage clocks of the tcgen05 conv kernel inside a real plan.
Needs the trace build:  B2P_TRACE_BUILD=1 python -m autonomous_driving_with_diffusion_model_b200.build
Usage: B2P_TRACE_BUILD=1 python scripts/tc_trace.py [precision] [B]
Stages (CTA (0,0) of every launch, SM clocks): start -> dependency wait done (producer) -> first operands landed (MMA
thread) -> last MMA committed -> accumulators visible to the epilogue -> GroupNorm done -> stores issued."""
import sys, os, ctypes as C
os.environ["B2P_TRACE_BUILD"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import _lib
from autonomous_driving_with_diffusion_model_b200 import synthetic as W

dev = "cuda:0"; prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
T = 20
cfg = P.load_cfg(B200=dict(PRECISION=prec, SMALL_BATCH_MAX=0), EVAL=dict(SAMPLE_STEPS=T))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE")); m = m.to(dev).eval()
s = P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)); pl = P.DiffusionPlanner(m, s, cfg)
x = W.synth_inputs(B, 0, 1); xd, fd = x["x"].to(dev), x["feat"].to(dev)
pl.plan(xd, fd); torch.cuda.synchronize()          # capture (launch ids 0..n-1 are assigned at capture time)
lib = _lib.load()
lib.b2p_debug_tc_trace.restype = C.c_int
n1 = C.c_int(0)
buf = np.zeros(8192 * 16, dtype=np.uint64)
for _ in range(3): pl.plan(xd, fd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pl.plan(xd, fd); e1.record(); torch.cuda.synchronize()
assert lib.b2p_debug_tc_trace(buf.ctypes.data_as(C.c_void_p), C.byref(n1)) == 0
n = n1.value
tr = buf.reshape(8192, 16)[:n].astype(np.int64)
per_step = n // T
print(f"{prec} B={B}: plan {e0.elapsed_time(e1) * 1e3 / T:.1f} us/step, {per_step} tcgen05 launches per step")
names = ["start->wait done", "wait->operands", "MMA issue", "MMA drain", "taps+GN", "add+store"]
step = T // 2
rows = tr[step * per_step:(step + 1) * per_step]
tot = np.zeros(6)
tails, rels = [], []
print("layer  " + "  ".join(f"{n_:>16s}" for n_ in names) + "   next_kernel_ns   tail_ns(last CTA end - CTA0 end)   release_ns(next wait return - last CTA end)")
for i, r in enumerate(rows):
    d = [r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5]]
    nxt = rows[i + 1][7] - r[7] if i + 1 < len(rows) else 0
    tot += np.array(d)
    tail = r[8] - r[7]
    rel = rows[i + 1][9] - r[8] if i + 1 < len(rows) else 0
    tails.append(tail); rels.append(rel)
    print(f"{i:4d}   " + "  ".join(f"{v:16d}" for v in d) + f"   {nxt:8d} {tail:8d} {rel:8d}")
print("sum    " + "  ".join(f"{int(v):16d}" for v in tot))
print("epilogue detail (cycles): layer  tap combine (accumulators ready -> GroupNorm entry) | own-part statistics | slice exchange | cluster exchange | normalise + Mish")
ed = np.zeros(5)
for i, r in enumerate(rows):
    if r[10] and r[13]:
        d = [r[10] - r[4], r[11] - r[10], r[12] - r[11], r[13] - r[12], r[5] - r[13]]
        ed += np.array(d)
        print(f"{i:4d}   " + "  ".join(f"{v:8d}" for v in d))
print("sum    " + "  ".join(f"{int(v):8d}" for v in ed) + "   us: " + "  ".join(f"{v / 1965:.1f}" for v in ed))
hr = rows[-1]
print(f"head layer (TN=64): GroupNorm done {hr[5]-hr[4]}, end {hr[6]-hr[4]} cycles after the accumulators were ready")
print(f"tail total {sum(tails) / 1e3:.1f} us, release total {sum(rels) / 1e3:.1f} us (includes non-tcgen05 kernels between steps)")
print("us@1.965GHz " + "  ".join(f"{v / 1965:16.1f}" for v in tot))
