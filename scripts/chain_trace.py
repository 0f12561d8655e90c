"""Developer (GPU): in-kernel stage clocks of one seam launch of the chain kernel (CTA 0), B2P_CHAIN_TRACE=1."""
import ctypes as C, os, sys
os.environ["B2P_CHAIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import _lib, synthetic as W
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=10), B200=dict(PRECISION="bf16x3"))
m = P.build_model(cfg); m.load_state_dict(W.make_state_dict("NO_GUIDANCE", seed=0)); m = m.to(dev).eval()
pl = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
inp = W.synth_inputs(B, 0, 1)
for _ in range(3): pl.plan(inp["x"].to(dev), inp["feat"].to(dev))
torch.cuda.synchronize()
lib = C.CDLL(_lib.lib_path())
lib.b2p_debug_chain_trace.argtypes = [C.c_void_p, C.c_void_p]
buf = (C.c_uint64 * 256)()
assert lib.b2p_debug_chain_trace(m._handle_for(dev), buf) == 0
t = [[buf[i * 16 + j] for j in range(16)] for i in range(16)]
print(f"B={B}: seam launch, CTA 0, thread 0 (an epilogue thread of every op) + the MMA lane, cycles")
print("op   L  barrier_wait  mma_lane:weights_wait  mma_issue | addends  wait_half0  epi_half0(+wait_half1)  epi_half1  [head: partials+sync, sums+sync, scheduler]  next_top   total")
tot = 0
for i in range(10):
    r = t[i]
    nxt = t[i + 1][0] if i < 9 else t[10][0]
    head = f"{r[10]-r[8]:6d} {r[11]-r[10]:6d} {r[6]-r[11]:6d}" if r[10] else " " * 20
    print(f"{i:2d}  {r[1]-r[0]:8d}  {r[2]-r[14]:8d}  {r[3]-r[2]:8d} | {r[4]-r[1]:8d}  {r[5]-r[4]:8d}  {(r[9] if r[9] else r[8])-r[5]:8d}  {r[8]-r[9] if r[9] else 0:8d}  {head}  {nxt-r[6]:8d}  {nxt-r[0]:8d}   [MMA lane past the barrier {r[14]-r[1]} after thread 0; next weights issued {t[i+1][12]-r[1] if i < 9 and t[i+1][12] else 0} after the barrier]")
    tot += nxt - r[0]
print("sum", tot, "cycles =", tot / 1.965e3, "us")
