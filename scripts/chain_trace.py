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
print(f"B={B}: seam launch, CTA 0, cycles")
print("op  top->barrier  weights_wait(after barrier)  mma_issue  prefetch  mma_wait(exposed)  taps+GN  epilogue  total   [head: partial, sync1, sum+sync2, sched]")
tot = 0
for i in range(10):
    r = t[i]
    nxt = t[i + 1][0] if i < 9 else t[10][0]
    wl = f"  [weights: issued {r[12]-t[i-1][0] if i else 0} into the previous op, landed {r[2]-r[12] if r[12] else 0} cycles later]"
    extra = wl + f"   {r[9]-r[8]} {r[10]-r[9]} {r[11]-r[10]} {r[6]-r[11]}" if r[9] else wl
    print(f"{i:2d}  {r[1]-r[0]:8d}  {r[2]-r[1]:8d}  {r[3]-r[2]:8d}  {r[4]-r[1]:8d}  {r[5]-r[4]:8d}  {r[8]-r[5]:8d}  {r[6]-r[5]:8d}  {nxt-r[0]:8d}{extra}")
    tot += nxt - r[0]
print("sum", tot, "cycles =", tot / 1.965e3, "us")
