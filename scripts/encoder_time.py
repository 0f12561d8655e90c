"""Developer timing (GPU): the image encoder on 256 scenes, fp32/TF32 vs bf16, cudnn.benchmark off/on, with a per-kernel table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import synthetic as W

dev = "cuda:0"
scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = P.load_cfg()
m = P.build_model(cfg)
m.load_state_dict(W.make_state_dict("NO_GUIDANCE", seed=0))
m = m.to(dev).eval()
img = torch.randn(scenes, 3, 256, 900, device=dev)
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for prec, body in (("fp32", "cudnn"), ("bf16", "cudnn"), ("bf16", "tcgen05")):
        m.perception.set_precision(prec, body)
        for _ in range(2):
            m.perception(img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m.perception(img)
        e1.record(); torch.cuda.synchronize()
        print(f"benchmark={bench} {prec} body={body}: {e0.elapsed_time(e1) / 3:.2f} ms for {scenes} scenes", flush=True)
        if "--profile" in sys.argv and bench and body == "tcgen05":
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                m.perception(img); torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=90), flush=True)
