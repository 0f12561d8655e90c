"""Developer (GPU): b2p_encoder_conv_bf16 against torch's conv2d on the same bf16 operands (fp32 reference), every layer shape of the encoder body,
and its time on 256 frames next to cuDNN's."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from autonomous_driving_with_diffusion_model_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load()


def run(x, w, b, res, ksize, stride, relu):
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    oh, ow = (h - 1) // stride + 1, (wd - 1) // stride + 1
    wp = w.permute(2, 3, 0, 1).reshape(ksize * ksize, cout, cin).contiguous()
    out = torch.empty(n, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.b2p_encoder_conv_bf16(_lib.ptr(x), n, h, wd, cin, _lib.ptr(wp), _lib.ptr(b), _lib.ptr(res) if res is not None else None, _lib.ptr(out), cout, ksize, stride, relu, st)
    assert rc == 0, rc
    return out, wp


def ref(x, w, b, res, ksize, stride, relu):
    y = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, stride, 1 if ksize == 3 else 0)
    if res is not None:
        y = y + res.permute(0, 3, 1, 2).float()
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1)


torch.manual_seed(0)
bad = 0
cases = [(3, 64, 225, 64, 64, 3, 1), (2, 32, 113, 128, 128, 3, 1), (2, 16, 57, 256, 256, 3, 1), (3, 8, 29, 512, 512, 3, 1), (2, 64, 225, 64, 128, 3, 2), (2, 32, 113, 128, 256, 3, 2),
         (3, 16, 57, 256, 512, 3, 2), (2, 64, 225, 64, 128, 1, 2), (2, 16, 57, 256, 512, 1, 2), (1, 5, 3, 64, 64, 3, 1), (1, 17, 9, 64, 64, 3, 2), (5, 16, 8, 128, 64, 3, 1)]
for (n, h, w_, cin, cout, k, s) in cases:
    for use_res, relu in ((False, 1), (True, 1), (False, 0)):
        x = torch.randn(n, h, w_, cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, device=dev)
        oh, ow = (h - 1) // s + 1, (w_ - 1) // s + 1
        res = torch.randn(n, oh, ow, cout, device=dev).to(torch.bfloat16) if use_res else None
        out, _ = run(x, w, b, res, k, s, relu)
        torch.cuda.synchronize()
        r = ref(x, w, b, res, k, s, relu)
        err = (out.float() - r).abs().max().item()
        tol = 2e-2 * max(1.0, r.abs().max().item())
        flag = "" if err <= tol else "   <-- MISMATCH"
        bad += err > tol
        print(f"N={n} {h}x{w_} {cin}->{cout} k{k} s{s} res={int(use_res)} relu={relu}: max err {err:.3e} (ref max {r.abs().max().item():.2f}){flag}")
print("FAILED" if bad else "all cases match")
only = int(os.environ.get("ENC_ONLY", "0"))

# timing on 256 frames, per layer shape of ResNet-34 at 256 x 900 input
if len(sys.argv) > 1 and sys.argv[1] == "time":
    N = 256
    for (h, w_, cin, cout, k, s, count) in [(64, 225, 64, 64, 3, 1, 6), (64, 225, 64, 128, 3, 2, 1), (64, 225, 64, 128, 1, 2, 1), (32, 113, 128, 128, 3, 1, 7), (32, 113, 128, 256, 3, 2, 1),
                                            (16, 57, 256, 256, 3, 1, 11), (16, 57, 256, 512, 3, 2, 1), (8, 29, 512, 512, 3, 1, 5)]:
        if only and (cin != only or cout != only): continue
        x = torch.randn(N, h, w_, cin, device=dev).to(torch.bfloat16)
        w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, device=dev)
        oh, ow = (h - 1) // s + 1, (w_ - 1) // s + 1
        res = torch.randn(N, oh, ow, cout, device=dev).to(torch.bfloat16)
        for _ in range(2): run(x, w, b, res, k, s, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(x, w, b, res, k, s, 1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        xc = x.permute(0, 3, 1, 2)          # channels-last view
        wc = w.contiguous(memory_format=torch.channels_last)
        bc = b.to(torch.bfloat16)
        for _ in range(2): F.conv2d(xc, wc, bc, s, 1 if k == 3 else 0)
        e0.record()
        for _ in range(5): F.conv2d(xc, wc, bc, s, 1 if k == 3 else 0)
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 5
        for _ in range(2): run(x, w, b, None, k, s, 1)
        e0.record()
        for _ in range(5): run(x, w, b, None, k, s, 1)
        e1.record(); torch.cuda.synchronize()
        ms_nores = e0.elapsed_time(e1) / 5
        fl = 2.0 * N * oh * ow * cout * cin * k * k
        print(f"{h}x{w_} {cin}->{cout} k{k} s{s} (x{count} in ResNet-34): {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s (no residual: {ms_nores:.3f} ms) | torch conv2d {ms_t:.3f} ms = {fl / ms_t / 1e9:.0f} TFLOP/s")
