"""Drop-in ``TemporalMapUnet`` whose forward runs hand-written sm_100a kernels through the C ABI.

Mirrors the reference interface (modeling/temporal.py:58-258): constructor arguments, ``forward(x, img, time,
cond=None, return_action_and_time_only=False)``, attributes ``perception`` / ``state_pred`` / ``magic_num`` /
``use_cond``, and — because published checkpoints are loaded by key and EMA weights positionally
(interact.py:102-108, misc/load_param.py:4-8) — the exact ``state_dict()`` keys and ``parameters()`` order.

The module tree is built from a flat (key, shape) table, so parameters live in plain container modules; the
arithmetic lives in ``csrc/`` and is reached with ctypes.  There is no PyTorch/CPU fallback for the denoiser: calling
``forward`` without a CUDA device (or without the built library) raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .constant import GuidanceType

Spec = Tuple[str, Tuple[int, ...], str]  # key, shape, kind ("param" | "buffer:<init>")


# ----------------------------------------------------------------------------------------------------------
# key/shape tables
# ----------------------------------------------------------------------------------------------------------
def _norm2d(prefix: str, c: int) -> List[Spec]:
    return [(f"{prefix}.weight", (c,), "ones"), (f"{prefix}.bias", (c,), "zeros"),
            (f"{prefix}.running_mean", (c,), "buffer:zeros"), (f"{prefix}.running_var", (c,), "buffer:ones"),
            (f"{prefix}.num_batches_tracked", (), "buffer:long")]


def _encoder_table(prefix: str, out_dim: int) -> List[Spec]:
    """ResNet-34 (BasicBlock x [3,4,6,3]) with fc -> out_dim (modeling/resnet.py:163-296, temporal.py:83-84)."""
    t: List[Spec] = [(f"{prefix}.conv1.weight", (64, 3, 7, 7), "conv2d")] + _norm2d(f"{prefix}.bn1", 64)
    cin = 64
    for stage, (width, depth) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for blk in range(depth):
            p = f"{prefix}.layer{stage}.{blk}"
            t += [(f"{p}.conv1.weight", (width, cin, 3, 3), "conv2d")] + _norm2d(f"{p}.bn1", width)
            t += [(f"{p}.conv2.weight", (width, width, 3, 3), "conv2d")] + _norm2d(f"{p}.bn2", width)
            if blk == 0 and (stage > 1 or cin != width):
                t += [(f"{p}.downsample.0.weight", (width, cin, 1, 1), "conv2d")] + _norm2d(f"{p}.downsample.1", width)
            cin = width
    return t + [(f"{prefix}.fc.weight", (out_dim, 512), "linear"), (f"{prefix}.fc.bias", (out_dim,), "bias:512")]


def _conv_gn(prefix: str, cin: int, cout: int, k: int) -> List[Spec]:
    return [(f"{prefix}.block.0.weight", (cout, cin, k), "linear"), (f"{prefix}.block.0.bias", (cout,), f"bias:{cin * k}"),
            (f"{prefix}.block.2.weight", (cout,), "ones"), (f"{prefix}.block.2.bias", (cout,), "zeros")]


def _res_block(prefix: str, cin: int, cout: int, embed: int) -> List[Spec]:
    t = _conv_gn(f"{prefix}.blocks.0", cin, cout, 5) + _conv_gn(f"{prefix}.blocks.1", cout, cout, 5)
    t += [(f"{prefix}.time_mlp.1.weight", (cout, embed), "linear"), (f"{prefix}.time_mlp.1.bias", (cout,), f"bias:{embed}")]
    if cin != cout:
        t += [(f"{prefix}.residual_conv.weight", (cout, cin, 1), "linear"), (f"{prefix}.residual_conv.bias", (cout,), f"bias:{cin}")]
    return t


def _state_pred_table(prefix: str, out_dim: int, hidden: int = 64, layers: int = 2) -> List[Spec]:
    t: List[Spec] = [(f"{prefix}.input_proj.weight", (hidden, 3), "xavier"), (f"{prefix}.input_proj.bias", (hidden,), "bias:3")]
    for i in range(layers):
        p = f"{prefix}.encoder_traj.layers.{i}"
        t += [(f"{p}.self_attn.in_proj_weight", (3 * hidden, hidden), "xavier"), (f"{p}.self_attn.in_proj_bias", (3 * hidden,), "zeros"),
              (f"{p}.self_attn.out_proj.weight", (hidden, hidden), "xavier"), (f"{p}.self_attn.out_proj.bias", (hidden,), "zeros"),
              (f"{p}.linear1.weight", (4 * hidden, hidden), "xavier"), (f"{p}.linear1.bias", (4 * hidden,), f"bias:{hidden}"),
              (f"{p}.linear2.weight", (hidden, 4 * hidden), "xavier"), (f"{p}.linear2.bias", (hidden,), f"bias:{4 * hidden}"),
              (f"{p}.norm1.weight", (hidden,), "ones"), (f"{p}.norm1.bias", (hidden,), "zeros"),
              (f"{p}.norm2.weight", (hidden,), "ones"), (f"{p}.norm2.bias", (hidden,), "zeros")]
    return t + [(f"{prefix}.encoder_traj.norm.weight", (hidden,), "ones"), (f"{prefix}.encoder_traj.norm.bias", (hidden,), "zeros"),
                (f"{prefix}.output_proj.weight", (out_dim, hidden), "xavier"), (f"{prefix}.output_proj.bias", (out_dim,), f"bias:{hidden}")]


def parameter_table(use_cond: GuidanceType, transition_dim: int, dim: int, dim_mults: Sequence[int]) -> List[Spec]:
    """(key, shape, init) in the reference's registration order: perception, [cond_mlp],
    time_mlp, downs, ups, mid_block1, mid_block2, final_conv | act_conv + state_pred."""
    widths = [transition_dim] + [dim * m for m in dim_mults]
    pairs = list(zip(widths[:-1], widths[1:]))
    embed = 2 * dim
    t = _encoder_table("perception", dim)
    if use_cond == GuidanceType.FREE_GUIDANCE:
        t += [("cond_mlp.0.weight", (dim, 2), "linear"), ("cond_mlp.0.bias", (dim,), "bias:2"),
              ("cond_mlp.2.weight", (dim, dim), "linear"), ("cond_mlp.2.bias", (dim,), f"bias:{dim}")]
    t += [("time_mlp.1.weight", (4 * dim, dim), "linear"), ("time_mlp.1.bias", (4 * dim,), f"bias:{dim}"),
          ("time_mlp.3.weight", (dim, 4 * dim), "linear"), ("time_mlp.3.bias", (dim,), f"bias:{4 * dim}")]
    for i, (a, b) in enumerate(pairs):
        t += _res_block(f"downs.{i}.0", a, b, embed) + _res_block(f"downs.{i}.1", b, b, embed)
        if i < len(pairs) - 1:
            t += [(f"downs.{i}.3.conv.weight", (b, b, 3), "linear"), (f"downs.{i}.3.conv.bias", (b,), f"bias:{3 * b}")]
    for i, (a, b) in enumerate(reversed(pairs[1:])):
        t += _res_block(f"ups.{i}.0", 2 * b, a, embed) + _res_block(f"ups.{i}.1", a, a, embed)
        t += [(f"ups.{i}.3.conv.weight", (a, a, 4), "linear"), (f"ups.{i}.3.conv.bias", (a,), f"bias:{4 * a}")]
    mid = widths[-1]
    t += _res_block("mid_block1", mid, mid, embed) + _res_block("mid_block2", mid, mid, embed)
    w0 = widths[1]
    if use_cond == GuidanceType.CLASSIFIER_GUIDANCE:
        t += _conv_gn("act_conv.0", w0, w0, 5) + [("act_conv.1.weight", (3, w0, 1), "linear"), ("act_conv.1.bias", (3,), f"bias:{w0}")]
        t += _state_pred_table("state_pred", transition_dim - 3)
    else:
        t += _conv_gn("final_conv.0", w0, w0, 5)
        t += [("final_conv.1.weight", (transition_dim, w0, 1), "linear"), ("final_conv.1.bias", (transition_dim,), f"bias:{w0}")]
    return t


def _init_tensor(shape, kind: str) -> torch.Tensor:
    if kind == "ones":
        return torch.ones(shape)
    if kind == "zeros":
        return torch.zeros(shape)
    if kind.startswith("bias:"):
        bound = 1.0 / math.sqrt(int(kind[5:]))
        return torch.empty(shape).uniform_(-bound, bound)
    t = torch.empty(shape)
    if kind == "linear":      # nn.Linear / nn.Conv1d default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        fan_in = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else shape[0]
        bound = 1.0 / math.sqrt(fan_in)
        return t.uniform_(-bound, bound)
    if kind == "conv2d":      # modeling/resnet.py:211-213 kaiming_normal(fan_out, relu)
        return nn.init.kaiming_normal_(t, mode="fan_out", nonlinearity="relu")
    if kind == "xavier":      # modeling/helpers.py:48-51
        return nn.init.xavier_uniform_(t)
    raise ValueError(kind)


class _Tree(nn.Module):
    """Plain container node; children and leaves are attached by dotted key."""

    def _attach(self, path: List[str], shape, kind: str, special: Dict[str, type], prefix: str = ""):
        name = path[0]
        if len(path) == 1:
            if kind.startswith("buffer:"):
                init = kind[7:]
                val = torch.tensor(0, dtype=torch.long) if init == "long" else (torch.ones(shape) if init == "ones" else torch.zeros(shape))
                self.register_buffer(name, val)
            else:
                self.register_parameter(name, nn.Parameter(_init_tensor(shape, kind)))
            return
        if name not in self._modules:
            full = prefix + name
            self.add_module(name, special.get(full, _Tree)())
        self._modules[name]._attach(path[1:], shape, kind, special, prefix + name + ".")


class ImageEncoder(_Tree):
    """``model.perception``: ResNet-34 -> Linear(512, dim) in eval mode (modeling/resnet.py, modeling/temporal.py:83-84).
    SURVEY.md 8f rank 1, a 'next' row: library kernels (cuDNN through torch), hoisted out of the sampling loop because
    the feature is step-invariant.  Inference only — BatchNorm always uses its running statistics.

    On CUDA without autograd the forward is restructured for per-tick latency (a closed-loop agent encodes a new camera
    frame before every plan): BatchNorm folded into the convolution weights, channels-last, cuDNN fused
    conv+bias(+residual)+ReLU, and the whole network replayed as one CUDA graph per input shape (36 launches instead of
    ~110 eager ones).  TF32 follows ``torch.backends.cudnn.allow_tf32`` as for any torch convolution."""

    compute_dtype = "fp32"   # "fp32" (TF32 follows torch.backends.cudnn.allow_tf32) or "bf16" (bf16 channels-last;
    #                           feature error <= 3e-2 of its max-abs vs the fp32 golden, tests/test_gpu_parity.py) — set_precision()

    bf16_body = "tcgen05"    # "tcgen05": layer1..layer4 by csrc/encoder_conv.cu (hand-written implicit GEMM);  "cudnn": torch's fused cuDNN calls (A/B timing)

    def set_precision(self, precision: str, body: str = "tcgen05") -> "ImageEncoder":
        if precision not in ("fp32", "bf16"):
            raise ValueError("encoder precision must be 'fp32' or 'bf16'")
        if body not in ("tcgen05", "cudnn"):
            raise ValueError("encoder body must be 'tcgen05' or 'cudnn'")
        object.__setattr__(self, "compute_dtype", precision)
        object.__setattr__(self, "bf16_body", body)
        self.invalidate()
        return self

    def _bn(self, x, m):
        return F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, 1e-5)

    def _blocks(self):
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in stage._modules.values():
                yield blk, (2 if "downsample" in blk._modules else 1)   # ResNet-34: a projection shortcut <=> a stride-2 block

    def _forward_plain(self, img: torch.Tensor) -> torch.Tensor:
        x = F.relu(self._bn(F.conv2d(img, self.conv1.weight, None, 2, 3), self.bn1))
        x = F.max_pool2d(x, 3, 2, 1)
        for blk, stride in self._blocks():
            y = F.relu(self._bn(F.conv2d(x, blk.conv1.weight, None, stride, 1), blk.bn1))
            y = self._bn(F.conv2d(y, blk.conv2.weight, None, 1, 1), blk.bn2)
            if "downsample" in blk._modules:
                ds = blk.downsample._modules
                x = self._bn(F.conv2d(x, ds["0"].weight, None, stride, 0), ds["1"])
            x = F.relu(y + x)
        return F.linear(F.adaptive_avg_pool2d(x, 1).flatten(1), self.fc.weight, self.fc.bias)

    # ---- folded / fused / graphed inference path --------------------------------------------------------
    def _fold(self, conv_w, bn):
        dt = torch.bfloat16 if self.compute_dtype == "bf16" else torch.float32
        scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + 1e-5)
        w = (conv_w.detach().float() * scale.view(-1, 1, 1, 1)).to(dt).contiguous(memory_format=torch.channels_last)
        return w, (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).to(dt).contiguous()

    def _fold_packed(self, conv_w, bn):
        """conv + BatchNorm as the operands ``b2p_encoder_conv_bf16`` takes: bf16 [k*k][C_out][C_in] with the BatchNorm scale folded in, fp32 bias."""
        scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + 1e-5)
        w = conv_w.detach().float() * scale.view(-1, 1, 1, 1)
        co, ci, kh, kw = w.shape
        wp = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(torch.bfloat16).contiguous()
        return wp, (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()

    @staticmethod
    def _conv_tc(x, wp, bias, res, ksize, stride, relu):
        """One folded convolution on NHWC bf16 activations by csrc/encoder_conv.cu (x: [N,H,W,C_in] contiguous)."""
        from . import _lib
        import ctypes as C
        n, h, w, ci = x.shape
        co = wp.shape[1]
        out = torch.empty((n, (h - 1) // stride + 1, (w - 1) // stride + 1, co), dtype=torch.bfloat16, device=x.device)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().b2p_encoder_conv_bf16(_lib.ptr(x), n, h, w, ci, _lib.ptr(wp), _lib.ptr(bias), _lib.ptr(res) if res is not None else None,
                                                         _lib.ptr(out), co, ksize, stride, 1 if relu else 0, stream), None, "b2p_encoder_conv_bf16")
        return out

    def _stem_operands(self):
        """conv1 with bn1 folded, as the operand image ``b2p_encoder_stem_bf16`` takes (include/b200plan.h): K index
        (kernel row, kernel column, channel) padded 147 -> 192, bf16, [K chunk of 64][out channel][16-byte unit ^ (channel % 8)][8]."""
        bn = self.bn1
        scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + 1e-5)
        w = self.conv1.weight.detach().float() * scale.view(-1, 1, 1, 1)                       # [64, 3, 7, 7]
        wk = F.pad(w.permute(0, 2, 3, 1).reshape(64, 147), (0, 45)).to(torch.bfloat16).view(64, 3, 8, 8)
        n = torch.arange(64, device=w.device).view(64, 1, 1)
        src = (torch.arange(8, device=w.device).view(1, 1, 8) ^ (n & 7)).expand(64, 3, 8)
        image = torch.gather(wk, 2, src.unsqueeze(-1).expand(64, 3, 8, 8)).permute(1, 0, 2, 3).contiguous()
        bias = (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()
        return image, bias

    def _stem_bf16(self, img: torch.Tensor, image: torch.Tensor, bias: torch.Tensor, fused: bool = True) -> torch.Tensor:
        """conv1 + bn1 + relu + maxpool (modeling/resnet.py:279-282) by the hand-written kernels of csrc/encoder_stem.cu -> bf16 channels-last.
        ``fused``: one kernel (conv1's output never reaches HBM); otherwise the conv kernel followed by the pool kernel (same bits)."""
        from . import _lib
        import ctypes as C
        n, c, h, w = img.shape
        if c != 3:
            raise ValueError("the encoder expects [N,3,H,W] images")
        oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        ph, pw = (oh - 1) // 2 + 1, (ow - 1) // 2 + 1
        x = torch.empty((n, 64, ph, pw), dtype=torch.bfloat16, device=img.device, memory_format=torch.channels_last)
        lib = _lib.load()
        stream = C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)
        sn, sc, sh, sw = img.stride()
        with torch.cuda.device(img.device):
            if fused and min(sc, sh, sw) >= 0:
                _lib.check(lib.b2p_encoder_stem_pool_bf16(_lib.ptr(img), sn, sc, sh, sw, n, h, w, _lib.ptr(image), _lib.ptr(bias), _lib.ptr(x), stream),
                           None, "b2p_encoder_stem_pool_bf16")
                return x
            y = torch.empty((n, 64, oh, ow), dtype=torch.bfloat16, device=img.device, memory_format=torch.channels_last)
            _lib.check(lib.b2p_encoder_stem_bf16(_lib.ptr(img), sn, sc, sh, sw, n, h, w, _lib.ptr(image), _lib.ptr(bias), _lib.ptr(y), stream),
                       None, "b2p_encoder_stem_bf16")
            _lib.check(lib.b2p_maxpool3x3s2_nhwc_bf16(_lib.ptr(y), _lib.ptr(x), n, oh, ow, 64, stream), None, "b2p_maxpool3x3s2_nhwc_bf16")
        return x

    def _tensors(self):
        if getattr(self, "_tlist", None) is None:
            object.__setattr__(self, "_tlist", list(self.state_dict(keep_vars=True).values()))
        return self._tlist

    def _apply(self, fn, *args, **kwargs):
        object.__setattr__(self, "_tlist", None)
        return super()._apply(fn, *args, **kwargs)

    def invalidate(self):
        """Drop the folded weights and the captured graphs (needed after a write the version counters cannot see, e.g.
        through ``param.data``)."""
        object.__setattr__(self, "_tlist", None)
        object.__setattr__(self, "_fold_cache", None)
        object.__setattr__(self, "_graphs", {})
        object.__setattr__(self, "_gen", getattr(self, "_gen", 0) + 1)

    def weights_key(self) -> tuple:
        """Identifies the current encoder weights: (generation, sum of version counters)."""
        return (getattr(self, "_gen", 0), sum([t._version for t in self._tensors()]))

    def _folded(self):
        key = (self.compute_dtype, self.bf16_body) + tuple([(t.data_ptr(), t._version) for t in self._tensors()])
        cache = getattr(self, "_fold_cache", None)
        if cache is None or cache[0] != key:
            stem = self._stem_operands() if self.compute_dtype == "bf16" else self._fold(self.conv1.weight, self.bn1)
            blocks = []
            fold = self._fold_packed if (self.compute_dtype == "bf16" and self.bf16_body == "tcgen05") else self._fold
            for blk, stride in self._blocks():
                ds = None
                if "downsample" in blk._modules:
                    d = blk.downsample._modules
                    ds = fold(d["0"].weight, d["1"])
                blocks.append((stride, fold(blk.conv1.weight, blk.bn1), fold(blk.conv2.weight, blk.bn2), ds))
            cache = (key, stem, blocks)
            object.__setattr__(self, "_fold_cache", cache)
            object.__setattr__(self, "_graphs", {})
        return cache

    def _forward_fused(self, img: torch.Tensor) -> torch.Tensor:
        _, (w, b), blocks = self._folded()
        one, zero = [1, 1], [0, 0]
        if self.compute_dtype == "bf16":
            x = self._stem_bf16(img, w, b)
            if self.bf16_body == "tcgen05":     # every convolution of layer1..layer4 hand-written (csrc/encoder_conv.cu), NHWC bf16 throughout
                x = x.permute(0, 2, 3, 1)       # the stem's channels-last tensor viewed as [N,H,W,C] (contiguous)
                for stride, (w1, b1), (w2, b2), ds in blocks:
                    y = self._conv_tc(x, w1, b1, None, 3, stride, True)
                    r = self._conv_tc(x, ds[0], ds[1], None, 1, stride, False) if ds is not None else x
                    x = self._conv_tc(y, w2, b2, r, 3, 1, True)
                return F.linear(x.float().mean((1, 2)), self.fc.weight.float(), self.fc.bias.float())   # pooling and fc in fp32
        else:
            x = torch.cudnn_convolution_relu(img.contiguous(memory_format=torch.channels_last), w, b, [2, 2], [3, 3], one, 1)
            x = F.max_pool2d(x, 3, 2, 1)
        for stride, (w1, b1), (w2, b2), ds in blocks:
            y = torch.cudnn_convolution_relu(x, w1, b1, [stride, stride], one, one, 1)
            if ds is not None:
                x = F.conv2d(x, ds[0], ds[1], stride, 0)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, one, one, 1)
        return F.linear(x.float().mean((2, 3)), self.fc.weight.float(), self.fc.bias.float())   # pooling and fc in fp32

    def _forward_graphed(self, img: torch.Tensor) -> torch.Tensor:
        self._folded()   # (re)builds the folded weights and drops stale graphs when a parameter changed
        key = (tuple(img.shape), img.device, torch.backends.cudnn.allow_tf32, self.compute_dtype, self.bf16_body)
        g = self._graphs.get(key)
        if g is None:
            static_in = torch.empty_like(img, memory_format=torch.channels_last)
            static_in.copy_(img)
            side = torch.cuda.Stream(device=img.device)
            side.wait_stream(torch.cuda.current_stream(img.device))
            with torch.cuda.stream(side):
                for _ in range(2):      # cuDNN algorithm selection and workspace allocation happen outside the capture
                    self._forward_fused(static_in)
            torch.cuda.current_stream(img.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._forward_fused(static_in)
            g = (graph, static_in, static_out)
            while len(self._graphs) >= 4:       # a graph pins its activation pool (GBs at large batch): keep the 4 most recent shapes
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = g
        else:
            self._graphs[key] = self._graphs.pop(key)   # most recently used last
        graph, static_in, static_out = g
        static_in.copy_(img)
        graph.replay()
        return static_out.clone()

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        if img.is_cuda and img.dtype == torch.float32 and not (torch.is_grad_enabled() and img.requires_grad) and hasattr(torch, "cudnn_convolution_add_relu"):
            with torch.no_grad():
                if torch.cuda.is_current_stream_capturing():
                    return self._forward_fused(img)
                if self.compute_dtype == "bf16" and self.bf16_body == "tcgen05" and img.shape[0] >= 16:
                    # many frames: the 37 launches are GPU-bound, and replaying a graph would first copy the whole batch into its static input
                    # (0.5 ms for 256 frames); the graph is for the per-tick latency of a few frames
                    return self._forward_fused(img)
                return self._forward_graphed(img)
        return self._forward_plain(img)


class _StatePredFn(torch.autograd.Function):
    """TrajPredict forward with an analytic VJP wrt ``action`` so that the reference's
    ``torch.autograd.grad(loss, [x_guidance, action])`` (control/guidance.py:45-48) works unchanged on top of it."""

    @staticmethod
    def forward(ctx, action, time_embed, owner):
        model = owner()
        B, S, _ = action.shape
        full = torch.zeros(B, S + 1, 3, device=action.device, dtype=torch.float32)
        full[:, :S] = action
        te = time_embed.detach().contiguous().float()
        out = torch.empty(B, S, model.transition_dim - 3, device=action.device, dtype=torch.float32)
        h = model._handle_for(action.device)
        _lib.check(_lib.load().b2p_state_pred(h, _lib.ptr(full), _lib.ptr(te), _lib.ptr(out), B, model._stream()), h, "b2p_state_pred")
        ctx.owner, ctx.full, ctx.te = owner, full, te
        return out

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.owner()
        B = ctx.full.shape[0]
        g = grad_out.contiguous().float()
        ga = torch.empty_like(ctx.full)
        h = model._handle_for(g.device)
        _lib.check(_lib.load().b2p_state_pred_vjp(h, _lib.ptr(ctx.full), _lib.ptr(ctx.te), _lib.ptr(g), _lib.ptr(ga), B, model._stream()),
                   h, "b2p_state_pred_vjp")
        return ga[:, :-1], None, None


class StatePredictor(_Tree):
    """``model.state_pred(action[:, :-1], time_embed)`` (modeling/helpers.py:53-59, interact.py:158)."""

    def forward(self, action: torch.Tensor, time_embed: torch.Tensor) -> torch.Tensor:
        import weakref
        return _StatePredFn.apply(action.contiguous().float(), time_embed, weakref.ref(self._owner()))


class TemporalMapUnet(nn.Module):
    def __init__(self, horizon, transition_dim=2, attention=False, dim=128, dim_mults=(1, 2, 4, 8),
                 diffuser_building_block="concat", use_cond=GuidanceType.NO_GUIDANCE, precision: str = "fp32",
                 small_batch_max: int = 4):
        super().__init__()
        if diffuser_building_block != "concat":
            raise NotImplementedError  # modeling/temporal.py:72-75
        if attention:
            raise NotImplementedError("LinearAttention is never enabled by the shipped configs (MODEL.USE_ATTN=False) and is out of scope")
        if isinstance(use_cond, str):
            use_cond = GuidanceType[use_cond]
        self.horizon, self.transition_dim, self.dim, self.dim_mults = int(horizon), int(transition_dim), int(dim), tuple(int(m) for m in dim_mults)
        self.use_cond = use_cond
        self.precision = precision
        self.small_batch_max = int(small_batch_max)
        self.magic_num = 23.315
        special = {"perception": ImageEncoder, "state_pred": StatePredictor}
        root = _Tree()
        for key, shape, kind in parameter_table(use_cond, self.transition_dim, self.dim, self.dim_mults):
            root._attach(key.split("."), shape, kind, special)
        for name, child in root._modules.items():   # re-parent in registration order
            self.add_module(name, child)
        if use_cond == GuidanceType.CLASSIFIER_GUIDANCE:
            import weakref
            object.__setattr__(self.state_pred, "_owner", weakref.ref(self))
        self._handles: Dict[int, C.c_void_p] = {}
        self._packed: Dict[int, tuple] = {}
        self._feat_cache: Optional[tuple] = None
        self._tensor_list: Optional[list] = None
        self._tensor_gen = 0
        import weakref
        me = weakref.ref(self)
        for p in self.parameters():             # lets checkpoint.copy_parameters find the model that owns a parameter
            object.__setattr__(p, "_b2p_owner", me)

    # ---- C handle management --------------------------------------------------------------------------
    def _unet_items(self) -> Iterable[Tuple[str, torch.Tensor]]:
        for k, v in self.state_dict(keep_vars=True).items():
            if not k.startswith("perception."):
                yield k, v

    def _unet_tensors(self) -> list:
        """The denoiser's tensors, listed once (walking state_dict() costs ~0.5 ms, more than a 2-step plan); the list is
        dropped whenever the module is converted / moved / reloaded (_apply, load_state_dict).  In-place updates are seen
        through the tensors' version counters."""
        if self._tensor_list is None:
            self._tensor_list = [v for _, v in self._unet_items()]
        return self._tensor_list

    def _version_key(self) -> tuple:
        """Changes whenever a denoiser tensor is modified in place (version counters) or replaced / moved (the tensor list is
        rebuilt by _apply / load_state_dict, which bump the generation term)."""
        ts = self._unet_tensors()
        return (self._tensor_gen, sum([t._version for t in ts]))

    def _apply(self, fn, *args, **kwargs):
        self._tensor_list, self._tensor_gen, self._feat_cache = None, getattr(self, "_tensor_gen", 0) + 1, None
        out = super()._apply(fn, *args, **kwargs)
        import weakref
        me = weakref.ref(self)
        for p in self.parameters():             # .to()/.half() may replace Parameter objects
            object.__setattr__(p, "_b2p_owner", me)
        return out

    def load_state_dict(self, *args, **kwargs):
        self._tensor_list, self._tensor_gen, self._feat_cache = None, getattr(self, "_tensor_gen", 0) + 1, None
        return super().load_state_dict(*args, **kwargs)

    def invalidate_weights(self) -> "TemporalMapUnet":
        """Force the packed device weights, the folded encoder weights, the encoder graphs and the cached image feature to
        be rebuilt on the next call.  In-place updates (``p.copy_()``, optimizers, ``load_state_dict``, ``.to()``) are
        detected automatically through the tensors' version counters; a write through ``param.data`` (the idiom of the
        reference's misc/load_param.copy_parameters and of EMA ``copy_to``) does NOT bump them, so call this afterwards —
        ``checkpoint.copy_parameters`` does."""
        self._tensor_list, self._tensor_gen, self._feat_cache = None, self._tensor_gen + 1, None
        self.perception.invalidate()
        return self

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _handle_for(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("TemporalMapUnet runs on CUDA (sm_100a) only; there is no CPU fallback")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        lib = _lib.load()
        if idx not in self._handles:
            cfg = _lib.ModelConfig()
            cfg.horizon, cfg.transition_dim, cfg.dim, cfg.n_mults = self.horizon, self.transition_dim, self.dim, len(self.dim_mults)
            for i, m in enumerate(self.dim_mults):
                cfg.dim_mults[i] = m
            cfg.guidance = self.use_cond.value
            cfg.precision = _lib.PRECISIONS[self.precision]
            h = C.c_void_p()
            _lib.check(lib.b2p_create(C.byref(cfg), idx, C.byref(h)), None, "b2p_create")
            _lib.check(lib.b2p_set_small_batch_max(h, self.small_batch_max), h, "b2p_set_small_batch_max")
            if getattr(self, "_chain", None) is not None:
                _lib.check(lib.b2p_set_chain(h, int(self._chain)), h, "b2p_set_chain")
            self._handles[idx] = h
        h = self._handles[idx]
        key = self._version_key()
        if self._packed.get(idx) != key:
            for k, v in self._unet_items():
                host = v.detach().to("cpu", torch.float32).contiguous()
                _lib.check(lib.b2p_load_weight(h, k.encode(), _lib.ptr(host), host.numel()), h, f"b2p_load_weight({k})")
            _lib.check(lib.b2p_finalize_weights(h), h, "b2p_finalize_weights")
            self._packed[idx] = key
        return h

    def set_precision(self, precision: str) -> "TemporalMapUnet":
        """'fp32' (CUDA-core FFMA, exact fp32), 'bf16x3' (tcgen05, bf16 hi/lo split, fp32-class parity) or 'bf16' (tcgen05 single pass)."""
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.precision = precision
        lib = _lib.load()
        for h in self._handles.values():
            _lib.check(lib.b2p_set_precision(h, _lib.PRECISIONS[precision]), h, "b2p_set_precision")
        return self

    def set_small_batch_max(self, max_samples: int) -> "TemporalMapUnet":
        """Evaluations of at most `max_samples` trajectories (CFG doubling included) run the small-batch exact-fp32 GEMV
        kernels in every precision mode (single-trajectory closed-loop planning); 0 turns that path off."""
        if max_samples < 0:
            raise ValueError("max_samples must be >= 0")
        self.small_batch_max = int(max_samples)
        lib = _lib.load()
        for h in self._handles.values():
            _lib.check(lib.b2p_set_small_batch_max(h, self.small_batch_max), h, "b2p_set_small_batch_max")
        return self

    def set_chain(self, enabled: bool) -> "TemporalMapUnet":
        """Developer switch: False runs every layer of the tensor-core precisions as its own launch instead of the row-owned
        chain kernels (csrc/chain64.cu).  Same results up to fp32 summation order."""
        self._chain = bool(enabled)
        lib = _lib.load()
        for h in self._handles.values():
            _lib.check(lib.b2p_set_chain(h, int(self._chain)), h, "b2p_set_chain")
        return self

    def __del__(self):
        try:
            lib = _lib.load()
            for h in getattr(self, "_handles", {}).values():
                lib.b2p_destroy(h)
        except Exception:
            pass

    # ---- conditioning feature -------------------------------------------------------------------------
    def encode(self, img: torch.Tensor, use_cache: bool = True) -> torch.Tensor:
        """perception(img) with a one-entry cache: the reference re-runs the encoder inside every denoising step
        (modeling/temporal.py:203); in eval mode the feature is step-invariant, so it is computed once per image tensor.
        The cache key covers the image tensor (address, version, geometry) AND the encoder weights (generation + version
        counters).  Pass ``use_cache=False`` for a frame buffer that an external producer (DLPack, cupy, ``.data``) refills
        in place without bumping its version counter."""
        if img.dim() == 2:
            return img
        key = (img.data_ptr(), img._version, tuple(img.shape), tuple(img.stride()), img.device, self._tensor_gen, self.perception.weights_key())
        if use_cache and self._feat_cache is not None and self._feat_cache[0] == key:
            return self._feat_cache[1]
        with torch.no_grad():
            feat = self.perception(img).float().contiguous()
        self._feat_cache = (key, feat, img)   # the image is kept alive: its address cannot be recycled for another frame while cached
        return feat

    # ---- reference surface ----------------------------------------------------------------------------
    def forward(self, x, img, time, cond=None, return_action_and_time_only=False):
        """x [B,H,D]; img [B|S,3,h,w] image or [B|S,dim] precomputed feature; time [1]|[B] ints; cond None|[B,2]."""
        if x.device.type != "cuda":
            raise RuntimeError("TemporalMapUnet.forward needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        h = self._handle_for(x.device)
        B = x.shape[0]
        if x.dim() != 3 or tuple(x.shape[1:]) != (self.horizon, self.transition_dim):
            raise ValueError(f"x must be [B,{self.horizon},{self.transition_dim}], got {tuple(x.shape)}")
        time = torch.as_tensor(time, device=x.device)
        if B == 0:   # empty batch: nothing to launch (torch modules return empty tensors too)
            if self.use_cond == GuidanceType.CLASSIFIER_GUIDANCE and return_action_and_time_only:
                return x.new_zeros((0, self.horizon, 3)), x.new_zeros((0, self.dim))
            return x.new_zeros((0, self.horizon, self.transition_dim))
        x = x.detach().contiguous().float()
        feat = self.encode(img).detach().contiguous().float()
        t = time.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
        free = self.use_cond == GuidanceType.FREE_GUIDANCE
        if free:
            if cond is None:
                cond = torch.zeros((B, 2), device=x.device)   # modeling/temporal.py:207
            cond = cond.detach().to(x.device, torch.float32).contiguous()
            if cond.shape[0] != B:
                raise ValueError("cond rows must equal the batch of x")
        elif t.numel() not in (1, B) or feat.shape[0] not in (1, B):
            raise RuntimeError(f"Sizes of tensors must match: time {t.numel()}, feature {feat.shape[0]}, batch {B}")
        if B % feat.shape[0] or B % t.numel():
            raise RuntimeError(f"time ({t.numel()}) / feature ({feat.shape[0]}) rows must divide the batch ({B})")
        cls = self.use_cond == GuidanceType.CLASSIFIER_GUIDANCE
        out = None if (cls and return_action_and_time_only) else torch.empty_like(x)
        action = torch.empty(B, self.horizon, 3, device=x.device) if cls else None
        te = torch.empty(B, self.dim, device=x.device) if cls else None
        rc = lib.b2p_unet_forward(h, _lib.ptr(x), _lib.ptr(feat), feat.shape[0], _lib.ptr(t), t.numel(), _lib.ptr(cond if free else None),
                                  _lib.ptr(out), _lib.ptr(action), _lib.ptr(te), B, self._stream())
        _lib.check(rc, h, "b2p_unet_forward")
        if cls and return_action_and_time_only:
            return action, te
        return out

    def last_launch_count(self, device=None) -> int:
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        return int(_lib.load().b2p_last_launch_count(self._handles[idx]))


def build_model(cfg) -> TemporalMapUnet:
    """modeling/temporal.py:248-258 — accepts the reference's yacs tree or any object with the same attributes."""
    return TemporalMapUnet(horizon=cfg.MODEL.HORIZON, transition_dim=cfg.MODEL.TRANSITION_DIM, attention=cfg.MODEL.USE_ATTN,
                           dim=cfg.MODEL.DIM, dim_mults=cfg.MODEL.DIM_MULTS,
                           diffuser_building_block=cfg.MODEL.DIFFUSER_BUILDING_BLOCK, use_cond=GuidanceType[cfg.TRAIN.USE_COND],
                           precision=getattr(getattr(cfg, "B200", None), "PRECISION", "fp32") if hasattr(cfg, "B200") else "fp32",
                           small_batch_max=getattr(getattr(cfg, "B200", None), "SMALL_BATCH_MAX", 4) if hasattr(cfg, "B200") else 4)
