"""EXPERIMENTAL one-cluster-per-trajectory denoiser kernel (csrc/unet_cluster.cu, opt-in with B2P_CLUSTER_EVAL=1): its
host-side program (slot liveness, tap ranges, weight chunking, the 16 per-CTA weight streams) is built by the real C++
builder and walked on the CPU the way the kernel walks it (tests/native/uc_emulate.cu); the result must equal the CPU
oracle's denoiser output.  No GPU is needed: only the kernel's thread mapping and barriers stay unverified here."""
import json
import os
import shutil
import struct
import subprocess

import pytest
import torch
import torch.nn.functional as F

from oracle import unet as U
from oracle import weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "autonomous_driving_with_diffusion_model_b200")


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    nvcc = _nvcc()
    if nvcc is None:
        pytest.skip("nvcc not available")
    from autonomous_driving_with_diffusion_model_b200 import build as B

    B.build()   # object files of the library (the harness links the kernels' host stubs)
    objs = [os.path.join(PKG, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "api.cu"]
    exe = str(tmp_path_factory.mktemp("uc") / "uc_emulate")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-ffp-contract=off",
           os.path.join(ROOT, "tests", "native", "uc_emulate.cu"), *objs, "-lcuda", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "cannot find -lcuda" in r.stderr:
        pytest.skip("libcuda stub not available for linking")
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("seed,t", [(0, 50), (3, 99), (5, 0)])
def test_cluster_program_walk_equals_oracle(emulator, tmp_path, seed, t):
    sd = W.make_state_dict("NO_GUIDANCE", seed=seed, with_perception=False)
    inp = W.synth_inputs(1, 0, seed=seed + 11)
    x, feat = inp["x"], inp["feat"]
    tt = torch.tensor([t])
    with torch.no_grad():
        expect = U.unet_forward(sd, x, feat, tt, None, "NO_GUIDANCE")
        ci = torch.cat([U.time_embedding(sd, tt), feat], dim=-1)
    path = tmp_path / "case.bin"
    with open(path, "wb") as f:
        keys = [k for k in sd if not k.startswith("perception.")]
        f.write(struct.pack("<i", len(keys)))
        for k in keys:
            v = sd[k].detach().to(torch.float32).contiguous()
            kb = k.encode()
            f.write(struct.pack("<i", len(kb)) + kb + struct.pack("<q", v.numel()))
            f.write(v.numpy().tobytes())
        for v in (x, ci, expect):
            f.write(v.detach().to(torch.float32).contiguous().numpy().tobytes())
    r = subprocess.run([emulator, str(path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["max_abs_err"] <= 1e-4
    assert d["smem_bytes"] <= 227 * 1024 and d["max_chunk_floats"] <= 8192 and 30 <= d["n_ops"] <= 48
