"""GPU check of the EXPERIMENTAL one-cluster-per-trajectory denoiser kernel (csrc/unet_cluster.cu, opt-in through
B2P_CLUSTER_EVAL=1; not on any default path): same results as the CPU oracle and as the default small-batch kernels, for a
single evaluation and for a whole DDIM plan replayed from a CUDA graph.  (File name: runs after the parity tests.)"""
import os

import pytest
import torch

import autonomous_driving_with_diffusion_model_b200 as P
from oracle import plan as OP
from oracle import unet as U
from oracle import weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(sd, T, env):
    old = os.environ.get("B2P_CLUSTER_EVAL")
    os.environ["B2P_CLUSTER_EVAL"] = env           # read once per handle, when the handle is created
    try:
        cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T))
        m = P.build_model(cfg)
        m.load_state_dict(sd)
        m = m.to(DEV).eval()
        x = torch.zeros(1, 16, 7, device=DEV)
        m(x, torch.zeros(1, 64, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))   # creates the handle now
    finally:
        if old is None:
            os.environ.pop("B2P_CLUSTER_EVAL", None)
        else:
            os.environ["B2P_CLUSTER_EVAL"] = old
    return m, P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)


def test_cluster_evaluation_matches_oracle_and_default_path():
    B, T = 1, 10        # the configuration measured on B200 in round 1 (profiles/r01_cluster_eval_check.json)
    sd = W.make_state_dict("NO_GUIDANCE", seed=0)
    inp = W.synth_inputs(B, 0, seed=21)
    x, f = inp["x"].to(DEV), inp["feat"].to(DEV)
    t = torch.full((B,), 37, dtype=torch.long, device=DEV)
    m0, p0 = _model(sd, T, "0")
    m1, p1 = _model(sd, T, "1")
    y0, y1 = m0(x, f, t), m1(x, f, t)
    assert m1.last_launch_count() == 3 and m0.last_launch_count() > 40      # embed + time-term GEMM + ONE denoiser launch
    with torch.no_grad():
        ref = U.unet_forward(sd, inp["x"], inp["feat"], t.cpu(), None, "NO_GUIDANCE")
    assert float((y1.cpu() - ref).abs().max()) <= 1e-4
    assert float((y1 - y0).abs().max()) <= 1e-4
    a, b = p0.plan(x, f, postprocess=False), p1.plan(x, f, postprocess=False)
    b2 = p1.plan(x, f, postprocess=False)                                   # graph replay
    assert torch.equal(b, b2)
    ref_plan = OP.plan(sd, "NO_GUIDANCE", "guidance_ddim", inp["x"], inp["feat"], T, postprocess=False)
    assert float((b.cpu() - ref_plan).abs().max()) <= 1e-3 and float((a - b).abs().max()) <= 1e-3
