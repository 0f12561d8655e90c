"""CPU tests that pin the oracle: known-answer constants, golden vectors made by the real reference
(oracle/make_golden.py) and, when /root/reference is present, a live comparison with the reference code."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import plan as P
from oracle import reference_loader as RL
from oracle import schedulers as S
from oracle import unet as U
from oracle import weights as W
from oracle.make_golden import PLAN_CASES


def test_schedule_known_answers():
    # SURVEY.md §8c known-answer constants (T=100, squaredcos_cap_v2, fp32)
    ac = S.alphas_cumprod(100)
    kat = {0: 0.999368727, 1: 0.998252511, 10: 0.966716647, 50: 0.47826457, 80: 0.0851461962, 90: 0.0195443742,
           98: 0.000242857204, 99: 2.42854071e-07}
    for t, v in kat.items():
        assert abs(float(ac[t]) - v) <= 2e-7 * max(v, 1e-3), (t, float(ac[t]), v)
    assert list(S.leading_timesteps(100, 10)) == [90, 80, 70, 60, 50, 40, 30, 20, 10, 0]
    assert list(S.leading_timesteps(100, 2)) == [50, 0]
    assert list(S.leading_timesteps(100, 100)) == list(range(99, -1, -1))
    assert abs(float(S.ddim_variance(ac, 99, 98)) - 0.99875766) < 1e-6
    assert abs(float(S.ddim_variance(ac, 90, 80)) - 0.71890974) < 1e-6
    assert abs(float(S.ddim_variance(ac, 50, 0)) - 0.00063090731) < 1e-9
    assert float(S.ddpm_variance(ac, 0, -1)) == pytest.approx(1e-20)


def test_state_dict_spec_matches_reference(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, "state_dict_spec.json")))
    for mode in W.MODES:
        sd = W.make_state_dict(mode)
        assert [[k, list(v.shape)] for k, v in sd.items()] == spec[mode]["state_dict"]
    n = sum(int(np.prod(s)) for k, s, kind, _ in W.unet_specs("NO_GUIDANCE", with_perception=False))
    assert n == 16031815


def test_weights_are_reproducible(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_forward.npz"))
    for mode in W.MODES:
        assert W.state_dict_digest(W.make_state_dict(mode)) == str(g[mode + ".digest"])


def test_sched_steps_bit_exact_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "sched_steps.npz"))
    assert float(g["oracle_vs_reference"]) == 0.0
    ac = S.alphas_cumprod(100)
    n = 0
    for key in g.files:
        if not key.endswith(".prev"):
            continue
        kind, N, t, _ = key.split(".")
        N, t = int(N), int(t)
        cfg = S.SchedCfg(num_inference_steps=N)
        tag = f"{kind}/{N}/{t}"
        mo = 1.2 * W.hash_normal(tag + "/mo", (5, 16, 7))
        x = W.hash_normal(tag + "/x", (5, 16, 7))
        nz = W.hash_normal(tag + "/nz", (5, 16, 7))
        inp = W.synth_inputs(5, 0, 5)
        kw = dict(target_traj=inp["target_traj"], target_mask=inp["mask"], inpainting=True) if kind.startswith("inpainting") else {}
        fn = S.ddim_step if kind.endswith("ddim") else S.ddpm_step
        prev, x0 = fn(cfg, ac, mo, t, x, variance_noise=nz, **kw)
        assert np.array_equal(prev.numpy(), g[key]), key
        assert np.array_equal(x0.numpy(), g[key[:-5] + ".x0"]), key
        n += 1
    assert n == 32  # N=2 has only two distinct timesteps


def test_unet_forward_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_forward.npz"))
    inp = W.synth_inputs(3, 0, 31)
    t = torch.tensor([63, 5, 99])
    for mode in W.MODES:
        sd = W.make_state_dict(mode)
        y = U.unet_forward(sd, inp["x"], inp["feat"], t, inp["target"] if mode == "FREE_GUIDANCE" else None, mode)
        assert float((y - torch.from_numpy(g[mode])).abs().max()) <= 1e-5, mode


@pytest.mark.parametrize("name", ["cfg2_noguid_ddim10_b4", "cfg3_free_ddim10_b3", "cfg4_classifier_ddim2_b3",
                                  "cfg4b_inpaint_ddpm10_b2", "cfg3_free_ddpm10_b2", "cfg1_noguid_ddpm100_b1"])
def test_plan_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"plan_{name}.npz"))
    meta = json.loads(str(g["meta"]))
    assert float(g["oracle_vs_reference"]) <= 1e-4
    mode, kind, T, B, seed = PLAN_CASES[name]
    assert (meta["mode"], meta["scheduler"], meta["T"], meta["B"], meta["seed"]) == (mode, kind, T, B, seed)
    sd = W.make_state_dict(mode)
    assert W.state_dict_digest(sd) == meta["weights_digest"]
    inp = W.synth_inputs(B, T, seed)
    noise = inp["noise"] if (kind.endswith("ddpm") or kind.startswith("inpainting")) else None
    inpaint = kind.startswith("inpainting")
    out = P.plan(sd, mode, kind, inp["x"], inp["feat"], T, target=inp["target"] if mode != "NO_GUIDANCE" else None,
                 noise=noise, target_traj=inp["target_traj"] if inpaint else None, target_mask=inp["mask"] if inpaint else None)
    assert float((out - torch.from_numpy(g["trajs"])).abs().max()) <= 2e-4, name


def test_encoder_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_feature.npz"))
    sd = W.make_state_dict("NO_GUIDANCE")
    f = U.resnet34_feature(sd, W.synth_image(1, seed=2))
    assert float((f - torch.from_numpy(g["feat"])).abs().max()) <= 1e-3 * float(np.abs(g["feat"]).max())


@pytest.mark.skipif(not RL.available(), reason="reference tree not present (GPU box)")
def test_live_reference_hoisted_encoder_and_unet():
    sd = W.make_state_dict("NO_GUIDANCE")
    model = RL.build_reference_model("NO_GUIDANCE", sd)
    inp = W.synth_inputs(2, 0, 3)
    t = torch.tensor([42, 42])
    img = W.synth_image(1, seed=4, h=64, w=96).repeat(2, 1, 1, 1)
    with torch.no_grad():
        y_asis = model(inp["x"], img, t)                       # encoder inside the call, as the reference does
        feat = U.resnet34_feature(sd, img)
        y_hoisted = U.unet_forward(sd, inp["x"], feat, t)      # oracle with the feature hoisted
    assert float((y_asis - y_hoisted).abs().max()) <= 1e-5
