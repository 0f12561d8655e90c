"""CPU tests of the C-ABI boundary: the library builds/loads, exports every symbol include/b200plan.h declares, and the
host-only entry points (schedule tables, coefficients, error strings) behave.  No compute calls: there is no GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import autonomous_driving_with_diffusion_model_b200 as P
from autonomous_driving_with_diffusion_model_b200 import _lib, build
from oracle import schedulers as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_all_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "b200plan.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|uint64_t|const char\*)\s+(b2p_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.b2p_abi_version() == int(re.search(r"#define B2P_ABI_VERSION (\d+)", hdr).group(1))


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.ModelConfig) == 4 * (4 + 8 + 2)
    assert C.sizeof(_lib.SchedConfig) == 4 * 11
    assert C.sizeof(_lib.StepCoeffs) == 4 * 14
    assert C.sizeof(_lib.PlanConfig) == C.sizeof(_lib.SchedConfig) + 4 * 7


def test_status_strings_and_null_handle(lib):
    assert lib.b2p_status_string(0) == b"ok"
    assert b"invalid" in lib.b2p_status_string(-1)
    assert lib.b2p_last_error(None) == b""
    assert lib.b2p_num_weights(None) == 0
    assert lib.b2p_destroy(None) == 0


def test_create_without_gpu_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.ModelConfig(16, 7, 64, 4, (C.c_int32 * 8)(1, 2, 4, 8), 0, 0)
    h = C.c_void_p()
    rc = lib.b2p_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -6 and not h.value  # B2P_ERR_NO_DEVICE: no CPU fallback
    with pytest.raises(_lib.B2PError):
        _lib.check(rc, None, "b2p_create")


def test_schedule_tables_match_oracle(lib):
    for n in (100, 10):
        ac = np.empty(n, np.float32)
        assert lib.b2p_alphas_cumprod(b"squaredcos_cap_v2", n, 0.0, 0.0, ac.ctypes.data_as(_lib.c_float_p)) == 0
        assert np.array_equal(ac, S.alphas_cumprod(n).numpy())
    assert lib.b2p_alphas_cumprod(b"nope", 10, 0.0, 0.0, ac.ctypes.data_as(_lib.c_float_p)) == -1
    for n_inf in (100, 10, 2, 7):
        ts = np.empty(n_inf, np.int64)
        assert lib.b2p_timesteps(100, n_inf, ts.ctypes.data_as(_lib.c_int64_p)) == 0
        assert list(ts) == list(S.leading_timesteps(100, n_inf))
    assert lib.b2p_timesteps(100, 101, ts.ctypes.data_as(_lib.c_int64_p)) == -1


def test_step_coeffs_c_side_within_one_ulp_of_reference_order_torch_math(lib):
    cfg = P.load_cfg()
    for cls in (P.GuidanceDDIMScheduler, P.GuidanceDDPMScheduler):
        s = cls(cfg=cfg, **P.scheduler_kwargs(cfg))
        for n in (100, 10, 2):
            s.set_timesteps(n)
            sc, kc = s.sched_config(), _lib.StepCoeffs()
            ac = s.alphas_cumprod.numpy().copy()
            for t in s.timesteps.tolist():
                assert lib.b2p_step_coeffs_compute(C.byref(sc), ac.ctypes.data_as(_lib.c_float_p), n, t, 0.0, C.byref(kc)) == 0
                k = s.coeffs(t)
                for f, _ in _lib.StepCoeffs._fields_[2:]:
                    a, b = float(getattr(k, f)), float(getattr(kc, f))
                    assert abs(a - b) <= 2.5e-7 * max(abs(a), 1e-30), (f, t, a, b)


def test_scheduler_surface_matches_reference_contract():
    cfg = P.load_cfg()
    s = P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg))
    assert s.num_inference_steps is None
    with pytest.raises(ValueError, match="Number of inference steps is 'None'"):
        s.step(torch.zeros(1, 16, 7), 0, torch.zeros(1, 16, 7))
    with pytest.raises(ValueError, match="cannot be larger"):
        s.set_timesteps(101)
    s.set_timesteps(10)
    assert s.timesteps.tolist() == [90, 80, 70, 60, 50, 40, 30, 20, 10, 0] and s.timesteps.dtype == torch.int64
    assert s.config.prediction_type == "sample" and s.config.thresholding is True and s.config.sample_max_value == 1.0
    assert torch.equal(s.alphas_cumprod, S.alphas_cumprod(100))
    # known-answer first-step variances
    assert abs(s.coeffs(90).variance - 0.71890974) < 1e-6
    d = P.GuidanceDDPMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg))
    d.set_timesteps(100)
    assert d.coeffs(0).variance == pytest.approx(1e-20)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.step(torch.zeros(1, 16, 7), 90, torch.zeros(1, 16, 7))  # CPU tensors: no fallback


def test_model_contract_keys_and_parameter_order(golden_dir):
    import json
    spec = json.load(open(os.path.join(golden_dir, "state_dict_spec.json")))
    for mode in ("NO_GUIDANCE", "FREE_GUIDANCE", "CLASSIFIER_GUIDANCE"):
        cfg = P.load_cfg(TRAIN=dict(USE_COND=mode))
        m = P.build_model(cfg)
        assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == spec[mode]["state_dict"]
        assert [n for n, _ in m.named_parameters()] == spec[mode]["parameters"]   # positional EMA copy (misc/load_param.py:4-8)
        assert m.magic_num == 23.315 and m.use_cond == P.GuidanceType[mode]
        assert hasattr(m, "perception") and (mode != "CLASSIFIER_GUIDANCE" or hasattr(m, "state_pred"))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16, 7), torch.zeros(1, 64), torch.tensor([3]))


def test_config_loader_follows_base_chain(tmp_path):
    (tmp_path / "default.yaml").write_text("PROJECT_DIR: x\nTRAIN:\n  ROOT: data\n")
    g = tmp_path / "guidance"
    g.mkdir()
    (g / "free.yaml").write_text("_BASE_: ../default.yaml\nTRAIN:\n  USE_COND: FREE_GUIDANCE\nGUIDANCE:\n  USE_COND: FREE_GUIDANCE\n  FREE_SCALE: 7.5\nEVAL:\n  SAMPLE_STEPS: 10\n")
    cfg = P.load_cfg(str(g / "free.yaml"))
    assert cfg.TRAIN.USE_COND == "FREE_GUIDANCE" and cfg.GUIDANCE.FREE_SCALE == 7.5 and cfg.EVAL.SAMPLE_STEPS == 10
    assert cfg.MODEL.HORIZON == 16 and cfg.TRAIN.ROOT == "data" and cfg.EVAL.SCHEDULER == "ddim"


def test_shard_bounds_cover_batch():
    for B in (1, 7, 256, 4096):
        for ws in (1, 2, 4, 8):
            b = P.shard_bounds(B, ws)
            assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_preprocess_frames_argument_checks_without_a_gpu():
    """b2p_preprocess_frames validates its arguments before touching the device (n_pixels, NULLs, alignment, zero std)."""
    import ctypes as C
    from autonomous_driving_with_diffusion_model_b200 import _lib
    lib = _lib.load()
    mean, std, bad = (C.c_float * 3)(0.5, 0.5, 0.5), (C.c_float * 3)(1, 1, 1), (C.c_float * 3)(1, 0, 1)
    assert lib.b2p_preprocess_frames(None, None, 0, mean, std, None) == 0            # empty batch: nothing to do
    assert lib.b2p_preprocess_frames(None, None, -1, mean, std, None) != 0
    assert lib.b2p_preprocess_frames(None, None, 8, mean, std, None) != 0
    assert lib.b2p_preprocess_frames(C.c_void_p(4098), C.c_void_p(8192), 8, mean, std, None) != 0   # input not 4-byte aligned
    assert lib.b2p_preprocess_frames(C.c_void_p(4096), C.c_void_p(8200), 8, mean, std, None) != 0   # output not 16-byte aligned
    assert lib.b2p_preprocess_frames(C.c_void_p(4096), C.c_void_p(8192), 8, mean, bad, None) != 0   # zero std
    assert lib.b2p_set_small_batch_max(None, 4) != 0


def test_schedule_tables_property(lib):
    """Random (num_train_timesteps, schedule, beta range, inference steps): the C host tables equal the oracle's (which follows
    diffusers 0.28.0: scheduling_ddim.py betas / set_timesteps 'leading')."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(2, 400), sched=st.sampled_from(["squaredcos_cap_v2", "linear", "scaled_linear"]),
           b0=st.floats(1e-5, 5e-3), span=st.floats(1e-3, 5e-2), frac=st.floats(0.01, 1.0))
    def check(n, sched, b0, span, frac):
        b1 = b0 + span
        ac = np.empty(n, np.float32)
        assert lib.b2p_alphas_cumprod(sched.encode(), n, b0, b1, ac.ctypes.data_as(_lib.c_float_p)) == 0
        want = S.alphas_cumprod(n, sched, b0, b1).numpy()
        if sched == "squaredcos_cap_v2":
            assert np.array_equal(ac, want), (n, float(np.abs(ac - want).max()))      # the shipped schedule: Python float math, exact
        else:
            # torch.linspace's CPU kernel is vectorised (base + step * lane, width depends on the host's ISA), so its last bit is
            # machine dependent; the C table uses the scalar formula.  The Python scheduler classes take their betas from torch.
            # A last-bit difference in a beta then drifts through the fp32 cumulative product (n <= 400 factors): bound 2e-6 relative.
            assert np.allclose(ac, want, rtol=2e-6, atol=0), (n, sched, b0, b1, float(np.abs(ac - want).max()))
        n_inf = max(1, min(n, int(round(frac * n))))
        ts = np.empty(n_inf, np.int64)
        assert lib.b2p_timesteps(n, n_inf, ts.ctypes.data_as(_lib.c_int64_p)) == 0
        assert list(ts) == list(S.leading_timesteps(n, n_inf))
        assert all(0 <= t < n for t in ts) and all(ts[i] > ts[i + 1] for i in range(n_inf - 1)) or n_inf == 1 or len(set(ts)) < n_inf

    check()
