"""ctypes binding of libb200plan.so (C ABI: include/b200plan.h).  There is NO fallback: if the library is missing or
a call fails this module raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_LOCK = threading.Lock()
_LIB = None

c_float_p = C.POINTER(C.c_float)
c_int64_p = C.POINTER(C.c_int64)


class ModelConfig(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("transition_dim", C.c_int32), ("dim", C.c_int32), ("n_mults", C.c_int32),
                ("dim_mults", C.c_int32 * 8), ("guidance", C.c_int32), ("precision", C.c_int32)]


class SchedConfig(C.Structure):
    _fields_ = [("kind", C.c_int32), ("num_train_timesteps", C.c_int32), ("prediction_type", C.c_int32),
                ("thresholding", C.c_int32), ("clip_sample", C.c_int32), ("clip_sample_range", C.c_float),
                ("dynamic_thresholding_ratio", C.c_float), ("sample_max_value", C.c_float),
                ("beta_schedule", C.c_int32), ("beta_start", C.c_float), ("beta_end", C.c_float)]


class StepCoeffs(C.Structure):
    _fields_ = [("t", C.c_int32), ("t_prev", C.c_int32), ("alpha_prod_t", C.c_float), ("alpha_prod_t_prev", C.c_float),
                ("sqrt_alpha_prod_t", C.c_float), ("sqrt_beta_prod_t", C.c_float), ("sqrt_alpha_prod_t_prev", C.c_float),
                ("sqrt_one_minus_alpha_prod_t_prev", C.c_float), ("variance", C.c_float), ("std_dev_t", C.c_float),
                ("dir_coeff", C.c_float), ("x0_coeff", C.c_float), ("sample_coeff", C.c_float),
                ("guidance_grad_scale", C.c_float)]


class PlanConfig(C.Structure):
    _fields_ = [("sched", SchedConfig), ("num_inference_steps", C.c_int32), ("eta", C.c_float), ("free_scale", C.c_float),
                ("classifier_scale", C.c_float), ("magic_num", C.c_float), ("postprocess", C.c_int32), ("use_graph", C.c_int32)]


class ControlConfig(C.Structure):
    _fields_ = [("turn_kp", C.c_double), ("turn_ki", C.c_double), ("turn_kd", C.c_double), ("turn_n", C.c_int32),
                ("speed_kp", C.c_double), ("speed_ki", C.c_double), ("speed_kd", C.c_double), ("speed_n", C.c_int32),
                ("aim_dist", C.c_double), ("angle_thresh", C.c_double), ("dist_thresh", C.c_double), ("brake_speed", C.c_double),
                ("brake_ratio", C.c_double), ("clip_delta", C.c_double), ("max_throttle", C.c_double)]


ABI_VERSION = 8
SCHED_KINDS = {"guidance_ddim": 0, "guidance_ddpm": 1, "inpainting_ddim": 2, "inpainting_ddpm": 3}
PRED_TYPES = {"epsilon": 0, "sample": 1, "v_prediction": 2}
BETA_SCHEDULES = {"squaredcos_cap_v2": 0, "linear": 1, "scaled_linear": 2}
GUIDANCE = {"NO_GUIDANCE": 0, "FREE_GUIDANCE": 1, "CLASSIFIER_GUIDANCE": 2}
PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2}
STEP_ZERO_FIRST_WAYPOINT, STEP_FINAL_POSTPROCESS, STEP_USE_CLIPPED_OUTPUT = 1, 2, 4

# every symbol include/b200plan.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "b2p_abi_version": (C.c_int, []),
    "b2p_status_string": (C.c_char_p, [C.c_int]),
    "b2p_last_error": (C.c_char_p, [_VP]),
    "b2p_create": (C.c_int, [C.POINTER(ModelConfig), C.c_int, C.POINTER(_VP)]),
    "b2p_destroy": (C.c_int, [_VP]),
    "b2p_load_weight": (C.c_int, [_VP, C.c_char_p, _VP, C.c_int64]),
    "b2p_num_weights": (C.c_int, [_VP]),
    "b2p_weight_info": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_char_p), c_int64_p]),
    "b2p_finalize_weights": (C.c_int, [_VP]),
    "b2p_set_precision": (C.c_int, [_VP, C.c_int]),
    "b2p_set_small_batch_max": (C.c_int, [_VP, C.c_int]),
    "b2p_set_chain": (C.c_int, [_VP, C.c_int]),
    "b2p_preprocess_frames": (C.c_int, [_VP, _VP, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), _VP]),
    "b2p_encoder_stem_bf16": (C.c_int, [_VP, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP]),
    "b2p_encoder_stem_pool_bf16": (C.c_int, [_VP, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP]),
    "b2p_maxpool3x3s2_nhwc_bf16": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "b2p_encoder_conv_bf16": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "b2p_unet_forward": (C.c_int, [_VP, _VP, _VP, C.c_int32, _VP, C.c_int32, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "b2p_state_pred": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "b2p_state_pred_vjp": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "b2p_classifier_guidance": (C.c_int, [_VP, _VP, _VP, _VP, C.c_float, C.c_float, C.c_int32, _VP]),
    "b2p_alphas_cumprod": (C.c_int, [C.c_char_p, C.c_int32, C.c_float, C.c_float, c_float_p]),
    "b2p_timesteps": (C.c_int, [C.c_int32, C.c_int32, c_int64_p]),
    "b2p_step_coeffs_compute": (C.c_int, [C.POINTER(SchedConfig), c_float_p, C.c_int32, C.c_int32, C.c_float, C.POINTER(StepCoeffs)]),
    "b2p_sched_step": (C.c_int, [C.POINTER(SchedConfig), C.POINTER(StepCoeffs), _VP, _VP, C.c_float, _VP, _VP, _VP, _VP, _VP, _VP,
                                 C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, _VP]),
    "b2p_plan": (C.c_int, [_VP, C.POINTER(PlanConfig), _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "b2p_plan_host": (C.c_int, [_VP, C.POINTER(PlanConfig), _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32]),
    "b2p_plan_host_async": (C.c_int, [_VP, C.POINTER(PlanConfig), _VP, _VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, C.c_int32]),
    "b2p_sync": (C.c_int, [_VP]),
    "b2p_plan_sharded_host": (C.c_int, [C.POINTER(_VP), C.c_int32, C.POINTER(PlanConfig), _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32]),
    "b2p_set_noise_seed": (C.c_int, [_VP, C.c_uint64]),
    "b2p_last_noise_key": (C.c_uint64, [_VP]),
    "b2p_philox_normal": (C.c_int, [C.c_uint64, C.c_int32, C.c_int64, _VP, _VP]),
    "b2p_fleet_state_bytes": (C.c_int64, [C.POINTER(ControlConfig), C.c_int32]),
    "b2p_fleet_reset": (C.c_int, [C.POINTER(ControlConfig), _VP, C.c_int32, _VP]),
    "b2p_fleet_control_pid": (C.c_int, [C.POINTER(ControlConfig), _VP, _VP, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP]),
    "b2p_fleet_post_process": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "b2p_last_launch_count": (C.c_int64, [_VP]),
    "b2p_unet_flops_per_sample": (C.c_int64, [_VP]),
    "b2p_weight_bytes": (C.c_int64, [_VP]),
}


class B2PError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen libb200plan.so (must have been built in-tree: `python -m autonomous_driving_with_diffusion_model_b200.build`
    or `__graft_entry__.build()`).  Raises if absent — there is no Python/CPU fallback."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.exists(path):
            raise B2PError(f"{path} not found: build the CUDA library first (python -m autonomous_driving_with_diffusion_model_b200.build); "
                           "there is no CPU fallback")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.b2p_abi_version() != ABI_VERSION:
            raise B2PError(f"ABI mismatch: library {lib.b2p_abi_version()} vs binding {ABI_VERSION}; rebuild")
        _LIB = lib
        return lib


def check(rc: int, handle=None, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.b2p_status_string(rc).decode()
    detail = lib.b2p_last_error(handle).decode() if handle else ""
    text = f"{what}: {msg}" + (f" ({detail})" if detail else "") + f" [rc={rc}]"
    if rc in (-1, -3):
        raise ValueError(text)
    if rc == -2:
        raise KeyError(text)
    raise B2PError(text)


def ptr(t):
    """device/host address of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
