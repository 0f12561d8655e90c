"""Plan post-processing -> vehicle control.

Two forms: the host-side mirror of the reference's per-vehicle controller (``Controller``, ``PIDController``,
``post_process_control`` — numpy, pinned bit-for-bit against the reference's classes) and the FLEET form on the device
(``FleetController``, ``post_process_control_batch`` — hand-written kernels in csrc/control.cu, one thread per vehicle, the
per-vehicle PID windows resident in device memory) for many vehicles planned in one batch.

``post_process_control`` is what the shipped 7-dim configuration uses (interact.py:218-229, 296-297: the control triple is
the last three columns of the first waypoint); ``Controller.control_pid`` is the waypoint-following PID used when the model
only predicts (x, y) (interact.py:231-239, control/controller.py:29-76, control/pid.py:16-28).  Plain numpy, float64 like
the reference; one Controller per vehicle (the PID windows are per-vehicle state)."""
from __future__ import annotations

from collections import deque

import numpy as np


class PIDController:
    """control/pid.py:7-28."""

    def __init__(self, K_P=1.0, K_I=0.0, K_D=0.0, n=20):
        self._K_P, self._K_I, self._K_D = K_P, K_I, K_D
        self._window = deque([0 for _ in range(n)], maxlen=n)
        self._max = 0.0
        self._min = 0.0

    def step(self, error):
        self._window.append(error)
        self._max = max(self._max, abs(error))
        self._min = -abs(self._max)
        if len(self._window) >= 2:
            integral = np.mean(self._window)
            derivative = self._window[-1] - self._window[-2]
        else:
            integral, derivative = 0.0, 0.0
        return self._K_P * error + self._K_I * integral + self._K_D * derivative


def _heading(v):
    """Angle of a 2-vector measured from the +y axis, in units of 90 degrees (control/controller.py:48-50)."""
    return np.degrees(np.pi / 2 - np.arctan2(v[1], v[0])) / 90


class Controller:
    """control/controller.py:7-76 (same constructor argument: a config tree with PID.* and CONTROL.*)."""

    def __init__(self, cfg):
        p, c = cfg.PID, cfg.CONTROL
        self.turn_controller = PIDController(K_P=p.TURN_KP, K_I=p.TURN_KI, K_D=p.TURN_KD, n=p.TURN_N)
        self.speed_controller = PIDController(K_P=p.SPEED_KP, K_I=p.SPEED_KI, K_D=p.SPEED_KD, n=p.SPEED_N)
        self.aim_dist, self.angle_thresh, self.dist_thresh = c.AIM_DIST, c.ANGLE_THRESH, c.DIST_THRESH
        self.brake_speed, self.brake_ratio, self.clip_delta, self.max_throttle = c.BRAKE_SPEED, c.BRAKE_RATIO, c.CLIP_DELTA, c.MAX_THROTTLE

    def control_pid(self, waypoints, velocity, target):
        """waypoints [N,2], velocity [1], target [2] (torch tensors or arrays, ego frame) -> (throttle, steer, brake)."""
        to_np = lambda t: t.data.cpu().numpy() if hasattr(t, "data") and hasattr(t.data, "cpu") else np.asarray(t)  # noqa: E731
        waypoints, target = to_np(waypoints), to_np(target)
        num_pairs = len(waypoints) - 1
        best_norm, desired_speed, aim = 1e5, 0, waypoints[0]
        for i in range(num_pairs):
            # desired speed: mean segment length x 2; aim point: the waypoint whose segment midpoint is closest to aim_dist
            desired_speed += np.linalg.norm(waypoints[i + 1] - waypoints[i]) * 2.0 / num_pairs
            norm = np.linalg.norm((waypoints[i + 1] + waypoints[i]) / 2.0)
            if abs(self.aim_dist - best_norm) > abs(self.aim_dist - norm):
                aim, best_norm = waypoints[i], norm
        angle, angle_last, angle_target = _heading(aim), _heading(waypoints[-1] - waypoints[-2]), _heading(target)
        use_target_to_aim = np.abs(angle_target) < np.abs(angle)
        use_target_to_aim = use_target_to_aim or (np.abs(angle_target - angle_last) > self.angle_thresh and target[1] < self.dist_thresh)
        steer = np.clip(self.turn_controller.step(angle_target if use_target_to_aim else angle), -1.0, 1.0)
        speed = to_np(velocity[0])
        brake = desired_speed < self.brake_speed or (speed / desired_speed) > self.brake_ratio
        delta = np.clip(desired_speed - speed, 0.0, self.clip_delta)
        throttle = np.clip(self.speed_controller.step(delta), 0.0, self.max_throttle)
        return (throttle if not brake else 0.0), steer, brake


def post_process_control(throttle_res, steer_res, brake_res):
    """interact.py:218-229."""
    if brake_res < 0.05:
        brake_res = 0.0
    if throttle_res > brake_res:
        brake_res = 0.0
    if brake_res > 0.5:
        brake_res, steer_res, throttle_res = 1.0, 0.0, 0.0
    return np.array([throttle_res, steer_res, brake_res])


def post_process_control_batch(trajs):
    """Fleet form of interact.py:296-297 + 218-229: trajs [B,H,D>=5] (CUDA tensor) -> [B,3] controls (throttle, steer, brake)
    read from the first waypoint's last three columns, one row per vehicle (``b2p_fleet_post_process``)."""
    import ctypes as C

    import torch

    from . import _lib
    if trajs.device.type != "cuda":
        raise RuntimeError("post_process_control_batch runs on CUDA tensors (use post_process_control per vehicle on the host)")
    t = trajs.detach().to(torch.float32).contiguous()
    B, H, D = t.shape
    out = torch.empty((B, 3), device=t.device, dtype=torch.float32)
    if B:
        with torch.cuda.device(t.device):
            rc = _lib.load().b2p_fleet_post_process(_lib.ptr(t), _lib.ptr(out), B, H, D, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, None, "b2p_fleet_post_process")
    return out


class FleetController:
    """``Controller.control_pid`` (control/controller.py:29-76) for ``n_vehicles`` vehicles at once: one thread per vehicle,
    float64 arithmetic, the two PID windows of every vehicle (the deques of control/pid.py:10) resident on the device and
    advanced by every call.  Same constructor argument as ``Controller`` (a config tree with PID.* and CONTROL.*)."""

    def __init__(self, cfg, n_vehicles: int, device="cuda"):
        import ctypes as C

        import torch

        from . import _lib
        p, c = cfg.PID, cfg.CONTROL
        self._cfg = _lib.ControlConfig(p.TURN_KP, p.TURN_KI, p.TURN_KD, int(p.TURN_N), p.SPEED_KP, p.SPEED_KI, p.SPEED_KD, int(p.SPEED_N),
                                       c.AIM_DIST, c.ANGLE_THRESH, c.DIST_THRESH, c.BRAKE_SPEED, c.BRAKE_RATIO, c.CLIP_DELTA, c.MAX_THROTTLE)
        self.n_vehicles, self.device = int(n_vehicles), torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FleetController runs on CUDA (use Controller per vehicle on the host)")
        nbytes = _lib.load().b2p_fleet_state_bytes(C.byref(self._cfg), self.n_vehicles)
        if nbytes <= 0:
            raise ValueError("invalid PID configuration (window lengths must be in 1..127) or fleet size")
        self._state = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.reset()

    def _stream(self):
        import ctypes as C

        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self) -> None:
        import ctypes as C

        import torch

        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().b2p_fleet_reset(C.byref(self._cfg), _lib.ptr(self._state), self.n_vehicles, self._stream()), None, "b2p_fleet_reset")

    def control_pid(self, waypoints, velocity, target):
        """waypoints [V,N,2], velocity [V] or [V,1], target [V,2] (ego frame, metres) -> [V,3] = (throttle, steer, brake in {0,1})."""
        import ctypes as C

        import torch

        from . import _lib
        f = lambda t: t.detach().to(self.device, torch.float32).contiguous()  # noqa: E731
        w, v, tg = f(waypoints), f(velocity).reshape(-1), f(target).reshape(-1, 2)
        V, N = w.shape[0], w.shape[1]
        if V != self.n_vehicles or v.numel() != V or tg.shape[0] != V or w.dim() != 3 or w.shape[2] != 2 or N < 2:
            raise ValueError(f"expected waypoints [{self.n_vehicles},N>=2,2], velocity [{self.n_vehicles}], target [{self.n_vehicles},2]")
        out = torch.empty((V, 3), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            rc = _lib.load().b2p_fleet_control_pid(C.byref(self._cfg), _lib.ptr(self._state), _lib.ptr(w), N, _lib.ptr(v), _lib.ptr(tg),
                                                   _lib.ptr(out), V, self._stream())
        _lib.check(rc, None, "b2p_fleet_control_pid")
        return out
