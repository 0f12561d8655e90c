"""Plan post-processing -> vehicle control: host-side mirror of the reference's controller.

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
    """Fleet form of interact.py:296-297 + 218-229: trajs [B,H,D>=5] (torch tensor, any device) -> [B,3] controls
    (throttle, steer, brake) read from the first waypoint's last three columns, one row per vehicle."""
    import torch

    c = trajs[:, 0, -3:].to(torch.float32)
    throttle, steer, brake = c[:, 0], c[:, 1], c[:, 2]
    brake = torch.where(brake < 0.05, torch.zeros_like(brake), brake)
    brake = torch.where(throttle > brake, torch.zeros_like(brake), brake)
    hard = brake > 0.5
    one, zero = torch.ones_like(brake), torch.zeros_like(brake)
    return torch.stack([torch.where(hard, zero, throttle), torch.where(hard, zero, steer), torch.where(hard, one, brake)], dim=-1)
