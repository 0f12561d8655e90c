// Plan post-processing -> vehicle control for a FLEET of vehicles: the step immediately after the
// sampling loop.  Replaces, one thread per vehicle,
//   * interact.py:296-297 + 218-229   post_process_control on the control triple of the first waypoint  (fleet_post_process_kernel)
//   * control/controller.py:29-76 + control/pid.py:16-28   waypoint-following PID with per-vehicle windows (fleet_pid_kernel)
// The reference runs a Python loop per vehicle over numpy scalars; here the per-vehicle PID windows (the deques of pid.py:10) live
// in device memory as ring buffers [n][V] (element i of all vehicles contiguous: coalesced), every vehicle advances in lock
// step, and all arithmetic is float64 like numpy's.  np.mean's summation order over the window (8 running sums, then the tail) is
// reproduced so that the integral term only differs from the host through libm (atan2).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b2p {

struct FleetState {            // header of the caller-provided state block, followed by the two windows
  long long tick;              // appends so far (== deque position): shared by the fleet, kept on the device so that a captured graph advances it
  long long pad;
};

// numpy's pairwise sum for n < 128 (numpy/core/src/umath/loops_utils.h.src): 8 accumulators over blocks of 8, then the tail
__device__ double np_mean_window(const double* win, int n, int V, int v, int oldest) {
  auto at = [&](int i) { int p = oldest + i; if (p >= n) p -= n; return win[(size_t)p * V + v]; };   // deque order: oldest first
  double res;
  if (n < 8) {
    res = 0.0;
    for (int i = 0; i < n; ++i) res += at(i);
  } else {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = at(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += at(i + j);
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += at(i);
  }
  return res / (double)n;
}

// PIDController.step (control/pid.py:16-28) on the ring buffer of vehicle v; `t` = appends before this one
__device__ double pid_step(double* win, int n, int V, int v, long long t, double kp, double ki, double kd, double error) {
  const int p = (int)(t % n);
  const int prev = p == 0 ? n - 1 : p - 1;
  const double last = win[(size_t)prev * V + v];          // window[-2] after the append (zero-initialised deque)
  win[(size_t)p * V + v] = error;
  double integral = 0.0, derivative = 0.0;
  if (n >= 2) {
    integral = np_mean_window(win, n, V, v, p + 1 == n ? 0 : p + 1);
    derivative = error - last;
  }
  return kp * error + ki * integral + kd * derivative;
}

__device__ __forceinline__ double heading(double x, double y) {   // control/controller.py:48-50
  return (M_PI / 2 - atan2(y, x)) * (180.0 / M_PI) / 90.0;
}
__device__ __forceinline__ double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

__global__ void __launch_bounds__(128) fleet_pid_kernel(b2p_control_config c, FleetState* st, double* turn_win, double* speed_win,
                                                        const float* __restrict__ wp, int N, const float* __restrict__ vel,
                                                        const float* __restrict__ tgt, float* __restrict__ out, int V) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const long long t = st->tick;
  const float* w = wp + (size_t)v * N * 2;
  // desired speed = mean segment length x 2; aim = the waypoint whose segment midpoint is closest to aim_dist (controller.py:34-46)
  double best_norm = 1e5, desired = 0.0, ax = w[0], ay = w[1];
  const int pairs = N - 1;
  for (int i = 0; i < pairs; ++i) {
    const double x0 = w[2 * i], y0 = w[2 * i + 1], x1 = w[2 * i + 2], y1 = w[2 * i + 3];
    const double dx = x1 - x0, dy = y1 - y0;
    desired += sqrt(dx * dx + dy * dy) * 2.0 / pairs;
    const double mx = (x1 + x0) / 2.0, my = (y1 + y0) / 2.0;
    const double norm = sqrt(mx * mx + my * my);
    if (fabs(c.aim_dist - best_norm) > fabs(c.aim_dist - norm)) { ax = x0; ay = y0; best_norm = norm; }
  }
  const double tx = tgt[2 * v], ty = tgt[2 * v + 1];
  const double angle = heading(ax, ay);
  const double angle_last = heading((double)w[2 * (N - 1)] - (double)w[2 * (N - 2)], (double)w[2 * (N - 1) + 1] - (double)w[2 * (N - 2) + 1]);
  const double angle_target = heading(tx, ty);
  bool use_target = fabs(angle_target) < fabs(angle);
  use_target = use_target || (fabs(angle_target - angle_last) > c.angle_thresh && ty < c.dist_thresh);
  const double steer = clipd(pid_step(turn_win, c.turn_n, V, v, t, c.turn_kp, c.turn_ki, c.turn_kd, use_target ? angle_target : angle), -1.0, 1.0);
  const double speed = vel[v];
  const bool brake = desired < c.brake_speed || (speed / desired) > c.brake_ratio;
  const double delta = clipd(desired - speed, 0.0, c.clip_delta);
  const double throttle = clipd(pid_step(speed_win, c.speed_n, V, v, t, c.speed_kp, c.speed_ki, c.speed_kd, delta), 0.0, c.max_throttle);
  out[3 * v + 0] = (float)(brake ? 0.0 : throttle);
  out[3 * v + 1] = (float)steer;
  out[3 * v + 2] = brake ? 1.f : 0.f;
}

__global__ void fleet_tick_kernel(FleetState* st) { st->tick += 1; }

// interact.py:296-297 + 218-229 on the first waypoint's last three columns (throttle, steer, brake)
__global__ void __launch_bounds__(128) fleet_post_process_kernel(const float* __restrict__ trajs, int HD, int D, float* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* c = trajs + (size_t)b * HD + (D - 3);
  float throttle = c[0], steer = c[1], brake = c[2];
  if (brake < 0.05f) brake = 0.f;
  if (throttle > brake) brake = 0.f;
  if (brake > 0.5f) { brake = 1.f; steer = 0.f; throttle = 0.f; }
  out[3 * b + 0] = throttle; out[3 * b + 1] = steer; out[3 * b + 2] = brake;
}

static bool cfg_ok(const b2p_control_config* c) { return c && c->turn_n >= 1 && c->turn_n <= 127 && c->speed_n >= 1 && c->speed_n <= 127; }

}  // namespace b2p

using namespace b2p;

extern "C" int64_t b2p_fleet_state_bytes(const b2p_control_config* c, int32_t V) {
  if (!cfg_ok(c) || V <= 0) return -1;
  return (int64_t)sizeof(FleetState) + (int64_t)sizeof(double) * V * (c->turn_n + c->speed_n);
}

extern "C" int b2p_fleet_reset(const b2p_control_config* c, void* state, int32_t V, void* stream) {
  if (!cfg_ok(c) || !state || V <= 0) return B2P_ERR_INVALID_ARG;
  B2P_CUDA_TRY(cudaMemsetAsync(state, 0, (size_t)b2p_fleet_state_bytes(c, V), (cudaStream_t)stream));   // deque([0] * n)
  return B2P_OK;
}

extern "C" int b2p_fleet_control_pid(const b2p_control_config* c, void* state, const float* waypoints, int32_t N, const float* velocity,
                                     const float* target, float* out, int32_t V, void* stream) {
  if (!cfg_ok(c) || !state || !waypoints || !velocity || !target || !out || V <= 0 || N < 2) return B2P_ERR_INVALID_ARG;
  FleetState* st = reinterpret_cast<FleetState*>(state);
  double* turn = reinterpret_cast<double*>(st + 1);
  double* speed = turn + (size_t)c->turn_n * V;
  cudaStream_t s = (cudaStream_t)stream;
  fleet_pid_kernel<<<(V + 127) / 128, 128, 0, s>>>(*c, st, turn, speed, waypoints, N, velocity, target, out, V);
  fleet_tick_kernel<<<1, 1, 0, s>>>(st);
  return (int)cudaGetLastError();
}

extern "C" int b2p_fleet_post_process(const float* trajs, float* out, int32_t B, int32_t H, int32_t D, void* stream) {
  if (!trajs || !out || B <= 0 || H < 1 || D < 3) return B2P_ERR_INVALID_ARG;
  fleet_post_process_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(trajs, H * D, D, out, B);
  return (int)cudaGetLastError();
}
