// abrb_launch.hpp — host-side launch interface between api.cu and the per-joint-count kernel TUs.
#pragma once
#include <cuda_runtime.h>

#include "abrb_host.hpp"

namespace abrb {

struct RbdCall {
  int frame;
  const double *xoff;  // host, 3 values or nullptr
  const void *q, *dq;  // device
  int64_t B;
  abrb_rbd_out out;    // device pointers
  bool f32;
  cudaStream_t stream;
};

// Fused all-gather of the control outputs over NVLink peer memory (BASELINE config 5): the OSC kernel's epilogue stores
// every rank's rows straight into the gathered (B_total, n) array of EVERY rank (buffers mapped with CUDA IPC), so the
// exchange overlaps the arithmetic tile by tile instead of following it as a separate collective.
struct GatherArgs {
  void *peer_u[kMaxPeers];                  // base of the gathered array on each rank (own rank included)
  unsigned long long *peer_flag[kMaxPeers]; // &flags[my_rank] on each rank: receives `epoch` when all my rows are there
  int n_peer = 0;
  int self = 0;                             // this rank's index in peer_u / peer_flag
  int64_t row0 = 0;                         // first row of this rank's block in the gathered array
  unsigned long long epoch = 0;
  unsigned *cta_counter = nullptr;          // local: CTAs of this launch that have finished
};

struct OscCall {
  int frame;
  const double *xoff;
  const void *q, *dq, *target, *tv;  // device
  int target_stride, tv_stride;
  void *u, *train;
  int64_t B;
  bool f32;
  cudaStream_t stream;
  void *ierr = nullptr;  // (B, 6) integrated task-space error, in/out (device), only with ki != 0
  const GatherArgs *gather = nullptr;
  int *sched = nullptr;  // device: {next tile, CTAs done}, zero between launches (api.cu, sched_slot), or nullptr
};

struct RolloutCall {
  int frame;
  const double *xoff;
  void *q, *dq;  // device, in/out
  const void *target;
  int target_stride;
  int steps;
  double dt;
  void *q_traj, *dq_traj, *u_traj;
  int64_t B;
  bool f32;
  cudaStream_t stream;
  void *ierr = nullptr;  // (B, 6) integrated task-space error, in/out (device), only with ki != 0
};

struct NullCall {
  const void *q, *dq;
  void *u;
  int64_t B;
  bool f32;
  cudaStream_t stream;
};

// Joint (kind 0) and Floating (kind 1) controllers
struct CtrlCall {
  int kind;
  double kp, kv;
  int flag_a, flag_b;              // Joint: account_for_gravity, -   Floating: task_space, dynamic
  const void *q, *dq, *target, *tv;
  int target_stride, tv_stride;
  void *u;
  int64_t B;
  bool f32;
  cudaStream_t stream;
};

struct SlidingCall {
  double kd, lamb;
  int cartesian, frame;
  const double *xoff;  // host, 3 values or nullptr
  const void *q, *dq, *target, *tv, *ta;
  int target_stride, tv_stride, ta_stride;
  void *u, *s;
  int64_t B;
  bool f32;
  cudaStream_t stream;
};

struct IkCall {
  double max_dx, max_dr, max_dq, dt;
  int method, steps;
  const void *position, *target;
  int target_stride;
  void *pos_path, *vel_path;
  int64_t B;
  bool f32;
  cudaStream_t stream;
};

// Each returns a cudaError_t (0 = success).  Defined once per joint count in kernels.cu (-DABRB_N=<n>).
template <int N> int launch_rbd(const ChainHost &h, const RbdCall &c);
template <int N> int launch_osc(const ChainHost &h, const abrb_osc_params &p, const OscCall &c);
template <int N> int launch_rollout(const ChainHost &h, const abrb_osc_params &p, const RolloutCall &c);
template <int N> int launch_null(const ChainHost &h, const abrb_null_params &z, const NullCall &c);
template <int N> int launch_ctrl(const ChainHost &h, const CtrlCall &c);
template <int N> int launch_sliding(const ChainHost &h, const SlidingCall &c);
template <int N> int launch_ik(const ChainHost &h, const IkCall &c);

void count_launch();

}  // namespace abrb
