/*
 * abrb.h — C ABI of libabrb.so: batched rigid-body quantities and operational-space control
 *          for serial robot arms on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of abr/abr_control.  Each entry
 * point states the reference interface it replaces (paths relative to the reference checkout).
 * In the reference the inner boundary is a run-time generated Cython function
 *      def autofunc_c(double q0, ..., [double dq0, ...], [double x, double y, double z]) -> ndarray
 * wrapping  void autofunc(double q0, ..., double *out)   (emitted by
 * abr_control/arms/base_config.py:125-146, loaded at :148-201), called once per joint state.
 * Here one call evaluates a whole batch of B joint states on the GPU.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (`stream` is a cudaStream_t passed as void*,
 *     NULL = the legacy default stream);
 *   - unless the name contains `_host`, every data pointer is a DEVICE pointer on the current CUDA device,
 *     aligned to its element type (any row of a contiguous array is a valid start), row-major, batch-major: q is (B, n_joints), J is (B, 6, n_joints), M is (B, n, n) ...
 *     exactly the per-state shapes the reference returns, stacked;
 *   - the library never allocates or frees caller buffers; device-pointer calls are asynchronous with
 *     respect to the host (enqueued on `stream`); `_host` calls copy in, run, copy out and synchronise;
 *   - every function returns 0 (ABRB_OK) or a negative ABRB_E* code; abrb_last_error() gives the
 *     message for the calling thread;
 *   - handles are immutable after creation, so concurrent calls on different streams are allowed;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with ABRB_ECUDA.
 */
#ifndef ABRB_H_
#define ABRB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABRB_VERSION 200 /* 0.2.0: integrated_error arguments, asynchronous host slots, peer-gather epilogue */
#define ABRB_MAX_JOINTS 7
#define ABRB_MAX_NULL 4
#define ABRB_MAX_OBSTACLES 16

enum {
  ABRB_OK = 0,
  ABRB_EINVAL = -1,   /* bad argument (NULL handle, B < 0, misaligned pointer, ...) */
  ABRB_EFRAME = -2,   /* reference: Exception("Invalid transformation name: ...") (arms/ur5/config.py:336-337) */
  ABRB_ESHAPE = -3,   /* n_joints / n_links outside what the library was built for */
  ABRB_EUNSUP = -4,   /* reference: NotImplementedError / Exception("Invalid algorithm number") (controllers/osc.py:190-194) */
  ABRB_ECUDA = -5,    /* CUDA runtime error (no device, launch failure, ...) */
  ABRB_ENOMEM = -6
};

/* ---------------------------------------------------------------------------------------------------
 * Arm model.  Replaces the per-arm `Config` classes' model data (arms/ur5/config.py:52-339,
 * arms/jaco2/config.py:56-356, arms/threejoint/config.py:45-223, arms/twojoint/config.py:30-181) in the
 * flattened chain form of SURVEY.md Appendix A.1:
 *     link0 = L0,  joint_i = link_i.A[i],  link_{i+1} = joint_i.Rz(q_i).B[i],  EE = link_n.E
 * All matrices are 3x4 row-major [R|t] blocks of the 4x4 homogeneous factors.  The rotation blocks are
 * NOT required to be orthonormal (Jaco2's measured frames are not).
 * ------------------------------------------------------------------------------------------------- */
typedef struct abrb_chain_desc {
  int32_t n_joints;                              /* 1 .. ABRB_MAX_JOINTS */
  int32_t n_links;                               /* must equal n_joints + 1 (link0 .. link_n) */
  double L0[12];
  double A[ABRB_MAX_JOINTS][12];
  double B[ABRB_MAX_JOINTS][12];
  double E[12];
  double link_inertia[ABRB_MAX_JOINTS + 1][6];   /* diag(m,m,m,Ixx,Iyy,Izz) of `_M_LINKS[l]` (base_config.py:625-632) */
  double gravity[6];                             /* `self.gravity`, base_config.py:123: [0,0,-9.81,0,0,0] */
} abrb_chain_desc;

typedef struct abrb_model abrb_model;

int abrb_version(void);
const char *abrb_last_error(void);
/* number of CUDA devices visible, or ABRB_ECUDA */
int abrb_device_count(void);

int abrb_model_create(const abrb_chain_desc *desc, abrb_model **out);
int abrb_model_destroy(abrb_model *m);
int abrb_model_n_joints(const abrb_model *m);
/* 1 if every constant rotation block is orthonormal to 1e-12 (selects the cross-product kernels) */
int abrb_model_is_orthonormal(const abrb_model *m);
/* Frame name -> id.  Names as in the reference: "link0".."link<n>", "joint0".."joint<n-1>", "EE".
 * Unknown names return ABRB_EFRAME (reference raises Exception, arms/ur5/config.py:336-337). */
int abrb_frame_id(const abrb_model *m, const char *name);

/* ---------------------------------------------------------------------------------------------------
 * Batched rigid-body quantities.  One call replaces, for B states at once, the reference methods
 *   Tx (base_config.py:371-392 / :739-789)     T (:338-369)          R (:287-301 / :647-676)
 *   T_inv (:394-415 / :791-837)                quaternion (:304-318) J (:249-270 / :522-592)
 *   dJ (:225-247 / :470-520)                   M (:272-285 / :594-645)
 *   g (:210-223 / :417-468)                    C (:320-336 / :678-727)
 * A NULL output pointer means "not wanted".  Frame-dependent outputs (Tx, T, R, T_inv, quat, J, dJ) are
 * for frame `frame_id` and the point `x_off` (3 host doubles, NULL = origin) inside that frame; M, g, C
 * do not depend on them.  `dq` may be NULL unless dJ or C is requested.
 * Shapes per state: Tx[3]  T[4][4]  R[3][3]  T_inv[4][4]  quat[4] (w,x,y,z)  J[6][n]  dJ[6][n]  M[n][n]
 * g[n]  C[n][n].   Unlike the reference's public wrappers nothing is rounded to float32 in the f64 variant.
 * ------------------------------------------------------------------------------------------------- */
typedef struct abrb_rbd_out {
  void *Tx, *T, *R, *T_inv, *quat, *J, *dJ, *M, *g, *C;
} abrb_rbd_out;

int abrb_rbd_eval_f64(const abrb_model *m, int frame_id, const double *x_off, const double *q,
                      const double *dq, int64_t B, const abrb_rbd_out *out, void *stream);
int abrb_rbd_eval_f32(const abrb_model *m, int frame_id, const double *x_off, const float *q,
                      const float *dq, int64_t B, const abrb_rbd_out *out, void *stream);
/* same, q/dq/outputs are HOST pointers; H2D + kernel + D2H + sync inside the call */
int abrb_rbd_eval_host_f64(const abrb_model *m, int frame_id, const double *x_off, const double *q,
                           const double *dq, int64_t B, const abrb_rbd_out *out);
int abrb_rbd_eval_host_f32(const abrb_model *m, int frame_id, const double *x_off, const float *q,
                           const float *dq, int64_t B, const abrb_rbd_out *out);

/* ---------------------------------------------------------------------------------------------------
 * Operational-space controller.  abrb_osc_params mirrors the constructor of
 * abr_control.controllers.OSC (controllers/osc.py:53-118) plus its secondary ("null space") controllers:
 *   ABRB_NULL_DAMPING        controllers/damping.py:21-32          M (-kv dq)
 *   ABRB_NULL_RESTING        controllers/resting_config.py:25-42 + controllers/joint.py:104-131
 *   ABRB_NULL_AVOID          controllers/avoid_obstacles.py:38-120
 * abrb_osc_generate_* evaluates OSC.generate (controllers/osc.py:217-320) for B states.
 * ------------------------------------------------------------------------------------------------- */
enum { ABRB_NULL_DAMPING = 1, ABRB_NULL_RESTING = 2, ABRB_NULL_AVOID = 3, ABRB_NULL_JOINT_LIMITS = 4 };

typedef struct abrb_null_params {
  int32_t kind;
  int32_t n_obstacles;                            /* AVOID */
  double kp, kv;                                  /* DAMPING: kv; RESTING: kp, kv (Joint: kv default sqrt(kp)) */
  double rest_angles[ABRB_MAX_JOINTS];            /* RESTING */
  int32_t rest_mask[ABRB_MAX_JOINTS];             /* RESTING: 0 where the reference has None */
  int32_t _pad;
  double threshold, gain, maximum;                /* AVOID (avoid_obstacles.py:25-36) */
  double obstacles[ABRB_MAX_OBSTACLES][4];        /* AVOID: x, y, z, radius */
  /* JOINT_LIMITS (avoid_joint_limits.py:36-86): the controller's attributes AFTER its constructor, i.e. limits
   * shifted by -pi and swapped where cross_zero is set; NaN = no limit on that side; max_torque default 1 */
  double limit_min[ABRB_MAX_JOINTS], limit_max[ABRB_MAX_JOINTS], limit_torque[ABRB_MAX_JOINTS];
  int32_t limit_cross_zero[ABRB_MAX_JOINTS], limit_gradient[ABRB_MAX_JOINTS];
} abrb_null_params;

typedef struct abrb_osc_params {
  double kp, ko, kv, ki;                          /* resolved gains: caller applies the ko/kv defaults (osc.py:71-74) */
  double vmax[2];                                 /* used if use_vmax */
  double mx_threshold;                            /* `_Mx(threshold=1e-3)`, osc.py:120 */
  int32_t use_vmax;
  int32_t ctrlr_dof[6];
  int32_t use_g, use_C;
  int32_t orientation_algorithm;                  /* 0 or 1, else ABRB_EUNSUP */
  int32_t n_null;
  int32_t _pad;
  abrb_null_params null[ABRB_MAX_NULL];
} abrb_osc_params;

typedef struct abrb_osc abrb_osc;

int abrb_osc_create(const abrb_model *m, const abrb_osc_params *p, abrb_osc **out);
int abrb_osc_destroy(abrb_osc *c);

/* Execution options of one controller (no effect on results).  Names:
 *   "host_chunk_states"  states per pipeline chunk of the *_host entry points; 0 = automatic (default: the blocking
 *                        calls split batches of 49 152 states and more into two chunks, four above 196 608; the
 *                        asynchronous calls, which overlap whole calls on the two slots, keep one chunk up to 196 608
 *                        states; the environment variable ABRB_HOST_CHUNK sets another default).
 *   "host_upload_streams" copy streams per chunk of the *_host entry points: 1 (q, dq, target one after the other), 2
 *                        (default: dq beside q) or 3 (per-state targets on a stream of their own as well); default from
 *                        ABRB_HOST_STREAMS.  Which is fastest depends on the host: one pinned host->device stream
 *                        reaches 18 to 54 GB/s on different B200 boxes of the same pool.
 * Returns ABRB_EINVAL for an unknown name.  Not thread safe against concurrent generate calls on the same handle. */
int abrb_osc_set_option(abrb_osc *c, const char *name, double value);

/* OSC.generate(q, dq, target, target_velocity=None, ref_frame="EE", xyz_offset=None) for B states.
 *   target           (B,6) if target_stride == 6, or one (6,) row broadcast to all states if target_stride == 0
 *   target_velocity  NULL (the reference's `np.all(target_velocity == 0)` joint-space damping branch,
 *                    osc.py:275-278) or (B,6)/(6,) per tv_stride (task-space branch, osc.py:279-282)
 *   u                (B,n) out;   training_signal (B,n) out or NULL (osc.py:297)
 *   integrated_error (B,6) in/out: the reference's per-controller `self.integrated_error` (osc.py:81-82), one row per
 *                    state, updated as osc.py:262-264 (`integrated_error += u_task; u_task += ki * integrated_error`).
 *                    Must be given if and only if the controller was created with ki != 0 (else ABRB_EINVAL); the
 *                    caller owns it and zeroes it to reset the integrator.
 *   frame_id/x_off   ref_frame and xyz_offset (x_off: 3 host doubles or NULL) */
int abrb_osc_generate_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q,
                          const double *dq, const double *target, int target_stride,
                          const double *target_velocity, int tv_stride, double *u,
                          double *training_signal, double *integrated_error, int64_t B, void *stream);
int abrb_osc_generate_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q,
                          const float *dq, const float *target, int target_stride,
                          const float *target_velocity, int tv_stride, float *u,
                          float *training_signal, float *integrated_error, int64_t B, void *stream);
/* HOST-pointer variants (H2D + kernel + D2H + sync inside) */
int abrb_osc_generate_host_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q,
                               const double *dq, const double *target, int target_stride,
                               const double *target_velocity, int tv_stride, double *u,
                               double *training_signal, double *integrated_error, int64_t B);
int abrb_osc_generate_host_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q,
                               const float *dq, const float *target, int target_stride,
                               const float *target_velocity, int tv_stride, float *u,
                               float *training_signal, float *integrated_error, int64_t B);
/* Asynchronous HOST-pointer variants for control loops that evaluate batch after batch: the call enqueues the H2D
 * copies, the kernel and the D2H copies of one batch on pipeline slot `slot` (0 or 1) of the calling thread and
 * returns; abrb_osc_host_wait(c, slot) blocks until that batch's outputs are in the caller's buffers.  With the two
 * slots used alternately, batch k+1's H2D runs under batch k's kernel and D2H (PCIe is full duplex), so a stream of
 * calls costs max(H2D, kernel, D2H) per batch instead of their sum.  The host buffers (page-locked for real
 * asynchrony) must stay valid and untouched until the wait; enqueueing on a slot first waits for its previous batch.
 * The synchronous variants above are slot 0 + wait. */
int abrb_osc_generate_host_async_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q,
                                     const double *dq, const double *target, int target_stride,
                                     const double *target_velocity, int tv_stride, double *u,
                                     double *training_signal, double *integrated_error, int64_t B, int slot);
int abrb_osc_generate_host_async_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q,
                                     const float *dq, const float *target, int target_stride,
                                     const float *target_velocity, int tv_stride, float *u,
                                     float *training_signal, float *integrated_error, int64_t B, int slot);
int abrb_osc_host_wait(const abrb_osc *c, int slot);

/* ---------------------------------------------------------------------------------------------------
 * Fused all-gather of the control outputs across the GPUs of one box (BASELINE config 5): one
 * process per GPU, every rank evaluates its own row block, and the OSC kernel's epilogue stores each finished tile of
 * `u` straight into the gathered (B_total, n) array of EVERY rank through NVLink peer memory (buffers mapped with CUDA
 * IPC) — the exchange overlaps the arithmetic instead of following it as a separate NCCL collective.
 *   abrb_gather_create   allocates this rank's region: n_buffers gathered arrays of bytes_per_buffer each (use two and
 *                        alternate them call by call: a fast rank may already write call k+1 while a slow one still
 *                        reads call k) plus the completion flags
 *   abrb_gather_export / abrb_gather_import   exchange the 64-byte CUDA IPC handles (any transport: MPI,
 *                        torch.distributed.all_gather_object, a file); every rank imports every other rank once
 *   abrb_osc_generate_gather_*   abrb_osc_generate_* whose rows additionally land at row `row0` of buffer
 *                        `buffer_index` on every rank (`u` may be NULL if only the gathered copy is wanted); all ranks
 *                        must make the same sequence of gather calls
 *   abrb_gather_wait     enqueues, on `stream`, a wait until every rank's rows of the latest gather call have arrived
 *                        in this rank's buffer (device-side spin on the flags; gives up after ~3 s and sets the status)
 *   abrb_gather_buffer   device pointer of this rank's gathered array `buffer_index`
 *   abrb_gather_status   0, or 1 if a wait ever timed out (call after synchronising the stream)
 * ------------------------------------------------------------------------------------------------- */
typedef struct abrb_gather abrb_gather;
int abrb_gather_create(int rank, int world, int64_t bytes_per_buffer, int n_buffers, abrb_gather **out);
int abrb_gather_destroy(abrb_gather *g);
int abrb_gather_export(const abrb_gather *g, unsigned char handle[64]);
int abrb_gather_import(abrb_gather *g, int peer_rank, const unsigned char handle[64]);
void *abrb_gather_buffer(const abrb_gather *g, int buffer_index);
int abrb_gather_wait(abrb_gather *g, void *stream);
int abrb_gather_status(const abrb_gather *g);
int abrb_osc_generate_gather_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q,
                                 const double *dq, const double *target, int target_stride,
                                 const double *target_velocity, int tv_stride, double *u, double *training_signal,
                                 double *integrated_error, int64_t B, abrb_gather *g, int buffer_index, int64_t row0,
                                 void *stream);
int abrb_osc_generate_gather_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q,
                                 const float *dq, const float *target, int target_stride,
                                 const float *target_velocity, int tv_stride, float *u, float *training_signal,
                                 float *integrated_error, int64_t B, abrb_gather *g, int buffer_index, int64_t row0,
                                 void *stream);

/* Standalone secondary controller: Damping / RestingConfig / AvoidObstacles / AvoidJointLimits
 * `.generate(q, dq)` -> (B,n). */
int abrb_null_generate_f64(const abrb_model *m, const abrb_null_params *p, const double *q,
                           const double *dq, double *u, int64_t B, void *stream);
int abrb_null_generate_f32(const abrb_model *m, const abrb_null_params *p, const float *q,
                           const float *dq, float *u, int64_t B, void *stream);

/* Joint-space PD controller  Joint.generate(q, dq, target, target_velocity=None)
 * (controllers/joint.py:104-131):  u = M (kp q_tilde + kv (target_velocity - dq)) [- g],
 * q_tilde = ((target - q + pi) mod 2 pi) - pi  (joint.py:42-46; the ball-joint/quaternion branch is MuJoCo-only).
 *   target (B,n) if target_stride == n, one (n,) row if 0;  target_velocity NULL or per tv_stride. */
int abrb_joint_generate_f64(const abrb_model *m, double kp, double kv, int account_for_gravity, const double *q,
                            const double *dq, const double *target, int target_stride,
                            const double *target_velocity, int tv_stride, double *u, int64_t B, void *stream);
int abrb_joint_generate_f32(const abrb_model *m, double kp, double kv, int account_for_gravity, const float *q,
                            const float *dq, const float *target, int target_stride,
                            const float *target_velocity, int tv_stride, float *u, int64_t B, void *stream);

/* Gravity compensation  Floating.generate(q, dq)  (controllers/floating.py:27-71):
 *   joint space:  u = -g;   task space:  u = J^T (-(M^-1 J^T Mx)^T g) with J = J("EE")[:3] and the reference's
 *   inv / pinv(rcond=1e-4) switch at |det| > 1e-3;   dynamic: u -= M dq.   dq may be NULL unless dynamic. */
int abrb_floating_generate_f64(const abrb_model *m, int task_space, int dynamic, const double *q, const double *dq,
                               double *u, int64_t B, void *stream);
int abrb_floating_generate_f32(const abrb_model *m, int task_space, int dynamic, const float *q, const float *dq,
                               float *u, int64_t B, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Closed-loop rollout: `steps` iterations of
 *     u = OSC.generate(q, dq, target);  ddq = M^-1 (u + g - C dq);  dq += ddq dt;  q += dq dt
 * (semi-implicit Euler as the reference's in-repo plant, arms/twojoint/arm_sim.py:131-132; the reference
 * has no plant for UR5/Jaco2, this one uses the same M, g, C the controller uses).
 *   q, dq   (B,n) in/out (final state);  target (B,6) or (6,) per target_stride
 *   q_traj / dq_traj / u_traj: NULL or (steps, B, n) outputs
 *   integrated_error: (B,6) in/out, given if and only if ki != 0 (as for abrb_osc_generate_*)
 * ------------------------------------------------------------------------------------------------- */
int abrb_osc_rollout_f64(const abrb_osc *c, int frame_id, const double *x_off, double *q, double *dq,
                         const double *target, int target_stride, int steps, double dt,
                         double *q_traj, double *dq_traj, double *u_traj, double *integrated_error, int64_t B,
                         void *stream);
int abrb_osc_rollout_f32(const abrb_osc *c, int frame_id, const double *x_off, float *q, float *dq,
                         const float *target, int target_stride, int steps, double dt,
                         float *q_traj, float *dq_traj, float *u_traj, float *integrated_error, int64_t B,
                         void *stream);

/* Kernel launch counter for this process (every launch of a libabrb kernel increments it). */
int64_t abrb_launch_count(void);

/* Sliding-mode controller  Sliding.generate(q, dq, target, target_velocity=0, target_acc=0, ref_frame="EE",
 * offset=None)  (controllers/sliding.py:34-99):
 *   cartesian != 0:  J = J(frame, x_off)[:3], dq_ref = pinv(J)(tv + lamb (target - Tx)),
 *                    ddq_ref = pinv(J)(ta + lamb (tv - J dq) - dJ[:3] dq_ref);   target/tv/ta rows have 3 values
 *   cartesian == 0:  dq_ref = tv - lamb (q - target), ddq_ref = ta - lamb (dq - tv);   rows have n values
 *   s = dq - dq_ref,  u = M ddq_ref + C dq_ref + g - kd s.
 * target: one row per state (stride = row width) or one broadcast row (stride 0); target_velocity / target_acc
 * likewise or NULL (zero).  s (B,n) receives the reference's `self.s` (the adaptation signal) or may be NULL. */
int abrb_sliding_generate_f64(const abrb_model *m, double kd, double lamb, int cartesian, int frame_id,
                              const double *x_off, const double *q, const double *dq, const double *target,
                              int target_stride, const double *target_velocity, int tv_stride,
                              const double *target_acc, int ta_stride, double *u, double *s, int64_t B,
                              void *stream);
int abrb_sliding_generate_f32(const abrb_model *m, double kd, double lamb, int cartesian, int frame_id,
                              const double *x_off, const float *q, const float *dq, const float *target,
                              int target_stride, const float *target_velocity, int tv_stride,
                              const float *target_acc, int ta_stride, float *u, float *s, int64_t B, void *stream);

/* Iterative inverse-kinematics path  InverseKinematics(robot_config, max_dx, max_dr, max_dq).generate_path(position,
 * target_position, n_timesteps, dt, method)  (controllers/path_planners/inverse_kinematics.py:28-166), one path per
 * state: n_timesteps resolved-motion steps towards target (x, y, z, Euler a, b, g) at frame "EE", the task-space
 * step clipped to max_dx dt / max_dr dt and the joint step to max_dq dt.  method 1: pinv(J); 2: damped least
 * squares; 3: position first, orientation in its null space (the reference's default).
 *   position (B,n);  target (B,6) if target_stride == 6, one (6,) row if 0
 *   position_path, velocity_path: (n_timesteps, B, n) out — row t holds q_t and the step dq_t taken from it */
int abrb_ik_path_f64(const abrb_model *m, double max_dx, double max_dr, double max_dq, int method, double dt,
                     int n_timesteps, const double *position, const double *target, int target_stride,
                     double *position_path, double *velocity_path, int64_t B, void *stream);
int abrb_ik_path_f32(const abrb_model *m, double max_dx, double max_dr, double max_dq, int method, double dt,
                     int n_timesteps, const float *position, const float *target, int target_stride,
                     float *position_path, float *velocity_path, int64_t B, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ABRB_H_ */
