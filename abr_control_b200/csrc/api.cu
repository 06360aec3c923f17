// api.cu — the C ABI of libabrb.so (include/abrb.h): handle management, argument checking, dispatch on the
// joint count to the per-N kernel translation units, host-pointer convenience variants.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <utility>

#include "abrb_launch.hpp"

using namespace abrb;

struct abrb_model {
  abrb_chain_desc desc;
  ChainHost host;
};

// Index queue of the two-launch OSC mode (kernels.cu): one per (device, stream) the controller has been used on.
struct SlowQueue {
  int *buf = nullptr;       // 4 + cap ints
  void *records = nullptr;  // osc_record_len(n) * cap doubles (also used by the float kernels)
  int64_t cap = 0;
};

struct abrb_osc {
  const abrb_model *model;
  abrb_osc_params params;
  int64_t two_launch_min = 0;  // 0: single launch always
  mutable std::mutex mu;
  mutable std::map<std::pair<int, cudaStream_t>, SlowQueue> queues;
};

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

int cuda_fail(int e, const char *where) {
  return fail(ABRB_ECUDA, std::string(where) + ": " + cudaGetErrorString((cudaError_t)e));
}

#ifndef ABRB_N_LIST
#define ABRB_N_LIST(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#endif

bool n_supported(int n) {
#define X(k) if (n == k) return true;
  ABRB_N_LIST(X)
#undef X
  return false;
}

int ensure_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(ABRB_ECUDA, "no CUDA device available (libabrb has no CPU fallback)");
  }
  return ABRB_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// grow-only device workspace for the *_host entry points (one per host thread)
struct Workspace {
  void *ptr = nullptr;
  size_t cap = 0;
  int dev = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t lanes[3] = {nullptr, nullptr, nullptr};  // chunk pipeline of the *_host OSC call
  int ensure(size_t bytes) {
    int cur = 0;
    cudaError_t ed = cudaGetDevice(&cur);
    if (ed != cudaSuccess) return (int)ed;
    if (dev != cur) {  // the calling thread switched devices: streams and memory belong to the device they were made on
      if (dev >= 0) {
        cudaSetDevice(dev);
        cudaFree(ptr);
        if (stream) cudaStreamDestroy(stream);
        for (auto &l : lanes)
          if (l) cudaStreamDestroy(l);
        cudaSetDevice(cur);
      }
      ptr = nullptr;
      cap = 0;
      stream = nullptr;
      for (auto &l : lanes) l = nullptr;
      dev = cur;
    }
    if (stream == nullptr) {
      cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
      if (e != cudaSuccess) return (int)e;
      for (auto &l : lanes) {
        e = cudaStreamCreateWithFlags(&l, cudaStreamNonBlocking);
        if (e != cudaSuccess) return (int)e;
      }
    }
    if (bytes <= cap) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    cap = bytes;
    return 0;
  }
};
thread_local Workspace g_ws;

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

// a cudaMemcpyAsync that is rejected (bad pointer, wrong direction) fails at the call, NOT at the later synchronise
#define ABRB_CU(call, where)                                \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) return cuda_fail((int)e_, where); \
  } while (0)

}  // namespace

namespace abrb {
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace abrb

extern "C" {

int abrb_version(void) { return ABRB_VERSION; }
const char *abrb_last_error(void) { return g_err.c_str(); }
int64_t abrb_launch_count(void) { return g_launches.load(); }

int abrb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return fail(ABRB_ECUDA, "cudaGetDeviceCount failed");
  }
  return n;
}

int abrb_model_create(const abrb_chain_desc *desc, abrb_model **out) {
  if (!desc || !out) return fail(ABRB_EINVAL, "abrb_model_create: NULL argument");
  *out = nullptr;
  if (desc->n_joints < 1 || desc->n_joints > ABRB_MAX_JOINTS || !n_supported(desc->n_joints))
    return fail(ABRB_ESHAPE, "abrb_model_create: n_joints not supported by this build");
  abrb_model *m = new (std::nothrow) abrb_model;
  if (!m) return fail(ABRB_ENOMEM, "abrb_model_create: out of memory");
  m->desc = *desc;
  std::string e = chain_from_desc(*desc, m->host);
  if (!e.empty()) {
    delete m;
    return fail(ABRB_ESHAPE, "abrb_model_create: " + e);
  }
  *out = m;
  return ABRB_OK;
}

int abrb_model_destroy(abrb_model *m) {
  delete m;
  return ABRB_OK;
}

int abrb_model_n_joints(const abrb_model *m) { return m ? m->host.n : fail(ABRB_EINVAL, "NULL model"); }
int abrb_model_is_orthonormal(const abrb_model *m) { return m ? (m->host.ortho ? 1 : 0) : fail(ABRB_EINVAL, "NULL model"); }

int abrb_frame_id(const abrb_model *m, const char *name) {
  if (!m) return fail(ABRB_EINVAL, "NULL model");
  int id = parse_frame(m->host.n, name);
  if (id < 0) return fail(ABRB_EFRAME, std::string("Invalid transformation name: ") + (name ? name : "(null)"));
  return id;
}

// ------------------------------------------------------------------------------------------------ rbd
static int rbd_eval(const abrb_model *m, int frame_id, const double *x_off, const void *q, const void *dq,
                    int64_t B, const abrb_rbd_out *out, void *stream, bool f32) {
  if (!m || !out) return fail(ABRB_EINVAL, "abrb_rbd_eval: NULL model/out");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_rbd_eval: B < 0");
  const int n = m->host.n;
  if (frame_id < 0 || frame_id > 2 * n + 1) return fail(ABRB_EFRAME, "abrb_rbd_eval: invalid frame id");
  if ((out->dJ || out->C) && !dq) return fail(ABRB_EINVAL, "abrb_rbd_eval: dJ / C need dq");
  if (B == 0) return ABRB_OK;
  if (!q) return fail(ABRB_EINVAL, "abrb_rbd_eval: NULL q");
  const void *ptrs[] = {q, dq, out->Tx, out->T, out->R, out->T_inv, out->quat, out->J, out->dJ, out->M, out->g, out->C};
  for (const void *p : ptrs)
    if (p && !aligned16(p)) return fail(ABRB_EINVAL, "abrb_rbd_eval: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  RbdCall c{frame_id, x_off, q, dq, B, *out, f32, (cudaStream_t)stream};
  int e = cudaErrorInvalidValue;
  switch (n) {
#define X(k) case k: e = launch_rbd<k>(m->host, c); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, "abrb_rbd_eval") : ABRB_OK;
}

int abrb_rbd_eval_f64(const abrb_model *m, int frame_id, const double *x_off, const double *q, const double *dq,
                      int64_t B, const abrb_rbd_out *out, void *stream) {
  return rbd_eval(m, frame_id, x_off, q, dq, B, out, stream, false);
}
int abrb_rbd_eval_f32(const abrb_model *m, int frame_id, const double *x_off, const float *q, const float *dq,
                      int64_t B, const abrb_rbd_out *out, void *stream) {
  return rbd_eval(m, frame_id, x_off, q, dq, B, out, stream, true);
}

static int rbd_eval_host(const abrb_model *m, int frame_id, const double *x_off, const void *q, const void *dq,
                         int64_t B, const abrb_rbd_out *out, bool f32) {
  if (!m || !out) return fail(ABRB_EINVAL, "abrb_rbd_eval_host: NULL model/out");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_rbd_eval_host: B < 0");
  if (B == 0) return ABRB_OK;
  if (!q) return fail(ABRB_EINVAL, "abrb_rbd_eval_host: NULL q");
  int rc = ensure_device();
  if (rc) return rc;
  const size_t es = f32 ? 4 : 8, n = (size_t)m->host.n;
  const size_t len[10] = {3, 16, 9, 16, 4, 6 * n, 6 * n, n * n, n, n * n};
  void *const host_out[10] = {out->Tx, out->T, out->R, out->T_inv, out->quat, out->J, out->dJ, out->M, out->g, out->C};
  size_t total = 2 * align_up((size_t)B * n * es);
  for (int i = 0; i < 10; ++i)
    if (host_out[i]) total += align_up((size_t)B * len[i] * es);
  int e = g_ws.ensure(total);
  if (e) return cuda_fail(e, "abrb_rbd_eval_host(workspace)");
  char *base = static_cast<char *>(g_ws.ptr);
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base + off; off += align_up(bytes); return (void *)p; };
  void *dq_q = take((size_t)B * n * es), *dq_dq = take((size_t)B * n * es);
  cudaStream_t s = g_ws.stream;
  ABRB_CU(cudaMemcpyAsync(dq_q, q, (size_t)B * n * es, cudaMemcpyHostToDevice, s), "abrb_rbd_eval_host(q)");
  if (dq) ABRB_CU(cudaMemcpyAsync(dq_dq, dq, (size_t)B * n * es, cudaMemcpyHostToDevice, s), "abrb_rbd_eval_host(dq)");
  void *dev_out[10];
  for (int i = 0; i < 10; ++i) dev_out[i] = host_out[i] ? take((size_t)B * len[i] * es) : nullptr;
  abrb_rbd_out d{dev_out[0], dev_out[1], dev_out[2], dev_out[3], dev_out[4], dev_out[5], dev_out[6], dev_out[7], dev_out[8], dev_out[9]};
  rc = rbd_eval(m, frame_id, x_off, dq_q, dq ? dq_dq : nullptr, B, &d, s, f32);
  if (rc) return rc;
  for (int i = 0; i < 10; ++i)
    if (host_out[i])
      ABRB_CU(cudaMemcpyAsync(host_out[i], dev_out[i], (size_t)B * len[i] * es, cudaMemcpyDeviceToHost, s),
              "abrb_rbd_eval_host(result)");
  cudaError_t ce = cudaStreamSynchronize(s);
  return ce ? cuda_fail(ce, "abrb_rbd_eval_host") : ABRB_OK;
}

int abrb_rbd_eval_host_f64(const abrb_model *m, int frame_id, const double *x_off, const double *q, const double *dq,
                           int64_t B, const abrb_rbd_out *out) {
  return rbd_eval_host(m, frame_id, x_off, q, dq, B, out, false);
}
int abrb_rbd_eval_host_f32(const abrb_model *m, int frame_id, const double *x_off, const float *q, const float *dq,
                           int64_t B, const abrb_rbd_out *out) {
  return rbd_eval_host(m, frame_id, x_off, q, dq, B, out, true);
}

// ------------------------------------------------------------------------------------------------ osc
int abrb_osc_create(const abrb_model *m, const abrb_osc_params *p, abrb_osc **out) {
  if (!m || !p || !out) return fail(ABRB_EINVAL, "abrb_osc_create: NULL argument");
  *out = nullptr;
  std::string e = check_osc(m->host.n, *p);
  if (!e.empty()) return fail(ABRB_EUNSUP, "abrb_osc_create: " + e);
  abrb_osc *c = new (std::nothrow) abrb_osc;
  if (!c) return fail(ABRB_ENOMEM, "abrb_osc_create: out of memory");
  c->model = m;
  c->params = *p;
  if (const char *v = std::getenv("ABRB_OSC_DEFER_MIN")) c->two_launch_min = (int64_t)std::atoll(v);
  *out = c;
  return ABRB_OK;
}

int abrb_osc_set_option(abrb_osc *c, const char *name, double value) {
  if (!c || !name) return fail(ABRB_EINVAL, "abrb_osc_set_option: NULL argument");
  if (std::strcmp(name, "two_launch_min") == 0) {
    c->two_launch_min = value > 0 ? (int64_t)value : 0;
    return ABRB_OK;
  }
  return fail(ABRB_EINVAL, std::string("abrb_osc_set_option: unknown option ") + name);
}

int abrb_osc_destroy(abrb_osc *c) {
  if (c) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto &kv : c->queues) {
      cudaSetDevice(kv.first.first);
      cudaFree(kv.second.buf);
      cudaFree(kv.second.records);
    }
    cudaSetDevice(cur);
  }
  delete c;
  return ABRB_OK;
}

// Two-launch mode (kernels.cu, osc_kernel<DEFER> + osc_slow_kernel) for the 6-row task space, from
// `two_launch_min` states up.  Returns the queue for this (device, stream), growing it on demand, or nullptr for the
// single-launch mode.
static SlowQueue slow_queue_for(const abrb_osc *c, int64_t B, cudaStream_t stream, int *err) {
  const int64_t min_b = c->two_launch_min > 0 ? c->two_launch_min : -1;
  *err = 0;
  const abrb_osc_params &p = c->params;
  if (min_b < 0 || B < min_b || B > (int64_t)0x7fffffff - 8) return SlowQueue();
  if (!(p.ctrlr_dof[3] || p.ctrlr_dof[4] || p.ctrlr_dof[5])) return SlowQueue();
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(c->mu);
  SlowQueue &sq = c->queues[std::make_pair(dev, stream)];
  if (sq.cap < B) {
    cudaError_t e = cudaSuccess;
    if (sq.buf) e = cudaFree(sq.buf);  // synchronises with any launch still using it
    if (sq.records) cudaFree(sq.records);
    sq = SlowQueue();
    if (e == cudaSuccess) e = cudaMalloc(&sq.buf, (size_t)(B + 4) * sizeof(int));
    if (e == cudaSuccess)
      e = cudaMalloc(&sq.records, (size_t)osc_record_len(c->model->host.n) * (size_t)B * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(sq.buf, 0, 4 * sizeof(int), stream);
    if (e != cudaSuccess) {
      cudaFree(sq.buf);
      cudaFree(sq.records);
      sq = SlowQueue();
      *err = (int)e;
      return SlowQueue();
    }
    sq.cap = B;
  }
  return sq;
}

static int osc_generate(const abrb_osc *c, int frame_id, const double *x_off, const void *q, const void *dq,
                        const void *target, int target_stride, const void *tv, int tv_stride, void *u, void *train,
                        int64_t B, void *stream, bool f32) {
  if (!c) return fail(ABRB_EINVAL, "abrb_osc_generate: NULL controller");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_osc_generate: B < 0");
  const int n = c->model->host.n;
  if (frame_id < 0 || frame_id > 2 * n + 1) return fail(ABRB_EFRAME, "abrb_osc_generate: invalid frame id");
  if ((target_stride != 0 && target_stride != 6) || (tv && tv_stride != 0 && tv_stride != 6))
    return fail(ABRB_EINVAL, "abrb_osc_generate: stride must be 0 (broadcast) or 6");
  if (B == 0) return ABRB_OK;
  if (!q || !dq || !target || !u) return fail(ABRB_EINVAL, "abrb_osc_generate: NULL q/dq/target/u");
  const void *ptrs[] = {q, dq, u, train};
  for (const void *p : ptrs)
    if (p && !aligned16(p)) return fail(ABRB_EINVAL, "abrb_osc_generate: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  OscCall k{frame_id, x_off, q, dq, target, tv, target_stride, tv_stride, u, train, B, f32, (cudaStream_t)stream};
  int qe = 0;
  const SlowQueue sq = slow_queue_for(c, B, (cudaStream_t)stream, &qe);
  if (qe) return cuda_fail(qe, "abrb_osc_generate (two-launch workspace)");
  k.queue = sq.buf;
  k.records = sq.records;
  k.rec_stride = sq.cap;
  int e = cudaErrorInvalidValue;
  switch (n) {
#define X(j) case j: e = launch_osc<j>(c->model->host, c->params, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, "abrb_osc_generate") : ABRB_OK;
}

int abrb_osc_generate_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q, const double *dq,
                          const double *target, int target_stride, const double *target_velocity, int tv_stride,
                          double *u, double *training_signal, int64_t B, void *stream) {
  return osc_generate(c, frame_id, x_off, q, dq, target, target_stride, target_velocity, tv_stride, u,
                      training_signal, B, stream, false);
}
int abrb_osc_generate_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q, const float *dq,
                          const float *target, int target_stride, const float *target_velocity, int tv_stride,
                          float *u, float *training_signal, int64_t B, void *stream) {
  return osc_generate(c, frame_id, x_off, q, dq, target, target_stride, target_velocity, tv_stride, u,
                      training_signal, B, stream, true);
}

static int osc_generate_host(const abrb_osc *c, int frame_id, const double *x_off, const void *q, const void *dq,
                             const void *target, int target_stride, const void *tv, int tv_stride, void *u,
                             void *train, int64_t B, bool f32) {
  if (!c) return fail(ABRB_EINVAL, "abrb_osc_generate_host: NULL controller");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_osc_generate_host: B < 0");
  if (B == 0) return ABRB_OK;
  if (!q || !dq || !target || !u) return fail(ABRB_EINVAL, "abrb_osc_generate_host: NULL q/dq/target/u");
  int rc = ensure_device();
  if (rc) return rc;
  const size_t es = f32 ? 4 : 8, n = (size_t)c->model->host.n;
  const size_t sz_state = (size_t)B * n * es;
  const size_t sz_t = (target_stride ? (size_t)B : 1) * 6 * es, sz_tv = tv ? (tv_stride ? (size_t)B : 1) * 6 * es : 0;
  int e = g_ws.ensure(4 * align_up(sz_state) + align_up(sz_t) + align_up(sz_tv) + 256);
  if (e) return cuda_fail(e, "abrb_osc_generate_host(workspace)");
  char *base = static_cast<char *>(g_ws.ptr);
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base + off; off += align_up(bytes); return (void *)p; };
  void *d_q = take(sz_state), *d_dq = take(sz_state), *d_u = take(sz_state), *d_tr = take(sz_state);
  void *d_t = take(sz_t), *d_tv = tv ? take(sz_tv) : nullptr;
  // Chunked 3-stage pipeline over three streams: the H2D copy of chunk c+1 overlaps the kernel of chunk c and the
  // D2H copy of chunk c-1 (the copy engines are full duplex), so a large batch costs ~max(H2D, kernel, D2H).
  const size_t row = n * es;
  static const int64_t chunk_env = [] {  // tuning knob: states per pipeline chunk
    const char *v = std::getenv("ABRB_HOST_CHUNK");
    return v ? (int64_t)std::atoll(v) : (int64_t)0;
  }();
  // measured on B200 (tools/dbg/e2e_probe.py, UR5 6-DOF fp64, B = 65536): 1 chunk 347 us, 2 chunks 306 us, 3: 310,
  // 4: 330, 8: 373 — every chunk costs ~20 us of copy/launch overheads, so two chunks up to ~200 k states, four above
  int64_t chunk = B < 49152 ? B : ((B + (B <= 196608 ? 1 : 3)) / (B <= 196608 ? 2 : 4) + 127) / 128 * 128;
  if (chunk_env > 0) chunk = (chunk_env < B ? chunk_env : B + 127) / 128 * 128;
  if (chunk <= 0) chunk = B;
  cudaError_t ce = cudaSuccess;
  if (!target_stride)
    ABRB_CU(cudaMemcpyAsync(d_t, target, sz_t, cudaMemcpyHostToDevice, g_ws.stream), "abrb_osc_generate_host(target)");
  if (tv && !tv_stride)
    ABRB_CU(cudaMemcpyAsync(d_tv, tv, sz_tv, cudaMemcpyHostToDevice, g_ws.stream), "abrb_osc_generate_host(target_velocity)");
  if (!target_stride || (tv && !tv_stride)) {  // broadcast rows must be resident before any lane starts
    ce = cudaStreamSynchronize(g_ws.stream);
    if (ce) return cuda_fail(ce, "abrb_osc_generate_host");
  }
  int lane = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, lane = (lane + 1) % 3) {
    const int64_t nb = B - b0 < chunk ? B - b0 : chunk;
    cudaStream_t s = g_ws.lanes[lane];
    const size_t off_s = (size_t)b0 * row, off_t = (size_t)b0 * 6 * es;
    auto at = [](const void *p, size_t o) { return (const void *)((const char *)p + o); };
    auto atw = [](void *p, size_t o) { return (void *)((char *)p + o); };
    const char *where = "abrb_osc_generate_host(copy in)";
    ABRB_CU(cudaMemcpyAsync(atw(d_q, off_s), at(q, off_s), (size_t)nb * row, cudaMemcpyHostToDevice, s), where);
    ABRB_CU(cudaMemcpyAsync(atw(d_dq, off_s), at(dq, off_s), (size_t)nb * row, cudaMemcpyHostToDevice, s), where);
    if (target_stride)
      ABRB_CU(cudaMemcpyAsync(atw(d_t, off_t), at(target, off_t), (size_t)nb * 6 * es, cudaMemcpyHostToDevice, s), where);
    if (tv && tv_stride)
      ABRB_CU(cudaMemcpyAsync(atw(d_tv, off_t), at(tv, off_t), (size_t)nb * 6 * es, cudaMemcpyHostToDevice, s), where);
    rc = osc_generate(c, frame_id, x_off, at(d_q, off_s), at(d_dq, off_s), target_stride ? at(d_t, off_t) : d_t,
                      target_stride, tv ? (tv_stride ? at(d_tv, off_t) : d_tv) : nullptr, tv_stride, atw(d_u, off_s),
                      train ? atw(d_tr, off_s) : nullptr, nb, s, f32);
    if (rc) return rc;
    ABRB_CU(cudaMemcpyAsync(atw(u, off_s), at(d_u, off_s), (size_t)nb * row, cudaMemcpyDeviceToHost, s),
            "abrb_osc_generate_host(result)");
    if (train)
      ABRB_CU(cudaMemcpyAsync(atw(train, off_s), at(d_tr, off_s), (size_t)nb * row, cudaMemcpyDeviceToHost, s),
              "abrb_osc_generate_host(training signal)");
  }
  const int used = (int)((B + chunk - 1) / chunk) < 3 ? (int)((B + chunk - 1) / chunk) : 3;
  for (int l = 0; l < used; ++l) {
    ce = cudaStreamSynchronize(g_ws.lanes[l]);
    if (ce) return cuda_fail(ce, "abrb_osc_generate_host");
  }
  return ABRB_OK;
}

int abrb_osc_generate_host_f64(const abrb_osc *c, int frame_id, const double *x_off, const double *q, const double *dq,
                               const double *target, int target_stride, const double *target_velocity,
                               int tv_stride, double *u, double *training_signal, int64_t B) {
  return osc_generate_host(c, frame_id, x_off, q, dq, target, target_stride, target_velocity, tv_stride, u,
                           training_signal, B, false);
}
int abrb_osc_generate_host_f32(const abrb_osc *c, int frame_id, const double *x_off, const float *q, const float *dq,
                               const float *target, int target_stride, const float *target_velocity, int tv_stride,
                               float *u, float *training_signal, int64_t B) {
  return osc_generate_host(c, frame_id, x_off, q, dq, target, target_stride, target_velocity, tv_stride, u,
                           training_signal, B, true);
}

// ------------------------------------------------------------------------------------------------ null
static int null_generate(const abrb_model *m, const abrb_null_params *p, const void *q, const void *dq, void *u,
                         int64_t B, void *stream, bool f32) {
  if (!m || !p) return fail(ABRB_EINVAL, "abrb_null_generate: NULL argument");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_null_generate: B < 0");
  std::string e = check_null(m->host.n, *p);
  if (!e.empty()) return fail(ABRB_EUNSUP, "abrb_null_generate: " + e);
  if (B == 0) return ABRB_OK;
  if (!q || !dq || !u) return fail(ABRB_EINVAL, "abrb_null_generate: NULL q/dq/u");
  if (!aligned16(q) || !aligned16(dq) || !aligned16(u))
    return fail(ABRB_EINVAL, "abrb_null_generate: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  NullCall k{q, dq, u, B, f32, (cudaStream_t)stream};
  int ce = cudaErrorInvalidValue;
  switch (m->host.n) {
#define X(j) case j: ce = launch_null<j>(m->host, *p, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return ce ? cuda_fail(ce, "abrb_null_generate") : ABRB_OK;
}

int abrb_null_generate_f64(const abrb_model *m, const abrb_null_params *p, const double *q, const double *dq,
                           double *u, int64_t B, void *stream) {
  return null_generate(m, p, q, dq, u, B, stream, false);
}
int abrb_null_generate_f32(const abrb_model *m, const abrb_null_params *p, const float *q, const float *dq, float *u,
                           int64_t B, void *stream) {
  return null_generate(m, p, q, dq, u, B, stream, true);
}

// ------------------------------------------------------------------------------------------------ sliding
static int sliding_generate(const abrb_model *m, double kd, double lamb, int cartesian, int frame_id,
                            const double *x_off, const void *q, const void *dq, const void *target, int target_stride,
                            const void *tv, int tv_stride, const void *ta, int ta_stride, void *u, void *s, int64_t B,
                            void *stream, bool f32) {
  if (!m) return fail(ABRB_EINVAL, "abrb_sliding_generate: NULL model");
  if (B < 0) return fail(ABRB_EINVAL, "abrb_sliding_generate: B < 0");
  const int n = m->host.n, w = cartesian ? 3 : n;
  if (frame_id < 0 || frame_id > 2 * n + 1) return fail(ABRB_EFRAME, "abrb_sliding_generate: invalid frame id");
  if ((target_stride != 0 && target_stride != w) || (tv && tv_stride != 0 && tv_stride != w) ||
      (ta && ta_stride != 0 && ta_stride != w))
    return fail(ABRB_EINVAL, "abrb_sliding_generate: stride must be 0 (broadcast) or the row width (3 or n_joints)");
  if (B == 0) return ABRB_OK;
  if (!q || !dq || !target || !u) return fail(ABRB_EINVAL, "abrb_sliding_generate: NULL q/dq/target/u");
  if (!aligned16(q) || !aligned16(dq) || !aligned16(u) || (s && !aligned16(s)))
    return fail(ABRB_EINVAL, "abrb_sliding_generate: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  SlidingCall k{kd, lamb, cartesian, frame_id, x_off, q, dq, target, tv, ta, target_stride, tv_stride, ta_stride,
                u, s, B, f32, (cudaStream_t)stream};
  int e = cudaErrorInvalidValue;
  switch (n) {
#define X(j) case j: e = launch_sliding<j>(m->host, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, "abrb_sliding_generate") : ABRB_OK;
}

int abrb_sliding_generate_f64(const abrb_model *m, double kd, double lamb, int cartesian, int frame_id,
                              const double *x_off, const double *q, const double *dq, const double *target,
                              int target_stride, const double *target_velocity, int tv_stride,
                              const double *target_acc, int ta_stride, double *u, double *s, int64_t B, void *stream) {
  return sliding_generate(m, kd, lamb, cartesian, frame_id, x_off, q, dq, target, target_stride, target_velocity,
                          tv_stride, target_acc, ta_stride, u, s, B, stream, false);
}
int abrb_sliding_generate_f32(const abrb_model *m, double kd, double lamb, int cartesian, int frame_id,
                              const double *x_off, const float *q, const float *dq, const float *target,
                              int target_stride, const float *target_velocity, int tv_stride, const float *target_acc,
                              int ta_stride, float *u, float *s, int64_t B, void *stream) {
  return sliding_generate(m, kd, lamb, cartesian, frame_id, x_off, q, dq, target, target_stride, target_velocity,
                          tv_stride, target_acc, ta_stride, u, s, B, stream, true);
}

// ------------------------------------------------------------------------------------------------ inverse kinematics
static int ik_path(const abrb_model *m, double max_dx, double max_dr, double max_dq, int method, double dt, int steps,
                   const void *position, const void *target, int target_stride, void *pos_path, void *vel_path,
                   int64_t B, void *stream, bool f32) {
  if (!m) return fail(ABRB_EINVAL, "abrb_ik_path: NULL model");
  if (B < 0 || steps < 0) return fail(ABRB_EINVAL, "abrb_ik_path: negative size");
  if (method < 1 || method > 3) return fail(ABRB_EUNSUP, "abrb_ik_path: method must be 1, 2 or 3");
  if (target_stride != 0 && target_stride != 6) return fail(ABRB_EINVAL, "abrb_ik_path: stride must be 0 or 6");
  if (B == 0 || steps == 0) return ABRB_OK;
  if (!position || !target || !pos_path || !vel_path) return fail(ABRB_EINVAL, "abrb_ik_path: NULL argument");
  if (!aligned16(position) || !aligned16(pos_path) || !aligned16(vel_path))
    return fail(ABRB_EINVAL, "abrb_ik_path: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  IkCall k{max_dx, max_dr, max_dq, dt, method, steps, position, target, target_stride, pos_path, vel_path, B, f32,
           (cudaStream_t)stream};
  int e = cudaErrorInvalidValue;
  switch (m->host.n) {
#define X(j) case j: e = launch_ik<j>(m->host, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, "abrb_ik_path") : ABRB_OK;
}

int abrb_ik_path_f64(const abrb_model *m, double max_dx, double max_dr, double max_dq, int method, double dt,
                     int n_timesteps, const double *position, const double *target, int target_stride,
                     double *position_path, double *velocity_path, int64_t B, void *stream) {
  return ik_path(m, max_dx, max_dr, max_dq, method, dt, n_timesteps, position, target, target_stride, position_path,
                 velocity_path, B, stream, false);
}
int abrb_ik_path_f32(const abrb_model *m, double max_dx, double max_dr, double max_dq, int method, double dt,
                     int n_timesteps, const float *position, const float *target, int target_stride,
                     float *position_path, float *velocity_path, int64_t B, void *stream) {
  return ik_path(m, max_dx, max_dr, max_dq, method, dt, n_timesteps, position, target, target_stride, position_path,
                 velocity_path, B, stream, true);
}

// ------------------------------------------------------------------------------------------------ joint / floating
static int ctrl_generate(const abrb_model *m, int kind, double kp, double kv, int fa, int fb, const void *q,
                         const void *dq, const void *target, int target_stride, const void *tv, int tv_stride, void *u,
                         int64_t B, void *stream, bool f32, const char *who) {
  if (!m) return fail(ABRB_EINVAL, std::string(who) + ": NULL model");
  if (B < 0) return fail(ABRB_EINVAL, std::string(who) + ": B < 0");
  const int n = m->host.n;
  if (kind == 0 && ((target_stride != 0 && target_stride != n) || (tv && tv_stride != 0 && tv_stride != n)))
    return fail(ABRB_EINVAL, std::string(who) + ": stride must be 0 (broadcast) or n_joints");
  if (B == 0) return ABRB_OK;
  if (!q || !u || (kind == 0 && (!dq || !target)) || (kind == 1 && fb && !dq))
    return fail(ABRB_EINVAL, std::string(who) + ": NULL q/dq/target/u");
  if (!aligned16(q) || !aligned16(u) || (dq && !aligned16(dq)))
    return fail(ABRB_EINVAL, std::string(who) + ": pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  CtrlCall k{kind, kp, kv, fa, fb, q, dq, target, tv, target_stride, tv_stride, u, B, f32, (cudaStream_t)stream};
  int e = cudaErrorInvalidValue;
  switch (n) {
#define X(j) case j: e = launch_ctrl<j>(m->host, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, who) : ABRB_OK;
}

int abrb_joint_generate_f64(const abrb_model *m, double kp, double kv, int account_for_gravity, const double *q,
                            const double *dq, const double *target, int target_stride, const double *target_velocity,
                            int tv_stride, double *u, int64_t B, void *stream) {
  return ctrl_generate(m, 0, kp, kv, account_for_gravity, 0, q, dq, target, target_stride, target_velocity, tv_stride, u,
                       B, stream, false, "abrb_joint_generate");
}
int abrb_joint_generate_f32(const abrb_model *m, double kp, double kv, int account_for_gravity, const float *q,
                            const float *dq, const float *target, int target_stride, const float *target_velocity,
                            int tv_stride, float *u, int64_t B, void *stream) {
  return ctrl_generate(m, 0, kp, kv, account_for_gravity, 0, q, dq, target, target_stride, target_velocity, tv_stride, u,
                       B, stream, true, "abrb_joint_generate");
}
int abrb_floating_generate_f64(const abrb_model *m, int task_space, int dynamic, const double *q, const double *dq,
                               double *u, int64_t B, void *stream) {
  return ctrl_generate(m, 1, 0, 0, task_space, dynamic, q, dq, nullptr, 0, nullptr, 0, u, B, stream, false,
                       "abrb_floating_generate");
}
int abrb_floating_generate_f32(const abrb_model *m, int task_space, int dynamic, const float *q, const float *dq, float *u,
                               int64_t B, void *stream) {
  return ctrl_generate(m, 1, 0, 0, task_space, dynamic, q, dq, nullptr, 0, nullptr, 0, u, B, stream, true,
                       "abrb_floating_generate");
}

// ------------------------------------------------------------------------------------------------ rollout
static int osc_rollout(const abrb_osc *c, int frame_id, const double *x_off, void *q, void *dq, const void *target,
                       int target_stride, int steps, double dt, void *q_traj, void *dq_traj, void *u_traj, int64_t B,
                       void *stream, bool f32) {
  if (!c) return fail(ABRB_EINVAL, "abrb_osc_rollout: NULL controller");
  if (B < 0 || steps < 0) return fail(ABRB_EINVAL, "abrb_osc_rollout: B < 0 or steps < 0");
  const int n = c->model->host.n;
  if (frame_id < 0 || frame_id > 2 * n + 1) return fail(ABRB_EFRAME, "abrb_osc_rollout: invalid frame id");
  if (target_stride != 0 && target_stride != 6) return fail(ABRB_EINVAL, "abrb_osc_rollout: stride must be 0 or 6");
  if (B == 0 || steps == 0) return ABRB_OK;
  if (!q || !dq || !target) return fail(ABRB_EINVAL, "abrb_osc_rollout: NULL q/dq/target");
  const void *ptrs[] = {q, dq, q_traj, dq_traj, u_traj};
  for (const void *p : ptrs)
    if (p && !aligned16(p)) return fail(ABRB_EINVAL, "abrb_osc_rollout: pointers must be 16-byte aligned");
  int rc = ensure_device();
  if (rc) return rc;
  RolloutCall k{frame_id, x_off, q, dq, target, target_stride, steps, dt, q_traj, dq_traj, u_traj, B, f32, (cudaStream_t)stream};
  int e = cudaErrorInvalidValue;
  switch (n) {
#define X(j) case j: e = launch_rollout<j>(c->model->host, c->params, k); break;
    ABRB_N_LIST(X)
#undef X
  }
  return e ? cuda_fail(e, "abrb_osc_rollout") : ABRB_OK;
}

int abrb_osc_rollout_f64(const abrb_osc *c, int frame_id, const double *x_off, double *q, double *dq,
                         const double *target, int target_stride, int steps, double dt, double *q_traj,
                         double *dq_traj, double *u_traj, int64_t B, void *stream) {
  return osc_rollout(c, frame_id, x_off, q, dq, target, target_stride, steps, dt, q_traj, dq_traj, u_traj, B, stream, false);
}
int abrb_osc_rollout_f32(const abrb_osc *c, int frame_id, const double *x_off, float *q, float *dq, const float *target,
                         int target_stride, int steps, double dt, float *q_traj, float *dq_traj, float *u_traj,
                         int64_t B, void *stream) {
  return osc_rollout(c, frame_id, x_off, q, dq, target, target_stride, steps, dt, q_traj, dq_traj, u_traj, B, stream, true);
}

}  // extern "C"
