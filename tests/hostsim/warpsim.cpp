// TEST INFRASTRUCTURE ONLY — never loaded by the abr_control_b200 package.
// Runs the warp-/CTA-cooperative part of the OSC kernels (abr_control_b200/csrc/abrb_coop.cuh: the six- or eight-lane
// groups, the round-robin Jacobi with its shuffles, the in-line pass over a ballot mask, the CTA's queue flush) on the
// CPU: every CUDA thread is an OS thread, `__shfl_sync` / `__any_sync` / `__syncwarp` / `__syncthreads` are barriers over
// the threads of the warp / CTA.  Slow, but it executes the very code the GPU executes, lane mapping included, so the
// `-m "not gpu"` suite covers it.      g++ -O1 -std=c++20 -pthread -shared -fPIC warpsim.cpp -o _warpsim.so
#include <barrier>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace warpsim {
constexpr int kMaxThreads = 256;
struct Cta {
  int n_threads;
  std::barrier<> cta_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  uint64_t slot[kMaxThreads];
  explicit Cta(int n) : n_threads(n), cta_bar(n) {
    for (int w = 0; w < n / 32; ++w) warp_bar.emplace_back(new std::barrier<>(32));
  }
};
thread_local Cta *t_cta = nullptr;
thread_local int t_tid = 0;
inline void warp_sync() { t_cta->warp_bar[t_tid >> 5]->arrive_and_wait(); }
}  // namespace warpsim

struct WsDim3 {
  int x;
};
thread_local WsDim3 threadIdx{0}, blockDim{32};

#define __device__
#define __forceinline__ inline
#define __noinline__

template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  warpsim::Cta &c = *warpsim::t_cta;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  c.slot[warpsim::t_tid] = bits;
  warpsim::warp_sync();
  const uint64_t r = c.slot[(warpsim::t_tid & ~31) + (src & 31)];
  warpsim::warp_sync();
  T out;
  std::memcpy(&out, &r, sizeof(T));
  return out;
}
inline unsigned __ballot_sync(unsigned, bool p) {
  warpsim::Cta &c = *warpsim::t_cta;
  c.slot[warpsim::t_tid] = p ? 1u : 0u;
  warpsim::warp_sync();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= unsigned(c.slot[(warpsim::t_tid & ~31) + l]) << l;
  warpsim::warp_sync();
  return m;
}
inline bool __any_sync(unsigned mask, bool p) { return __ballot_sync(mask, p) != 0u; }
inline void __syncwarp() { warpsim::warp_sync(); }
inline void __syncthreads() { warpsim::t_cta->cta_bar.arrive_and_wait(); }
inline unsigned __fns(unsigned mask, unsigned base, int offset) {  // the offset-th set bit of mask at or above bit `base`
  for (unsigned b = base; b < 32; ++b)
    if ((mask >> b) & 1u)
      if (--offset == 0) return b;
  return 0xffffffffu;
}
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

#include "../../abr_control_b200/csrc/abrb_coop.cuh"

using namespace abrb;

namespace {

template <class F>
void run_cta(int n_threads, F body) {
  warpsim::Cta cta(n_threads);
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t] {
      warpsim::t_cta = &cta;
      warpsim::t_tid = t;
      threadIdx.x = t;
      blockDim.x = n_threads;
      body(t);
    });
  for (auto &x : th) x.join();
}

template <int N, int KD>
struct Xch {  // exchange area of one warp as the in-line route sees it: y | z | A, slot-major
  static constexpr int kW = KD > N ? KD : N;
  struct Slot {
    static int at(int r, int k) { return 2 * kW + r * N + k; }
  };
  struct Layout {
    static constexpr int kY = 0, kZ = kW;
  };
  static constexpr int kSlots = 2 * kW + KD * N;
};

// One warp, lanes in `mask` wait with their own A (KD x N), y, z: -> wy, wz (N values per lane; untouched for other lanes)
template <int N, int KD>
void inline_route(unsigned mask, const double *A, const double *y, const double *z, double rcond, int two, double *wy,
                  double *wz) {
  typedef Xch<N, KD> X;
  std::vector<double> xch((size_t)X::kSlots * 32, 0.0);
  for (int l = 0; l < 32; ++l) {
    for (int r = 0; r < KD; ++r) {
      xch[(X::Layout::kY + r) * 32 + l] = y[l * KD + r];
      xch[(X::Layout::kZ + r) * 32 + l] = z[l * KD + r];
      for (int k = 0; k < N; ++k) xch[X::Slot::at(r, k) * 32 + l] = A[(l * KD + r) * N + k];
    }
  }
  run_cta(32, [&](int lane) {
    coop_pinv_warp<double, N, KD, typename X::Slot, typename X::Layout>(mask, xch.data(), xch.data(), lane, rcond, two != 0);
  });
  for (int l = 0; l < 32; ++l)
    for (int k = 0; k < N; ++k) {
      wy[l * N + k] = xch[(X::Layout::kY + k) * 32 + l];
      wz[l * N + k] = xch[(X::Layout::kZ + k) * 32 + l];
    }
}

// A CTA of n_threads empties a queue of n records (layout CoopRecord<N, KD>) into u / train (B x N)
template <int N, int KD>
void flush(int n_threads, int n, const double *qrec, const long long *qrow, double *u, double *train, double rcond,
           int two) {
  FlushOut<double> o{};
  o.u = u;
  o.train = train;
  o.n_peer = 0;
  o.self = 0;
  o.row0 = 0;
  run_cta(n_threads, [&](int) { coop_flush_cta<double, N, KD>(qrec, qrow, n, o, rcond, two != 0); });
}

// WarpCoop::pinv as osc_eval calls it: the lanes in `slow_mask` wait; the first `qcap` of them (in atomic order) leave
// a record in the CTA queue and go on with wy = wz = 0, the others are decomposed in line.  `both`: scratch in shared
// memory (fp64 kernels: A is read in place) or in registers (fp32 kernels: A is first copied into the exchange area).
template <int N, int KD, bool SHARED>
struct FakeKin {
  static constexpr bool kSharedScratch = SHARED;
  static int aslot(int r, int k) { return 3 + r * N + k; }  // some offset inside the scratch
  struct Store {
    const double *base;  // this lane's column of the warp's scratch (stride 32)
    double ld(int i) const { return base[i * 32]; }
  } s;
};

template <int N, int KD, bool SHARED>
void push_route(unsigned slow_mask, unsigned valid_mask, int qcap, const double *A, const double *Lfull, const double *y,
                const double *z, double rcond, int two, double *wy, double *wz, double *qrec, long long *qrow,
                int *qcount) {
  typedef FakeKin<N, KD, SHARED> K_;
  typedef WarpCoop<double, N, KD, K_> WC;
  std::vector<double> scratch((size_t)(3 + KD * N) * 32, 0.0), xch((size_t)WC::LY::kSlots * 32, 0.0);
  for (int l = 0; l < 32; ++l)
    for (int r = 0; r < KD; ++r)
      for (int k = 0; k < N; ++k) scratch[(size_t)K_::aslot(r, k) * 32 + l] = A[(l * KD + r) * N + k];
  *qcount = 0;
  run_cta(32, [&](int lane) {
    K_ K;
    K.s.base = scratch.data() + lane;
    WC coop{xch.data(), scratch.data(), lane, ((valid_mask >> lane) & 1u) != 0, qrec, qrow, qcount, 1000 + lane, qcap};
    double yy[KD], zz[KD], oy[N], oz[N];
    for (int r = 0; r < KD; ++r) {
      yy[r] = y[lane * KD + r];
      zz[r] = z[lane * KD + r];
    }
    for (int k = 0; k < N; ++k) oy[k] = oz[k] = -7.0;  // sentinel: lanes that do not wait keep their values
    auto Lget = [&](int a, int b) { return Lfull[(lane * N + a) * N + b]; };
    coop.template pinv<double, N, KD>(((slow_mask >> lane) & 1u) != 0, K, Lget, yy, zz, oy, oz, two != 0, rcond);
    for (int k = 0; k < N; ++k) {
      wy[lane * N + k] = oy[k];
      wz[lane * N + k] = oz[k];
    }
  });
}

}  // namespace

extern "C" void ws_push_6_6(int shared, unsigned slow_mask, unsigned valid_mask, int qcap, const double *A,
                            const double *Lfull, const double *y, const double *z, double rcond, int two, double *wy,
                            double *wz, double *qrec, long long *qrow, int *qcount) {
  if (shared)
    push_route<6, 6, true>(slow_mask, valid_mask, qcap, A, Lfull, y, z, rcond, two, wy, wz, qrec, qrow, qcount);
  else
    push_route<6, 6, false>(slow_mask, valid_mask, qcap, A, Lfull, y, z, rcond, two, wy, wz, qrec, qrow, qcount);
}

#define WS_EXPORT(N, KD)                                                                                               \
  extern "C" void ws_inline_##N##_##KD(unsigned mask, const double *A, const double *y, const double *z, double rcond, \
                                        int two, double *wy, double *wz) {                                            \
    inline_route<N, KD>(mask, A, y, z, rcond, two, wy, wz);                                                           \
  }                                                                                                                    \
  extern "C" void ws_flush_##N##_##KD(int n_threads, int n, const double *qrec, const long long *qrow, double *u,     \
                                       double *train, double rcond, int two) {                                        \
    flush<N, KD>(n_threads, n, qrec, qrow, u, train, rcond, two);                                                     \
  }                                                                                                                    \
  extern "C" int ws_layout_##N##_##KD(int what) {                                                                      \
    return what == 0 ? CoopGroup<N, KD>::kLanes : what == 1 ? CoopGroup<N, KD>::kPerWarp : CoopRecord<N, KD>::kLen;    \
  }
WS_EXPORT(6, 6)
WS_EXPORT(6, 3)
WS_EXPORT(6, 5)
WS_EXPORT(7, 6)
WS_EXPORT(7, 3)
WS_EXPORT(3, 3)
WS_EXPORT(2, 2)
