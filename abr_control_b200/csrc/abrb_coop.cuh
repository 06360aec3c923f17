// abrb_coop.cuh — the warp-cooperative truncating pseudo-inverse of the OSC kernels (device only).
//
// A few percent of the states (3.8 % of uniformly random UR5 6-DOF states, i.e. 70 % of the warps hold at least one)
// have a task-space inertia whose inverse the reference replaces by pinv(rcond = 1e-4)
// (/root/reference/abr_control/controllers/osc.py:138-145).  With one state per thread those lanes used to walk a
// ~20 us serial eigen-route while the other 30 lanes of their warp idled (47 % of the headline kernel's time).
// Here the whole warp takes that route TOGETHER: the lanes that need it are found with a ballot, and each group of
// six (eight for 7-joint arms) lanes works on one such state — lane r of the group owns row r of A = (L^-1 J^T)^T and
// the r-th entries of the rotated right-hand sides, the disjoint row pairs of one round of a one-sided Jacobi SVD
// (abrb_math.cuh, jacobi_pair) exchange their rows with shuffles, so a round is one parallel step and five (four)
// states are decomposed per pass.  The loops are
// rolled (a few hundred instructions in all), every lane is active, and the owner lanes get Mx y and Mx z back
// through the warp's exchange area in shared memory.
#pragma once
#include "abrb_math.cuh"

namespace abrb {

// Lanes per state: one lane per row of A (KD <= 6 of them) and, in the deferred variant, per row of du = -L w (N of them).
// Six-lane groups put FIVE states on a warp (lanes 30, 31 spare), eight-lane groups four: a CTA of four warps then
// empties up to 20 queued records in one pass instead of 16, which is what keeps a two-tile CTA of the 65 536-state UR5
// batch (9.6 records on average) out of the two-records-per-group pass that used to be the kernel's tail.
template <int N, int KD>
struct CoopGroup {
  static constexpr int kW = KD > N ? KD : N;
  static constexpr int kLanes = kW <= 6 ? 6 : 8;
  static constexpr int kPerWarp = 32 / kLanes;
  // lane -> (group, row in the group, first lane of the group); the spare lanes act as extra non-owners of the last group
  static __device__ __forceinline__ int group(int lane) { return lane / kLanes < kPerWarp ? lane / kLanes : kPerWarp; }
  static __device__ __forceinline__ int sub(int lane) { return lane / kLanes < kPerWarp ? lane % kLanes : kLanes; }
  static __device__ __forceinline__ int base(int lane) {
    return (lane / kLanes < kPerWarp ? lane / kLanes : kPerWarp - 1) * kLanes;
  }
};

// Exchange area of one warp (shared memory, slot-major like the kinematic scratch: value i of lane l at [i * 32 + l]):
//   [0, W)         in: y (KD values)   out: A^T Mx y (N values)        W = max(KD, N)
//   [W, 2 W)       in: z               out: A^T Mx z
//   [2 W, ...)     A(r, k) at 2 W + r * N + k     (only when the scratch lives in registers: COPY_A)
template <int N, int KD, bool COPY_A>
struct CoopLayout {
  static constexpr int kW = KD > N ? KD : N;
  static constexpr int kY = 0, kZ = kW, kA = 2 * kW;
  static constexpr int kSlots = 2 * kW + (COPY_A ? KD * N : 0);
};

// One-sided Jacobi SVD of NS independent KD x N matrices (NS = 1 or 2, interleaved) whose row `sub` and right-hand-side
// entries y[sub], z[sub] this lane loaded into `me[s]` (lanes sub >= KD of the group, and the spare lanes, hold zeros), followed by
// w = A^T pinv(A A^T, rcond) y for the two right-hand sides: on return EVERY lane of the group holds wy[s][] (and
// wz[s][] if `two`).  All 32 lanes of the warp must call it together.
template <int N, int KD, int NS>
__device__ __forceinline__ void coop_jacobi_group(JacobiRow<N, KD> (&me)[NS], int sub, int gbase, const bool (&owner)[NS],
                                                  double rcond, bool two, double (&wy)[NS][N], double (&wz)[NS][N]) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int NRR = KD + (KD & 1);
  JacobiRow<N, KD> other[NS];
#pragma unroll 1
  for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
    bool big = false;
#pragma unroll 1
    for (int r = 0; r < NRR - 1; ++r) {
      const int p = sub < NRR ? rr_partner(NRR, sub, r) : 0;
      const int src = gbase + p;  // (lanes that own no row read some lane of their group; the value is not used)
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int k = 0; k < N; ++k) other[s].b[k] = __shfl_sync(kFull, me[s].b[k], src);
        other[s].t[0] = __shfl_sync(kFull, me[s].t[0], src);
        other[s].t[1] = __shfl_sync(kFull, me[s].t[1], src);
      }
      if (sub < NRR) {
#pragma unroll
        for (int s = 0; s < NS; ++s) big = (jacobi_pair<N, KD>(sub < p, me[s], other[s]) == 2) || big;
      }
    }
    if (!__any_sync(kFull, big)) break;
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    // squared singular values, the cut-off (numpy.linalg.pinv: s <= rcond * max(s) is dropped) and the two products
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) s2 += me[s].b[k] * me[s].b[k];
    double smax = 0.0;  // (the row lanes of the group are gbase .. gbase + KD - 1, whatever the group width)
#pragma unroll
    for (int r = 0; r < KD; ++r) {
      const double t = __shfl_sync(kFull, s2, gbase + r);
      smax = t > smax ? t : smax;
    }
    const bool keep = owner[s] && s2 > rcond * smax;
    const double is2 = keep ? inv_t(s2) : 0.0;
    const double cy = me[s].t[0] * is2, cz = me[s].t[1] * is2;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double a = cy * me[s].b[k], c = cz * me[s].b[k];
      double sa = 0.0, sc = 0.0;
#pragma unroll
      for (int r = 0; r < KD; ++r) {
        sa += __shfl_sync(kFull, a, gbase + r);
        if (two) sc += __shfl_sync(kFull, c, gbase + r);
      }
      wy[s][k] = sa;
      wz[s][k] = sc;
    }
  }
}

// Decompose the states of the lanes in `mask` (warp-uniform, non-zero) IN LINE: group g of the warp takes the g-th
// waiting lane, four or five states per pass.  `ASlot`: where row r, column k of A of lane o is found:
// abase[ASlot::at(r, k) * 32 + o].  y, z are read from xyz[(LY::kY / kZ + r) * 32 + o], and A^T Mx y, A^T Mx z written
// over them.  Not inlined: the hot path only pays a call, and the routine's registers are its own.
template <typename T, int N, int KD, class ASlot, class LY>
__device__ __noinline__ void coop_pinv_warp(unsigned mask, const T *abase, T *xyz, int lane, double rcond, bool two) {
  typedef CoopGroup<N, KD> G;
  const int sub = G::sub(lane), grp = G::group(lane), gbase = G::base(lane);
  while (mask != 0u) {
    const unsigned found = grp < G::kPerWarp ? __fns(mask, 0u, grp + 1) : 0xffffffffu;  // the (grp+1)-th waiting lane, if any
    const bool have = found != 0xffffffffu;
    const int o = have ? (int)found : 0;
    JacobiRow<N, KD> me[1];
    const bool owner[1] = {have && sub < KD};
    const int row = sub < KD ? sub : 0;
#pragma unroll
    for (int k = 0; k < N; ++k) me[0].b[k] = owner[0] ? double(abase[ASlot::at(row, k) * 32 + o]) : 0.0;
    me[0].t[0] = owner[0] ? double(xyz[(LY::kY + row) * 32 + o]) : 0.0;
    me[0].t[1] = (owner[0] && two) ? double(xyz[(LY::kZ + row) * 32 + o]) : 0.0;
    double wy[1][N], wz[1][N];
    __syncwarp();  // every lane of the group has read y, z before they are overwritten
    coop_jacobi_group<N, KD, 1>(me, sub, gbase, owner, rcond, two, wy, wz);
    if (have && sub == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        xyz[(LY::kY + k) * 32 + o] = T(wy[0][k]);
        if (two) xyz[(LY::kZ + k) * 32 + o] = T(wz[0][k]);
      }
    }
    // drop the (up to) four or five states of this pass
#pragma unroll
    for (int i = 0; i < G::kPerWarp; ++i) mask &= mask - 1u;
  }
}

// ---- deferred variant.  A pass costs the same ~9 us whether one or four of a warp's groups have a state to work on,
// and a CTA cannot retire before its slowest warp: paying a pass per warp per tile (70 % of the warps of a UR5 batch)
// doubles the kernel.  Instead a waiting lane finishes its evaluation WITHOUT the task-space term and leaves a record
// (A, the Cholesky factor of M, y, z, its row) in a small queue of the CTA; the CTA empties the queue with all its
// groups (20 or 16) at once — after its last tile, or earlier when as many records have gathered — and adds the missing
//   du = -J^T Mx y - J^T Mx J M^-1 u_null = -L A^T (Mx y + Mx z)
// to the rows already written.  Records that do not fit are handled in line as above.
template <int N, int KD>
struct CoopRecord {
  static constexpr int kA = 0, kL = KD * N, kY = kL + N * (N + 1) / 2, kZ = kY + KD, kLen = kZ + KD;
};
constexpr int kCoopQueuePerWarp = 8;  // queue capacity of a CTA: eight records per warp

template <typename T>
struct FlushOut {
  T *u, *train;           // local (B, n) outputs (u may be null when only the gathered copy is wanted)
  T *peer[kMaxPeers];     // gathered arrays (fused all-gather), row0 = first row of this rank's block
  int n_peer, self;
  int64_t row0;
};

// One round of the CTA's groups over the queue: group j of the CTA takes records j, j + G, ... (NS at a time).
template <typename T, int N, int KD, int NS>
__device__ __forceinline__ void coop_flush_round(const T *qrec, const long long *qrow, int n, int first, int stride,
                                                 const FlushOut<T> &o, double rcond, bool two) {
  typedef CoopRecord<N, KD> RC;
  const int lane = threadIdx.x & 31;
  typedef CoopGroup<N, KD> G;
  const int sub = G::sub(lane), gbase = G::base(lane);
  const bool spare = G::group(lane) >= G::kPerWarp;
  JacobiRow<N, KD> me[NS];
  bool have[NS], owner[NS];
  const T *rec[NS];
  double wy[NS][N], wz[NS][N];
  const int row_ = sub < KD ? sub : 0;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int e = first + s * stride;
    have[s] = e < n && !spare;
    owner[s] = have[s] && sub < KD;
    rec[s] = qrec + (size_t)(have[s] ? e : 0) * RC::kLen;
#pragma unroll
    for (int k = 0; k < N; ++k) me[s].b[k] = owner[s] ? double(rec[s][RC::kA + row_ * N + k]) : 0.0;
    me[s].t[0] = owner[s] ? double(rec[s][RC::kY + row_]) : 0.0;
    me[s].t[1] = (owner[s] && two) ? double(rec[s][RC::kZ + row_]) : 0.0;
  }
  coop_jacobi_group<N, KD, NS>(me, sub, gbase, owner, rcond, two, wy, wz);
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    // du = -L (A^T Mx y + A^T Mx z): every lane holds the two vectors, lane i forms row i of the triangular product
    if (have[s] && sub < N) {
      double dy = 0.0, dz = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (k <= sub) {
          const double l = double(rec[s][RC::kL + sub * (sub + 1) / 2 + k]);
          dy += l * wy[s][k];
          dz += l * wz[s][k];
        }
      }
      const int64_t row = qrow[first + s * stride];
      const T *src = o.u != nullptr ? o.u + row * N + sub : o.peer[o.self] + (o.row0 + row) * N + sub;
      const T v = T(double(*src) - dy - (two ? dz : 0.0));
      if (o.u != nullptr) o.u[row * N + sub] = v;
      if (o.train != nullptr) o.train[row * N + sub] = T(double(o.train[row * N + sub]) - dy);
      for (int p = 0; p < o.n_peer; ++p) o.peer[p][(o.row0 + row) * N + sub] = v;
    }
  }
}

// Empties the CTA's queue (all threads of the CTA call it).  Up to one record per group it is one Jacobi pass; with
// more, each group takes two records through the pass together (the rounds are latency-bound, two interleaved chains
// cost ~1.2x one) — a CTA whose queue happens to hold more records than it has groups is otherwise the kernel's tail.
template <typename T, int N, int KD>
__device__ __noinline__ void coop_flush_cta(const T *qrec, const long long *qrow, int n, const FlushOut<T> &o, double rcond,
                                            bool two) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  constexpr int kPerWarp = CoopGroup<N, KD>::kPerWarp;
  const int groups = n_warps * kPerWarp, mine = warp * kPerWarp + CoopGroup<N, KD>::group(lane);
  if (n <= groups) {
    if (warp * kPerWarp < n) coop_flush_round<T, N, KD, 1>(qrec, qrow, n, mine, groups, o, rcond, two);  // warp-uniform
  } else {
    for (int base = 0; base < n; base += 2 * groups)  // CTA-uniform trip count
      if (base + warp * kPerWarp < n) coop_flush_round<T, N, KD, 2>(qrec, qrow, n, base + mine, groups, o, rcond, two);
  }
}

// The `coop` argument of osc_eval on the GPU (abrb_osc.cuh).  K_: the kernel's kinematic scratch type; when it lives
// in shared memory the routine reads A in place, otherwise the waiting lanes first copy A into the exchange area.
template <typename T, int N, int KD, class K_>
struct WarpCoop {
  static constexpr bool kCopyA = !K_::kSharedScratch;
  typedef CoopLayout<N, KD, kCopyA> LY;
  struct InScratch {
    static __device__ __forceinline__ int at(int r, int k) { return K_::aslot(r, k); }
  };
  struct InExchange {
    static __device__ __forceinline__ int at(int r, int k) { return LY::kA + r * N + k; }
  };
  T *xch;            // the warp's exchange area
  const T *scratch;  // the warp's kinematic scratch (lane 0's column) when it is in shared memory
  int lane;
  bool valid;        // this lane evaluates a state of its own (false: padding lane of a ragged last warp)
  // deferral (osc_kernel only; null queue = always in line)
  T *qrec = nullptr;
  long long *qrow = nullptr;
  int *qcount = nullptr;
  long long row = 0;  // this lane's state index
  int qcap = 0;       // queue capacity (records)

  template <typename T_, int N_, int KD_, class LGet>
  __device__ __forceinline__ void pinv(bool slow, K_ &K, LGet L, const T *y, const T *z, T *wy, T *wz, bool two,
                                       double rcond) {
    slow = slow && valid;
    unsigned mask = __ballot_sync(0xffffffffu, slow);
    if (mask == 0u) return;  // warp-uniform
    if (qrec != nullptr) {
      typedef CoopRecord<N, KD> RC;
      bool queued = false;
      if (slow) {
        const int pos = atomicAdd(qcount, 1);
        if (pos < qcap) {
          queued = true;
          T *rec = qrec + (size_t)pos * RC::kLen;
          qrow[pos] = row;
#pragma unroll
          for (int r = 0; r < KD; ++r) {
            rec[RC::kY + r] = y[r];
            rec[RC::kZ + r] = z[r];
#pragma unroll
            for (int k = 0; k < N; ++k) rec[RC::kA + r * N + k] = K.s.ld(K_::aslot(r, k));
          }
          int li = 0;
#pragma unroll
          for (int a = 0; a < N; ++a)
#pragma unroll
            for (int b = 0; b < N; ++b)
              if (b <= a) rec[RC::kL + li++] = L(a, b);
#pragma unroll
          for (int k = 0; k < N; ++k) {  // the owner finishes without the task-space term; the flush adds it to the stored row
            wy[k] = T(0);
            wz[k] = T(0);
          }
        }
      }
      slow = slow && !queued;
      mask = __ballot_sync(0xffffffffu, slow);
      if (mask == 0u) return;
    }
    if (slow) {
#pragma unroll
      for (int r = 0; r < KD; ++r) {
        xch[(LY::kY + r) * 32 + lane] = y[r];
        xch[(LY::kZ + r) * 32 + lane] = z[r];
      }
      if (kCopyA) {
#pragma unroll
        for (int r = 0; r < KD; ++r)
#pragma unroll
          for (int k = 0; k < N; ++k) xch[InExchange::at(r, k) * 32 + lane] = K.s.ld(K_::aslot(r, k));
      }
    }
    __syncwarp();
    if (kCopyA)
      coop_pinv_warp<T, N, KD, InExchange, LY>(mask, xch, xch, lane, rcond, two);
    else
      coop_pinv_warp<T, N, KD, InScratch, LY>(mask, scratch, xch, lane, rcond, two);
    __syncwarp();
    if (slow) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        wy[k] = xch[(LY::kY + k) * 32 + lane];
        wz[k] = two ? xch[(LY::kZ + k) * 32 + lane] : T(0);
      }
    }
  }
};

}  // namespace abrb
