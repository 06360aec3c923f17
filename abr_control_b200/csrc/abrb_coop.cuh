// abrb_coop.cuh — the warp-cooperative truncating pseudo-inverse of the OSC kernels (device only).
//
// A few percent of the states (3.8 % of uniformly random UR5 6-DOF states, i.e. 70 % of the warps hold at least one)
// have a task-space inertia whose inverse the reference replaces by pinv(rcond = 1e-4)
// (/root/reference/abr_control/controllers/osc.py:138-145).  With one state per thread those lanes used to walk a
// ~20 us serial eigen-route while the other 30 lanes of their warp idled (47 % of the headline kernel's time).
// Here the whole warp takes that route TOGETHER: the lanes that need it are found with a ballot, and each group of
// 8 lanes works on one such state — lane r of the group owns row r of A = (L^-1 J^T)^T and of the accumulated
// rotations, the disjoint row pairs of one round of a one-sided Jacobi SVD (abrb_math.cuh, jacobi_pair) exchange
// their rows with shuffles, so a round is one parallel step and four states are decomposed per pass.  The loops are
// rolled (a few hundred instructions in all), every lane is active, and the owner lanes get Mx y and Mx z back
// through the warp's exchange area in shared memory.
#pragma once
#include "abrb_math.cuh"

namespace abrb {

constexpr int kCoopGroup = 8;  // lanes per state (6 or 3 row owners + idle lanes: shuffles stay inside an 8-lane group)

// Exchange area of one warp (shared memory, slot-major like the kinematic scratch: value i of lane l at [i * 32 + l]):
//   [0, KD)        y  ->  Mx y
//   [KD, 2 KD)     z  ->  Mx z
//   [2 KD, ...)    A(r, k) at 2 KD + r * N + k     (only when the scratch lives in registers: COPY_A)
template <int N, int KD, bool COPY_A>
struct CoopLayout {
  static constexpr int kY = 0, kZ = KD, kA = 2 * KD;
  static constexpr int kSlots = 2 * KD + (COPY_A ? KD * N : 0);
};

// Decompose the states of the lanes in `mask` (warp-uniform, non-zero).  `arow`: where row r, column k of A of lane o
// is found: abase[aslot(r, k) * 32 + o].  Not inlined: the hot path only pays a call, and the routine's registers are
// its own.
template <typename T, int N, int KD, class ASlot>
__device__ __noinline__ void coop_pinv_warp(unsigned mask, const T *abase, T *xyz, int lane, double rcond, bool two) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int NRR = KD + (KD & 1);
  const int sub = lane & (kCoopGroup - 1), grp = lane / kCoopGroup, gbase = lane & ~(kCoopGroup - 1);
  while (mask != 0u) {
    const unsigned found = __fns(mask, 0u, grp + 1);  // the (grp+1)-th waiting lane, if any
    const bool have = found != 0xffffffffu;
    const int o = have ? (int)found : 0;
    JacobiRow<N, KD> me, other;
    const bool owner = have && sub < KD;
#pragma unroll
    for (int k = 0; k < N; ++k) me.b[k] = owner ? double(abase[ASlot::at(sub, k) * 32 + o]) : 0.0;
#pragma unroll
    for (int k = 0; k < KD; ++k) me.v[k] = k == sub ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
      bool big = false;
#pragma unroll 1
      for (int r = 0; r < NRR - 1; ++r) {
        const int p = sub < NRR ? rr_partner(NRR, sub, r) : sub;
#pragma unroll
        for (int k = 0; k < N; ++k) other.b[k] = __shfl_sync(kFull, me.b[k], gbase + p);
#pragma unroll
        for (int k = 0; k < KD; ++k) other.v[k] = __shfl_sync(kFull, me.v[k], gbase + p);
        if (sub < NRR) big = (jacobi_pair<N, KD>(sub < p, me, other) == 2) || big;
      }
      if (!__any_sync(kFull, big)) break;
    }
    // squared singular values, the cut-off (numpy.linalg.pinv: s <= rcond * max(s) is dropped) and the two products
    double s2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) s2 += me.b[k] * me.b[k];
    double smax = s2;
#pragma unroll
    for (int d = 1; d < kCoopGroup; d <<= 1) {
      const double t = __shfl_xor_sync(kFull, smax, d);
      smax = t > smax ? t : smax;
    }
    const bool keep = owner && s2 > rcond * smax;
    double cy = 0.0, cz = 0.0;
    if (keep) {
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        cy += me.v[k] * double(xyz[k * 32 + o]);
        cz += me.v[k] * double(xyz[(KD + k) * 32 + o]);
      }
#if ABRB_FAST_DIV
      const double is2 = inv_t(s2);
#else
      const double is2 = 1.0 / s2;
#endif
      cy *= is2;
      cz *= is2;
    }
    __syncwarp();  // every lane of the group has read y, z before they are overwritten
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      double xy = cy * me.v[k], xz = cz * me.v[k];
#pragma unroll
      for (int d = 1; d < kCoopGroup; d <<= 1) {
        xy += __shfl_xor_sync(kFull, xy, d);
        if (two) xz += __shfl_xor_sync(kFull, xz, d);
      }
      if (have && sub == 0) {
        xyz[k * 32 + o] = T(xy);
        if (two) xyz[(KD + k) * 32 + o] = T(xz);
      }
    }
    // drop the (up to) four states of this pass
#pragma unroll
    for (int i = 0; i < 32 / kCoopGroup; ++i) mask &= mask - 1u;
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

  template <typename T_, int N_, int KD_>
  __device__ __forceinline__ void pinv(bool slow, K_ &K, T *y, T *z, bool two, double rcond) {
    const unsigned mask = __ballot_sync(0xffffffffu, slow);
    if (mask == 0u) return;  // warp-uniform
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
      coop_pinv_warp<T, N, KD, InExchange>(mask, xch, xch, lane, rcond, two);
    else
      coop_pinv_warp<T, N, KD, InScratch>(mask, scratch, xch, lane, rcond, two);
    __syncwarp();
    if (slow) {
#pragma unroll
      for (int r = 0; r < KD; ++r) {
        y[r] = xch[(LY::kY + r) * 32 + lane];
        if (two) z[r] = xch[(LY::kZ + r) * 32 + lane];
      }
    }
  }
};

}  // namespace abrb
