// abrb_rbd.cuh — one state of the batched rigid-body quantities {Tx, T, R, T_inv, quaternion, J, dJ, M, g, C}
// (reference: /root/reference/abr_control/arms/base_config.py:210-415).
#pragma once
#include "abrb_math.cuh"

namespace abrb {

enum : unsigned {
  kWantTx = 1u << 0,
  kWantT = 1u << 1,
  kWantR = 1u << 2,
  kWantTinv = 1u << 3,
  kWantQuat = 1u << 4,
  kWantJ = 1u << 5,
  kWantdJ = 1u << 6,
  kWantM = 1u << 7,
  kWantg = 1u << 8,
  kWantC = 1u << 9,
};

enum { kOutTx = 0, kOutT, kOutR, kOutTinv, kOutQuat, kOutJ, kOutdJ, kOutM, kOutg, kOutC, kOutCount };

// One state.  Every requested quantity is handed to `out.template put<LEN>(which, record)` as soon as it is complete,
// so that no more than one or two output records are ever live in registers (the kernel's `put` stages the record
// through shared memory and writes it out coalesced; the host test shim copies it into an array).
// DYN: M and/or g requested; CMAT: C requested; XTRA: any of Tx/T/R/T_inv/quaternion/dJ requested (compiled out of
// the common {J, M, g, C} instantiations: less code to fetch, fewer live registers).  `want` is uniform over the launch.
// `K` is the caller-provided kinematic scratch (register- or shared-memory-backed, see Kin in abrb_math.cuh).
template <typename T, int N, bool DYN, bool CMAT, bool XTRA, class K_, class Out>
ABRB_HD void rbd_state(const ChainK<T, N> &P, const T *q, const T *dq, int frame, const T *xoff, unsigned want,
                       K_ &K, Out &out) {
  if (!XTRA) want &= (kWantJ | kWantM | kWantg | kWantC);
  K.sync();
  walk<T, N>(P, q, frame, K);
  K.sync();
  const int dep = frame_dep<N>(frame);
  T pF[3];
  frame_point(K.F, xoff, pF);
  if (XTRA && (want & kWantTx)) out.template put<3>(kOutTx, pF);
  if (XTRA && (want & kWantT)) {  // base_config.py:338-369
    T Tm[16];
    ABRB_UNROLL
    for (int i = 0; i < 12; ++i) Tm[i] = K.F[i];
    Tm[12] = Tm[13] = Tm[14] = T(0);
    Tm[15] = T(1);
    out.template put<16>(kOutT, Tm);
  }
  if (XTRA && (want & (kWantR | kWantQuat))) {  // base_config.py:647-676, :304-318
    T R[9];
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r)
      ABRB_UNROLL
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = K.F[r * 4 + c];
    if (want & kWantR) out.template put<9>(kOutR, R);
    if (want & kWantQuat) {
      T qt[4];
      quat_from_R(R, qt);
      out.template put<4>(kOutQuat, qt);
    }
  }
  if (XTRA && (want & kWantTinv)) {  // [[R^T, -R^T t],[0,1]] with the TRANSPOSE (base_config.py:820-824)
    T Ti[16];
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r) {
      T s = T(0);
      ABRB_UNROLL
      for (int c = 0; c < 3; ++c) {
        Ti[r * 4 + c] = K.F[c * 4 + r];
        s -= K.F[c * 4 + r] * K.F[c * 4 + 3];
      }
      Ti[r * 4 + 3] = s;
    }
    Ti[12] = Ti[13] = Ti[14] = T(0);
    Ti[15] = T(1);
    out.template put<16>(kOutTinv, Ti);
  }
  K.sync();
  if (want & (kWantJ | kWantdJ)) {
    T J[6][N];
    jacobian<T, N>(K, pF, dep, J);
    if (want & kWantJ) out.template put<6 * N>(kOutJ, &J[0][0]);
    if (XTRA && (want & kWantdJ)) {
      T dJ[6][N];
      jacobian_dot<T, N>(K, J, dq, dep, dJ);
      out.template put<6 * N>(kOutdJ, &dJ[0][0]);
    }
  }
  if (DYN) {
    T M[N][N], g[N];
    dynamics_Mg<T, N, false>(P, K, dq, M, g, nullptr);
    ABRB_UNROLL
    for (int a = 0; a < N; ++a)
      ABRB_UNROLL
    for (int b = 0; b < N; ++b)
      if (b < a) M[a][b] = M[b][a];
    if (want & kWantM) out.template put<N * N>(kOutM, &M[0][0]);
    if (want & kWantg) out.template put<N>(kOutg, g);
  }
  K.sync();
  if (CMAT) {
    T C[N][N];
    dynamics_C<T, N>(P, K, dq, C);
    out.template put<N * N>(kOutC, &C[0][0]);
  }
}

}  // namespace abrb
