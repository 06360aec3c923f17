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

template <typename T, int N>
struct RbdOut {
  T Tx[3];
  T Tm[16];
  T R[9];
  T Tinv[16];
  T quat[4];
  T J[6][N];
  T dJ[6][N];
  T M[N][N];
  T g[N];
  T C[N][N];
};

// DYN: M and/or g requested; CMAT: C requested (implies the dynamics pass).  `want` is uniform over the launch.
// `K` is the caller-provided kinematic scratch (register- or shared-memory-backed, see Kin in abrb_math.cuh).
template <typename T, int N, bool DYN, bool CMAT, class K_>
ABRB_HD void rbd_state(const ChainK<T, N> &P, const T *q, const T *dq, int frame, const T *xoff, unsigned want,
                       RbdOut<T, N> &o, K_ &K) {
  walk<T, N>(P, q, frame, K);
  const int dep = frame_dep<N>(frame);
  T pF[3];
  frame_point(K.F, xoff, pF);
  if (want & kWantTx) {
    ABRB_UNROLL
    for (int c = 0; c < 3; ++c) o.Tx[c] = pF[c];
  }
  if (want & kWantT) {  // base_config.py:338-369
    ABRB_UNROLL
    for (int i = 0; i < 12; ++i) o.Tm[i] = K.F[i];
    o.Tm[12] = o.Tm[13] = o.Tm[14] = T(0);
    o.Tm[15] = T(1);
  }
  if (want & kWantR) {  // base_config.py:647-676
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r)
      ABRB_UNROLL
    for (int c = 0; c < 3; ++c) o.R[r * 3 + c] = K.F[r * 4 + c];
  }
  if (want & kWantTinv) {  // [[R^T, -R^T t],[0,1]] with the TRANSPOSE (base_config.py:820-824)
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r) {
      T s = T(0);
      ABRB_UNROLL
      for (int c = 0; c < 3; ++c) {
        o.Tinv[r * 4 + c] = K.F[c * 4 + r];
        s -= K.F[c * 4 + r] * K.F[c * 4 + 3];
      }
      o.Tinv[r * 4 + 3] = s;
    }
    o.Tinv[12] = o.Tinv[13] = o.Tinv[14] = T(0);
    o.Tinv[15] = T(1);
  }
  if (want & kWantQuat) {  // base_config.py:304-318
    T R[9];
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r)
      ABRB_UNROLL
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = K.F[r * 4 + c];
    quat_from_R(R, o.quat);
  }
  if (want & (kWantJ | kWantdJ)) {
    jacobian<T, N>(K, pF, dep, o.J);
    if (want & kWantdJ) jacobian_dot<T, N>(K, o.J, dq, dep, o.dJ);
  }
  if (DYN || CMAT) {
    dynamics<T, N, CMAT, false>(P, K, dq, o.M, o.g, o.C, nullptr);
    ABRB_UNROLL
    for (int a = 0; a < N; ++a)
      ABRB_UNROLL
    for (int b = 0; b < N; ++b)
      if (b < a) o.M[a][b] = o.M[b][a];
  }
}

}  // namespace abrb
