// abrb_osc.cuh — one state of OSC.generate (and the secondary controllers), fully fused:
// chain walk -> J, M, g, (C dq) -> Cholesky(M) -> task-space inertia -> task PD -> joint torques ->
// null-space filtered secondary torques.  Reference: /root/reference/abr_control/controllers/osc.py:217-320,
// damping.py:21-32, resting_config.py:25-42 + joint.py:104-131, avoid_obstacles.py:38-120.
#pragma once
#include "abrb_math.cuh"

namespace abrb {

enum { kNullDamping = 1, kNullResting = 2, kNullAvoid = 3 };

template <typename T, int N>
struct NullK {
  int kind;
  int n_obs;
  unsigned rest_mask;
  int pad_;
  T kp, kv;
  T rest[N];
  T threshold, gain, maximum;
  T obs[kMaxObstacles][4];
};

template <typename T, int N>
struct OscK {
  T kp, ko, kv;
  T lim_xyz, lim_abg;  // vmax[0]/kp*kv, vmax[1]/ko*kv   (osc.py:109-115)
  T thr;               // |det| threshold of _Mx (osc.py:120,138)
  T xoff[3];
  unsigned dof_mask;   // bit r = ctrlr_dof[r]
  int use_vmax, use_g, use_C, alg, n_null, frame;
  NullK<T, N> nul[kMaxNull];
};

// AvoidObstacles.generate — the rare, data-dependent secondary controller; not inlined and self-contained
// (re-walks the chain) so that the main path's register allocation is unaffected.
// Lm: Cholesky factor of M (row-major N x N, lower).
template <typename T, int N, bool ORTHO>
ABRB_HD_NOINLINE void avoid_generate(const ChainK<T, N> &P, const NullK<T, N> &A, const T *q, const T *Lm,
                                     T *u_out) {
  Kin<T, N, ORTHO> K;  // private register/local copy: this path is rare
  T LFs[N][12];
  walk<T, N>(P, q, 2 * N + 1, K, LFs);  // K.F = EE frame, LFs[i] = link(i+1) frame
  T up[N];
  for (int k = 0; k < N; ++k) up[k] = T(0);
  const T thr = A.threshold;
  for (int seg = 0; seg < N; ++seg) {
    const T *LF = LFs[seg];
    T p1[3], p2[3];
    K.t(seg, p1);
    K.t(seg + 1 < N ? seg + 1 : N - 1, p2);
    if (seg == N - 1)
      for (int r = 0; r < 3; ++r) p2[r] = K.F[r * 4 + 3];
    for (int ob = 0; ob < A.n_obs; ++ob) {
      const T *O = A.obs[ob];
      T line[3], obl[3];
      for (int r = 0; r < 3; ++r) {
        line[r] = p2[r] - p1[r];
        obl[r] = O[r] - p1[r];
      }
      const T proj = dot3(obl, line) / dot3(line, line);
      T cl[3];
      for (int r = 0; r < 3; ++r) cl[r] = proj < T(0) ? p1[r] : (proj > T(1) ? p2[r] : p1[r] + proj * line[r]);
      T d[3] = {O[0] - cl[0], O[1] - cl[1], O[2] - cl[2]};
      const T dist = sqrt_t(dot3(d, d));
      T rho = dist - O[3];
      const T floor_ = thr / T(50);
      rho = rho > floor_ ? rho : floor_;
      if (!(rho < thr)) continue;
      const T mag = T(0.02) * (T(1) / rho - T(1) / thr) * T(1) / (rho * sqrt_t(rho));
      T F[3];
      for (int r = 0; r < 3; ++r) F[r] = mag * (d[r] / rho);
      // m = T_inv(link) [closest;1] with the reference's TRANSPOSE inverse (base_config.py:820-824)
      T dl[3] = {cl[0] - LF[3], cl[1] - LF[7], cl[2] - LF[11]};
      T m[3], pw[3];
      for (int cc = 0; cc < 3; ++cc) m[cc] = LF[0 * 4 + cc] * dl[0] + LF[1 * 4 + cc] * dl[1] + LF[2 * 4 + cc] * dl[2];
      frame_point(LF, m, pw);
      // Jp (3 x N) of that point, W = L^-1 Jp^T (N x 3)
      T Jp[3][N], Wc[3][N];
      for (int k = 0; k < N; ++k) {
        T tk[3], v[3];
        K.t(k, tk);
        T dd[3] = {pw[0] - tk[0], pw[1] - tk[1], pw[2] - tk[2]};
        omega_apply(K, k, dd, v);
        for (int r = 0; r < 3; ++r) Jp[r][k] = k < seg + 1 ? v[r] : T(0);
      }
      for (int r = 0; r < 3; ++r) {
        for (int i = 0; i < N; ++i) {
          T sacc = Jp[r][i];
          for (int k = 0; k < i; ++k) sacc -= Lm[i * N + k] * Wc[r][k];
          Wc[r][i] = sacc / Lm[i * N + i];
        }
      }
      T S3[9];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          T sacc = T(0);
          for (int k = 0; k < N; ++k) sacc += Wc[a][k] * Wc[b][k];
          S3[a * 3 + b] = sacc;
        }
      T x3[3];
      pinv_apply_sym<T, 3>(S3, 7u, T(0.01), F, x3);
      for (int k = 0; k < N; ++k) up[k] -= Jp[0][k] * x3[0] + Jp[1][k] * x3[1] + Jp[2][k] * x3[2];
    }
  }
  for (int k = 0; k < N; ++k) {
    T v = up[k] * A.gain;
    v = v > A.maximum ? A.maximum : v;
    v = v < -A.maximum ? -A.maximum : v;
    u_out[k] = v;
  }
}

// python-style (x mod 2pi) in [0, 2pi)
template <typename T>
ABRB_HD T wrap_pm_pi(T d) {
  const T two_pi = T(6.283185307179586476925286766559);
  const T pi = T(3.14159265358979323846264338327950288);
  T r = fmod_t(d + pi, two_pi);
  r = r < T(0) ? r + two_pi : r;
  return r - pi;
}

// One OSC evaluation.  KD = 3: only (a subset of) x,y,z controlled; KD = 6: any mask.
// PLANT: also return ddq = M^-1 (u + g - C dq) for the rollout kernel.
// DEFER: if the state needs the truncating-pinv path, return true WITHOUT computing u (and without that path's code
// in the instantiation); the caller queues such states for a second, densely packed launch (they are a few % of
// random UR5 states but would otherwise drag most warps through the divergent slow path).  Returns false when u has
// been produced.
// `K`: caller-provided kinematic scratch (registers or shared memory).  Once the dynamics are done its t_k / z_k
// slots are overwritten IN PLACE by the task Jacobian (column k of J only needs t_k, z_k), which later becomes
// A = (L^-1 J^T)^T; so J, A never occupy registers of their own.
template <typename T, int N, int KD, bool PLANT, bool DEFER = false, class K_>
ABRB_HD bool osc_state(const ChainK<T, N> &P, const OscK<T, N> &O, const T *q, const T *dq, const T *target,
                       const T *tv, T *u, T *train, T *ddq, K_ &K) {
  constexpr bool ORTHO = K_::kOrtho;
  typedef typename K_::S SL;
  K.sync();
  walk<T, N>(P, q, O.frame, K);
  K.sync();
  const int dep = frame_dep<N>(O.frame);
  T pF[3];
  frame_point(K.F, O.xoff, pF);

  // ---- task-space error (osc.py:250-272), needs only the frame
  T err[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  if (O.dof_mask & 7u) {
    ABRB_UNROLL
    for (int c = 0; c < 3; ++c) err[c] = pF[c] - target[c];
  }
  if (KD == 6 && (O.dof_mask & 56u)) {
    T R[9];
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r)
      ABRB_UNROLL
    for (int c = 0; c < 3; ++c) R[r * 3 + c] = K.F[r * 4 + c];
    if (O.alg == 0) {
      T qd[4], qe[4], qr[4];
      quat_from_euler_rxyz(target[3], target[4], target[5], qd);
      const T nd = T(1) / sqrt_t(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
      ABRB_UNROLL
      for (int i = 0; i < 4; ++i) qd[i] *= nd;
      quat_from_R(R, qe);
      qe[1] = -qe[1];
      qe[2] = -qe[2];
      qe[3] = -qe[3];
      quat_mul(qd, qe, qr);
      const T sg = qr[0] > T(0) ? T(1) : (qr[0] < T(0) ? T(-1) : T(0));  // numpy.sign
      ABRB_UNROLL
      for (int c = 0; c < 3; ++c) err[3 + c] = -qr[1 + c] * sg;
    } else {
      T Rd[9], Red[9], qed[4];
      R_from_euler_rxyz(target[3], target[4], target[5], Rd);
      ABRB_UNROLL
      for (int r = 0; r < 3; ++r)
        ABRB_UNROLL
      for (int c = 0; c < 3; ++c) Red[r * 3 + c] = R[0 * 3 + r] * Rd[0 * 3 + c] + R[1 * 3 + r] * Rd[1 * 3 + c] + R[2 * 3 + r] * Rd[2 * 3 + c];
      quat_from_R(Red, qed);
      ABRB_UNROLL
      for (int r = 0; r < 3; ++r) err[3 + r] = -(R[r * 3 + 0] * qed[1] + R[r * 3 + 1] * qed[2] + R[r * 3 + 2] * qed[3]);
    }
  }
  if (O.use_vmax) {  // osc.py:198-215
    const T nx = sqrt_t(err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
    const T na = sqrt_t(err[3] * err[3] + err[4] * err[4] + err[5] * err[5]);
    const T sx = nx > O.lim_xyz ? O.lim_xyz / nx : T(1);
    const T sa = na > O.lim_abg ? O.lim_abg / na : T(1);
    ABRB_UNROLL
    for (int c = 0; c < 3; ++c) {
      err[c] = O.kv * sx * (O.kp / O.kv) * err[c];
      err[3 + c] = O.kv * sa * (O.ko / O.kv) * err[3 + c];
    }
  } else {
    ABRB_UNROLL
    for (int c = 0; c < 3; ++c) {
      err[c] *= O.kp;
      err[3 + c] *= O.ko;
    }
  }

  // ---- joint-space dynamics
  K.sync();
  T M[N][N], g[N], cdq[N];
  if (PLANT || O.use_C)
    dynamics_Mg<T, N, true>(P, K, dq, M, g, cdq);
  else
    dynamics_Mg<T, N, false>(P, K, dq, M, g, nullptr);
  ABRB_UNROLL
  for (int a = 0; a < N; ++a)
    ABRB_UNROLL
  for (int b = 0; b < N; ++b)
    if (b < a) M[a][b] = M[b][a];

  // secondary controllers that are M.(something): accumulate the something
  T wn[N];
  bool any_null = false, any_avoid = false;
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) wn[k] = T(0);
  for (int i = 0; i < O.n_null; ++i) {
    const NullK<T, N> &Z = O.nul[i];
    any_null = true;
    if (Z.kind == kNullDamping) {
      ABRB_UNROLL
      for (int k = 0; k < N; ++k) wn[k] -= Z.kv * dq[k];
    } else if (Z.kind == kNullResting) {
      ABRB_UNROLL
      for (int k = 0; k < N; ++k) {
        const T qt = ((Z.rest_mask >> k) & 1u) ? wrap_pm_pi(Z.rest[k] - q[k]) : T(0);
        wn[k] += Z.kp * qt - Z.kv * dq[k];
      }
    } else {
      any_avoid = true;
    }
  }
  // velocity compensation (osc.py:275-282): joint space if the target velocity is all zero
  bool tv_zero = true;
  if (tv != nullptr) {
    ABRB_UNROLL
    for (int c = 0; c < 6; ++c) tv_zero = tv_zero && (tv[c] == T(0));
  }
  T un[N];
  ABRB_UNROLL
  for (int a = 0; a < N; ++a) {
    T s1 = T(0), s2 = T(0);
    ABRB_UNROLL
    for (int b = 0; b < N; ++b) {
      s1 += M[a][b] * dq[b];
      s2 += M[a][b] * wn[b];
    }
    u[a] = tv_zero ? -O.kv * s1 : T(0);
    un[a] = s2;
  }

  K.sync();
  // ---- task Jacobian rows of the controlled DOF written in place over t_k / z_k (osc.py:242-244):
  //      A(r,k): r<3 -> slot kT+3k+r,  r>=3 -> slot kZ+3k+r-3.  Uncontrolled rows are zero.
  auto Aslot = [](int r, int k) { return r < 3 ? SL::kT + 3 * k + r : SL::kZ + 3 * k + (r - 3); };
  T xdot[KD];
  ABRB_UNROLL
  for (int r = 0; r < KD; ++r) xdot[r] = T(0);
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) {
    T tk[3], zk[3], d[3], v[3];
    K.t(k, tk);
    K.z(k, zk);
    ABRB_UNROLL
    for (int c = 0; c < 3; ++c) d[c] = pF[c] - tk[c];
    omega_apply(K, k, d, v);
    const bool on = k < dep;
    ABRB_UNROLL
    for (int r = 0; r < KD; ++r) {
      const T val = (on && ((O.dof_mask >> r) & 1u)) ? (r < 3 ? v[r < 3 ? r : 0] : zk[r < 3 ? 0 : r - 3]) : T(0);
      K.s.st(Aslot(r, k), val);
      xdot[r] += val * dq[k];
    }
  }
  if (!tv_zero) {
    ABRB_UNROLL
    for (int r = 0; r < KD; ++r) err[r] += O.kv * (xdot[r] - tv[r]);
  }
  T y[KD];
  ABRB_UNROLL
  for (int r = 0; r < KD; ++r) y[r] = ((O.dof_mask >> r) & 1u) ? err[r] : T(0);

  K.sync();
  // ---- M = L L^T ;  A <- rows of (L^-1 J^T)^T ;  S = J M^-1 J^T = A A^T   (osc.py:136-137)
  T Mi[N];
  chol<T, N>(M, Mi);
  ABRB_UNROLL
  for (int r = 0; r < KD; ++r) {
    T row[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) row[k] = K.s.ld(Aslot(r, k));
    fwd_solve<T, N>(M, Mi, row);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) K.s.st(Aslot(r, k), row[k]);
  }
  // S is built straight into the array that is then factorised in place; only its trace is kept.  The rare
  // truncating branch re-forms S from A (shared memory) instead of keeping 36 more values live on the hot path.
  T Sc[KD][KD], Si[KD], trS = T(0);
  ABRB_UNROLL
  for (int a = 0; a < KD; ++a) {
    T ra[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) ra[k] = K.s.ld(Aslot(a, k));
    ABRB_UNROLL
    for (int b = 0; b < KD; ++b) {
      if (b <= a) {
        T s = T(0);
        ABRB_UNROLL
        for (int k = 0; k < N; ++k) s += ra[k] * K.s.ld(Aslot(b, k));
        const bool on = ((O.dof_mask >> a) & 1u) && ((O.dof_mask >> b) & 1u);
        Sc[a][b] = on ? s : (a == b ? T(1) : T(0));
        Sc[b][a] = Sc[a][b];
        if (a == b && on) trS += s;
      }
    }
  }
  // ---- Mx: inverse if |det| >= threshold else pinv(rcond = threshold*0.1)   (osc.py:138-145)
  const bool pd = chol<T, KD>(Sc, Si);
  T det = T(1);
  ABRB_UNROLL
  for (int a = 0; a < KD; ++a) det *= Sc[a][a] * Sc[a][a];
  bool fast = pd && (det >= O.thr);
  const T rcond = O.thr * T(0.1);
  if (pd && !fast) {
    // pinv == inv whenever no eigenvalue is truncated; certify that cheaply:
    // lambda_max <= trace(S_active), 1/lambda_min <= ||S^-1||_F  =>  no truncation if 1/||S^-1||_F > rcond*trace
    const T tr = trS;
    T fro = T(0);
    ABRB_UNROLL
    for (int a = 0; a < KD; ++a) {
      if ((O.dof_mask >> a) & 1u) {
        T e[KD];
        ABRB_UNROLL
        for (int b = 0; b < KD; ++b) e[b] = b == a ? T(1) : T(0);
        fwd_solve<T, KD>(Sc, Si, e);
        bwd_solve<T, KD>(Sc, Si, e);
        ABRB_UNROLL
        for (int b = 0; b < KD; ++b) fro += e[b] * e[b];
      }
    }
    fast = rcond * tr * sqrt_t(fro) < T(1);
  }
  if (DEFER && !fast) return true;
  // ---- secondary controllers that go through the null-space filter  I - J^T Mx J M^-1  (osc.py:310-318): their
  //      task-space image z = J M^-1 u_null is formed here so that Mx is applied to y and z in ONE place (the
  //      truncating branch computes its eigenvectors once for both right-hand sides)
  T z[KD];
  ABRB_UNROLL
  for (int r = 0; r < KD; ++r) z[r] = T(0);
  if (any_null) {
    if (any_avoid) {
      T Lf[N * N];
      ABRB_UNROLL
      for (int a = 0; a < N; ++a)
        ABRB_UNROLL
      for (int b = 0; b < N; ++b) Lf[a * N + b] = M[a][b];
      for (int i = 0; i < O.n_null; ++i) {
        if (O.nul[i].kind == kNullAvoid) {
          T ua[N], qa[N];  // private copies: only these (not the caller's register arrays) have their address taken
          ABRB_UNROLL
          for (int k = 0; k < N; ++k) qa[k] = q[k];
          avoid_generate<T, N, ORTHO>(P, O.nul[i], qa, Lf, ua);
          ABRB_UNROLL
          for (int k = 0; k < N; ++k) un[k] += ua[k];
        }
      }
    }
    T w[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) w[k] = un[k];
    fwd_solve<T, N>(M, Mi, w);  // L^-1 u_null
    ABRB_UNROLL
    for (int r = 0; r < KD; ++r) {
      T s = T(0);
      ABRB_UNROLL
      for (int k = 0; k < N; ++k) s += K.s.ld(Aslot(r, k)) * w[k];
      z[r] = ((O.dof_mask >> r) & 1u) ? s : T(0);
    }
  }
  // y <- Mx y,  z <- Mx z
  if (DEFER || fast) {
    fwd_solve<T, KD>(Sc, Si, y);
    bwd_solve<T, KD>(Sc, Si, y);
    if (any_null) {
      fwd_solve<T, KD>(Sc, Si, z);
      bwd_solve<T, KD>(Sc, Si, z);
    }
  } else {
    // cheap register-resident route first (inertia counts + inverse iteration, abrb_math.cuh); the rolled,
    // local-memory Jacobi eigen-decomposition only when that is inconclusive.  Static indices everywhere: a
    // rolled loop over S or v here would force them into local memory for the whole function.
    const unsigned mask = O.dof_mask & ((1u << KD) - 1u);
    // This branch always runs in double precision, also for the fp32 kernels: the matrices that end up here have
    // eigenvalue ratios down to 1e-8, where a float Cholesky breaks down (and the FP64 pipe is idle there anyway).
    double Sd[KD][KD], Ld[KD][KD], Sid[KD], yd[KD], xd[KD], zd[KD], xz[KD];
    const double trd = double(trS);
    ABRB_UNROLL
    for (int a = 0; a < KD; ++a) {
      yd[a] = double(y[a]);
      zd[a] = double(z[a]);
      double ra[N];
      ABRB_UNROLL
      for (int k = 0; k < N; ++k) ra[k] = double(K.s.ld(Aslot(a, k)));
      ABRB_UNROLL
      for (int b = 0; b < KD; ++b) {
        if (b <= a) {
          double acc = 0.0;
          ABRB_UNROLL
          for (int k = 0; k < N; ++k) acc += ra[k] * double(K.s.ld(Aslot(b, k)));
          const bool on = ((mask >> a) & 1u) && ((mask >> b) & 1u);
          // inactive rows: decoupled, diagonal >= lambda_max so that they are never counted as truncated
          const double val = on ? acc : (a == b ? trd : 0.0);
          Sd[a][b] = val;
          Sd[b][a] = val;
        }
      }
    }
    ABRB_UNROLL
    for (int a = 0; a < KD; ++a)
      ABRB_UNROLL
    for (int b = 0; b < KD; ++b) Ld[a][b] = Sd[a][b];
    const bool pdd = chol<double, KD>(Ld, Sid);
    bool done = pdd && pinv_solve_fast2<double, KD>(Sd, Ld, Sid, mask, double(rcond), yd, xd, zd, xz, any_null);
    if (!done) {
      double Sf[KD * KD], yf[KD], xf[KD];
      ABRB_UNROLL
      for (int a = 0; a < KD; ++a)
        ABRB_UNROLL
      for (int b = 0; b < KD; ++b) Sf[a * KD + b] = (a == b && !((mask >> a) & 1u)) ? 1.0 : Sd[a][b];
      ABRB_NOUNROLL
      for (int rhs = 0; rhs < (any_null ? 2 : 1); ++rhs) {
        ABRB_UNROLL
        for (int a = 0; a < KD; ++a) yf[a] = rhs == 0 ? yd[a] : zd[a];
        pinv_apply_sym<double, KD>(Sf, mask, double(rcond), yf, xf);
        ABRB_UNROLL
        for (int a = 0; a < KD; ++a) {
          if (rhs == 0)
            xd[a] = xf[a];
          else
            xz[a] = xf[a];
        }
      }
    }
    ABRB_UNROLL
    for (int a = 0; a < KD; ++a) {
      y[a] = T(xd[a]);
      if (any_null) z[a] = T(xz[a]);
    }
  }
  // J^T x = L (A^T x)   (osc.py:285-288)
  auto JT_apply = [&](const T *x, T *out) {
    T w[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) {
      T s = T(0);
      ABRB_UNROLL
      for (int r = 0; r < KD; ++r) s += K.s.ld(Aslot(r, k)) * x[r];
      w[k] = s;
    }
    ABRB_UNROLL
    for (int i = 0; i < N; ++i) {
      T s = T(0);
      ABRB_UNROLL
      for (int k = 0; k < N; ++k)
        if (k <= i) s += M[i][k] * w[k];
      out[i] = s;
    }
  };
  {
    T jt[N];
    JT_apply(y, jt);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] -= jt[k];
  }
  if (O.use_C) {
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] -= cdq[k];
  }
  if (train != nullptr) {
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) train[k] = u[k];  // osc.py:297
  }
  if (O.use_g) {
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] -= g[k];
  }
  // ---- secondary controllers, filtered:  u += u_null - J^T Mx J M^-1 u_null   (osc.py:310-318)
  if (any_null) {
    T jt[N];
    JT_apply(z, jt);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] += un[k] - jt[k];
  }
  if (PLANT) {
    T rhs[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) rhs[k] = u[k] + g[k] - cdq[k];
    fwd_solve<T, N>(M, Mi, rhs);
    bwd_solve<T, N>(M, Mi, rhs);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) ddq[k] = rhs[k];
  }
  return false;
}

// Standalone secondary controller (`Damping/RestingConfig/AvoidObstacles.generate`)
template <typename T, int N, class K_>
ABRB_HD void null_state(const ChainK<T, N> &P, const NullK<T, N> &Z, const T *q, const T *dq, T *u, K_ &K) {
  constexpr bool ORTHO = K_::kOrtho;
  walk<T, N>(P, q, 0, K);
  T M[N][N], g[N];
  dynamics_Mg<T, N, false>(P, K, dq, M, g, nullptr);
  ABRB_UNROLL
  for (int a = 0; a < N; ++a)
    ABRB_UNROLL
  for (int b = 0; b < N; ++b)
    if (b < a) M[a][b] = M[b][a];
  if (Z.kind == kNullAvoid) {
    T Mi[N];
    chol<T, N>(M, Mi);
    T Lf[N * N];
    ABRB_UNROLL
    for (int a = 0; a < N; ++a)
      ABRB_UNROLL
    for (int b = 0; b < N; ++b) Lf[a * N + b] = M[a][b];
    T ua[N], qa[N];
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) qa[k] = q[k];
    avoid_generate<T, N, ORTHO>(P, Z, qa, Lf, ua);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] = ua[k];
    return;
  }
  T w[N];
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) {
    if (Z.kind == kNullDamping) {
      w[k] = -Z.kv * dq[k];
    } else {
      const T qt = ((Z.rest_mask >> k) & 1u) ? wrap_pm_pi(Z.rest[k] - q[k]) : T(0);
      w[k] = Z.kp * qt - Z.kv * dq[k];
    }
  }
  ABRB_UNROLL
  for (int a = 0; a < N; ++a) {
    T s = T(0);
    ABRB_UNROLL
    for (int b = 0; b < N; ++b) s += M[a][b] * w[b];
    u[a] = s;
  }
}

// Joint.generate (controllers/joint.py:104-131)
template <typename T, int N, class K_>
ABRB_HD void joint_state(const ChainK<T, N> &P, T kp, T kv, bool gravity, const T *q, const T *dq, const T *target,
                         const T *tv, T *u, K_ &K) {
  walk<T, N>(P, q, 0, K);
  T M[N][N], g[N];
  dynamics_Mg<T, N, false>(P, K, dq, M, g, nullptr);
  T w[N];
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) w[k] = kp * wrap_pm_pi(target[k] - q[k]) + kv * ((tv != nullptr ? tv[k] : T(0)) - dq[k]);
  ABRB_UNROLL
  for (int a = 0; a < N; ++a) {
    T s = T(0);
    ABRB_UNROLL
    for (int b = 0; b < N; ++b) s += (b >= a ? M[a][b] : M[b][a]) * w[b];
    u[a] = gravity ? s - g[a] : s;
  }
}

// Floating.generate (controllers/floating.py:27-71)
template <typename T, int N, class K_>
ABRB_HD void floating_state(const ChainK<T, N> &P, bool task_space, bool dynamic, const T *q, const T *dq, T *u,
                            K_ &K) {
  constexpr bool ORTHO = K_::kOrtho;
  walk<T, N>(P, q, 2 * N + 1, K);
  T M[N][N], g[N];
  dynamics_Mg<T, N, false>(P, K, dq, M, g, nullptr);
  ABRB_UNROLL
  for (int a = 0; a < N; ++a)
    ABRB_UNROLL
  for (int b = 0; b < N; ++b)
    if (b < a) M[a][b] = M[b][a];
  T Mdq[N];
  ABRB_UNROLL
  for (int a = 0; a < N; ++a) {
    T s = T(0);
    ABRB_UNROLL
    for (int b = 0; b < N; ++b) s += M[a][b] * dq[b];
    Mdq[a] = s;
  }
  if (!task_space) {
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) u[k] = -g[k] - (dynamic ? Mdq[k] : T(0));
    return;
  }
  // J = J("EE")[:3];  A = (L^-1 J^T)^T;  S = J M^-1 J^T;  u = J^T (-Mx^T J M^-1 g) = -L A^T Mx A (L^-1 g)
  T pF[3] = {K.F[3], K.F[7], K.F[11]};
  T J[6][N];
  jacobian<T, N>(K, pF, N, J);
  T Mi[N];
  chol<T, N>(M, Mi);
  T A[3][N];
  ABRB_UNROLL
  for (int r = 0; r < 3; ++r) {
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) A[r][k] = J[r][k];
    fwd_solve<T, N>(M, Mi, A[r]);
  }
  T S[3][3], Sc[3][3], Si[3];
  ABRB_UNROLL
  for (int a = 0; a < 3; ++a)
    ABRB_UNROLL
  for (int b = 0; b < 3; ++b) {
    T s = T(0);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) s += A[a][k] * A[b][k];
    S[a][b] = s;
    Sc[a][b] = s;
  }
  const bool pd = chol<T, 3>(Sc, Si);
  T det = T(1);
  ABRB_UNROLL
  for (int a = 0; a < 3; ++a) det *= Sc[a][a] * Sc[a][a];
  T w[N], z[3];
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) w[k] = g[k];
  fwd_solve<T, N>(M, Mi, w);
  ABRB_UNROLL
  for (int r = 0; r < 3; ++r) {
    T s = T(0);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k) s += A[r][k] * w[k];
    z[r] = s;
  }
  if (pd && det > T(1e-3)) {  // note the strict '>' of floating.py:52 (osc.py uses '>=')
    fwd_solve<T, 3>(Sc, Si, z);
    bwd_solve<T, 3>(Sc, Si, z);
  } else {
    T Sf[9], zi[3], zo[3];
    ABRB_UNROLL
    for (int a = 0; a < 3; ++a) {
      zi[a] = z[a];
      ABRB_UNROLL
      for (int b = 0; b < 3; ++b) Sf[a * 3 + b] = S[a][b];
    }
    pinv_apply_sym<T, 3>(Sf, 7u, T(1e-4), zi, zo);
    ABRB_UNROLL
    for (int a = 0; a < 3; ++a) z[a] = zo[a];
  }
  ABRB_UNROLL
  for (int k = 0; k < N; ++k) {
    T s = T(0);
    ABRB_UNROLL
    for (int r = 0; r < 3; ++r) s += A[r][k] * z[r];
    w[k] = s;
  }
  ABRB_UNROLL
  for (int i = 0; i < N; ++i) {
    T s = T(0);
    ABRB_UNROLL
    for (int k = 0; k < N; ++k)
      if (k <= i) s += M[i][k] * w[k];
    u[i] = -s - (dynamic ? Mdq[i] : T(0));
  }
  (void)ORTHO;
}

}  // namespace abrb
