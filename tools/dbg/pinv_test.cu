#include <cstdio>
#include <cuda_runtime.h>
#include "../../abr_control_b200/csrc/abrb_math.cuh"
using namespace abrb;

template <typename T, int S_, int MODE>
__host__ __device__ void jac(const T *Sin, T *w, bool verbose) {
  T A[S_][S_];
  for (int i = 0; i < S_; ++i) for (int j = 0; j < S_; ++j) A[i][j] = Sin[i * S_ + j];
  const T eps = T(1e-30);
  int sweeps = 0;
  if (MODE == 1) {
#pragma unroll 1
    for (int sweep = 0; sweep < 24; ++sweep) {
      T off = 0, diag = 0;
#pragma unroll 1
      for (int i = 0; i < S_; ++i) { diag += A[i][i] * A[i][i];
#pragma unroll 1
        for (int j = i + 1; j < S_; ++j) off += A[i][j] * A[i][j]; }
      if (verbose) printf("  sweep %d off %.3e diag %.3e\n", sweep, (double)off, (double)diag);
      if (off <= eps * diag) break;
      ++sweeps;
#pragma unroll 1
      for (int p = 0; p < S_ - 1; ++p)
#pragma unroll 1
        for (int q = p + 1; q < S_; ++q) {
          const T apq = A[p][q];
          if (apq == T(0)) continue;
          const T theta = (A[q][q] - A[p][p]) / (T(2) * apq);
          const T t = (theta >= T(0) ? T(1) : T(-1)) / (abs_t(theta) + sqrt_t(theta * theta + T(1)));
          const T c = T(1) / sqrt_t(t * t + T(1)), s = t * c;
#pragma unroll 1
          for (int k = 0; k < S_; ++k) { const T akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
#pragma unroll 1
          for (int k = 0; k < S_; ++k) { const T apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        }
    }
  } else {
    for (int sweep = 0; sweep < 24; ++sweep) {
      T off = 0, diag = 0;
      for (int i = 0; i < S_; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < S_; ++j) off += A[i][j] * A[i][j]; }
      if (verbose) printf("  sweep %d off %.3e diag %.3e\n", sweep, (double)off, (double)diag);
      if (off <= eps * diag) break;
      ++sweeps;
      for (int p = 0; p < S_ - 1; ++p)
        for (int q = p + 1; q < S_; ++q) {
          const T apq = A[p][q];
          if (apq == T(0)) continue;
          const T theta = (A[q][q] - A[p][p]) / (T(2) * apq);
          const T t = (theta >= T(0) ? T(1) : T(-1)) / (abs_t(theta) + sqrt_t(theta * theta + T(1)));
          const T c = T(1) / sqrt_t(t * t + T(1)), s = t * c;
          for (int k = 0; k < S_; ++k) { const T akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
          for (int k = 0; k < S_; ++k) { const T apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        }
    }
  }
  for (int i = 0; i < S_; ++i) w[i] = A[i][i];
}
__global__ void kj0(const double* S, double* w) { jac<double,6,0>(S, w, true); }
__global__ void kj1(const double* S, double* w) { jac<double,6,1>(S, w, true); }
__global__ void k(const double* S, const double* y, double* x, unsigned active, double rcond) { pinv_apply_sym<double,6>(S, active, rcond, y, x); }
int main() {
  double l[6] = {3.03e-8, 1.42e-4, 1.57e-3, 5.4e-2, 0.191, 1.0};
  double Q[6][6];
  double v[6] = {0.3, -0.5, 0.2, 0.6, -0.4, 0.3}; double nv=0; for (int i=0;i<6;++i) nv+=v[i]*v[i];
  for (int i=0;i<6;++i) for (int j=0;j<6;++j) Q[i][j] = (i==j) - 2*v[i]*v[j]/nv;
  double S[36], y[6] = {1, -2, 0.5, 3, -1, 0.25}, xh[6], xd[6], wh[6], wd[6];
  for (int i=0;i<6;++i) for (int j=0;j<6;++j) { double s=0; for (int e=0;e<6;++e) s += Q[i][e]*l[e]*Q[j][e]; S[i*6+j]=s*37.0; }
  printf("host:\n"); jac<double,6,0>(S, wh, true);
  for (int i=0;i<6;++i) printf(" %.6e", wh[i]/37); printf("\n");
  pinv_apply_sym<double,6>(S, 63u, 1e-4, y, xh);
  double *dS,*dy,*dx,*dw; cudaMalloc(&dS,sizeof S); cudaMalloc(&dy,sizeof y); cudaMalloc(&dx,sizeof xd); cudaMalloc(&dw,sizeof wd);
  cudaMemcpy(dS,S,sizeof S,cudaMemcpyHostToDevice); cudaMemcpy(dy,y,sizeof y,cudaMemcpyHostToDevice);
  printf("dev mode0:\n"); kj0<<<1,1>>>(dS,dw); cudaDeviceSynchronize(); cudaMemcpy(wd,dw,sizeof wd,cudaMemcpyDeviceToHost);
  for (int i=0;i<6;++i) printf(" %.6e", wd[i]/37); printf("\n");
  printf("dev mode1 (no unroll):\n"); kj1<<<1,1>>>(dS,dw); cudaDeviceSynchronize(); cudaMemcpy(wd,dw,sizeof wd,cudaMemcpyDeviceToHost);
  for (int i=0;i<6;++i) printf(" %.6e", wd[i]/37); printf("\n");
  k<<<1,32>>>(dS,dy,dx,63u,1e-4); cudaError_t e=cudaDeviceSynchronize();
  cudaMemcpy(xd,dx,sizeof xd,cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(e));
  for (int i=0;i<6;++i) printf("%d host % .12e dev % .12e\n", i, xh[i], xd[i]);
  return 0;
}
