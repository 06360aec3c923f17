// timing + correctness of pinv_solve_fast on real slow-path matrices (tools/dbg/slow_states.bin), one per thread
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../abr_control_b200/csrc/abrb_math.cuh"
using namespace abrb;
template <typename T>
__global__ void k(const double* data, int n, double* xout, int* okout, long long* cyc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = data + (size_t)i * 48;
  T S[6][6], Sc[6][6], Si[6], y[6], x[6];
  for (int a = 0; a < 6; ++a) { y[a] = T(r[36 + a]); for (int b = 0; b < 6; ++b) S[a][b] = Sc[a][b] = T(r[a * 6 + b]); }
  bool pd = chol<T, 6>(Sc, Si);
  long long t0 = clock64();
  bool ok = pd && pinv_solve_fast<T, 6>(S, Sc, Si, 63u, T(1e-4), y, x);
  long long t1 = clock64();
  okout[i] = ok; cyc[i] = t1 - t0;
  for (int a = 0; a < 6; ++a) xout[i * 6 + a] = ok ? double(x[a]) : 0.0;
}
template <typename T> void run(const std::vector<double>& h, int n, const char* name) {
  double *d, *x; int* ok; long long* cyc;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&x, n * 48); cudaMalloc(&ok, n * 4); cudaMalloc(&cyc, n * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  k<T><<<(n + 31) / 32, 32>>>(d, n, x, ok, cyc); cudaDeviceSynchronize();
  k<T><<<(n + 31) / 32, 32>>>(d, n, x, ok, cyc); cudaError_t e = cudaDeviceSynchronize();
  std::vector<double> hx(n * 6); std::vector<int> hok(n); std::vector<long long> hc(n);
  cudaMemcpy(hx.data(), x, n * 48, cudaMemcpyDeviceToHost); cudaMemcpy(hok.data(), ok, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc.data(), cyc, n * 8, cudaMemcpyDeviceToHost);
  int nok = 0; double worst = 0; long long cmax = 0, csum = 0;
  for (int i = 0; i < n; ++i) { nok += hok[i]; csum += hc[i]; if (hc[i] > cmax) cmax = hc[i];
    if (hok[i]) { double sc = 0, er = 0; for (int a = 0; a < 6; ++a) { double rf = h[(size_t)i * 48 + 42 + a]; sc = fmax(sc, fabs(rf)); er = fmax(er, fabs(hx[i * 6 + a] - rf)); } worst = fmax(worst, er / sc); } }
  printf("%s: err=%s ok %d / %d, worst rel err %.2e, cycles avg %lld max %lld\n", name, cudaGetErrorString(e), nok, n, worst, csum / n, cmax);
}
int main() {
  FILE* f = fopen("tools/dbg/slow_states.bin", "rb"); if (!f) { printf("no data\n"); return 1; }
  std::vector<double> h; double buf[48]; while (fread(buf, 8, 48, f) == 48) h.insert(h.end(), buf, buf + 48); fclose(f);
  int n = h.size() / 48;
  run<double>(h, n, "double"); run<float>(h, n, "float");
  return 0;
}
