// Micro-benchmark: cycles per FP64 FMA warp instruction as a function of the number of active lanes and of the warps
// resident on the SM (is the FP64 pipe's issue time halved for a half-populated warp?  is the unit per sub-partition?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dbg/dfma_lanes tools/dbg/dfma_lanes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double *out, long long *cyc, int active, int iters) {
    const int lane = threadIdx.x & 31;
    double a[CHAINS];
    for (int c = 0; c < CHAINS; ++c) a[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    const double m = 1.0000001, b = 1e-7;
    long long t0 = 0, t1 = 0;
    if (lane < active) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, b);
        }
        t1 = clock64();
    }
    double s = 0;
    for (int c = 0; c < CHAINS; ++c) s += a[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16}) {
        for (int active : {32, 16, 8, 4, 1}) {
            for (int chains : {1, 8}) {
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    if (chains == 1) k<1><<<148, warps * 32>>>(out, cyc, active, iters);
                    else k<8><<<148, warps * 32>>>(out, cyc, active, iters);
                    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
                }
                printf("warps/SM %2d  active lanes %2d  independent chains %d : %.2f cycles per DFMA (per warp)\n", warps,
                       active, chains, double(h) / (double(iters) * chains));
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
