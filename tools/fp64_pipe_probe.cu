// tools/fp64_pipe_probe.cu -- do DADD/DMUL and DMMA.8x8x4 share one FP64 datapath on B200?
// Runs, with W warps per SM on every SM: (A) N independent DADDs per warp, (B) N/8 DMMAs per warp,
// (C) both interleaved; prints cycles per warp-instruction per SM sub-partition.  If time(C) ~ time(A) +
// time(B) the pipes are one datapath; if ~ max(A, B) they are separate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int n) {
  double a[8], c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
  const double b = out[1];
  for (int i = 0; i < 8; ++i) a[i] = out[0] + threadIdx.x * 1e-3 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    if (MODE & 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] + b;  // 32 independent-ish DADDs (8 chains)
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma(c0[i], c1[i], 1.0, b);  // 4 DMMAs (4 chains)
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  out[2 + (blockIdx.x * blockDim.x + threadIdx.x) % 1024] = s;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 16384); cudaMalloc(&c, 64);
  double h[2] = {1.0, 1e-9}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice);
  const int n = 2000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    long long t[4] = {0, 0, 0, 0};
    for (int mode = 1; mode <= 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 1) k<1><<<148, 32 * warps>>>(d, c, n);
        if (mode == 2) k<2><<<148, 32 * warps>>>(d, c, n);
        if (mode == 3) k<3><<<148, 32 * warps>>>(d, c, n);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&t[mode], c, 8, cudaMemcpyDeviceToHost);
    }
    const double wps = warps / 4.0;  // warps per SM sub-partition
    printf("warps/SM=%2d  DADD: %.2f cyc/instr/SMSP   DMMA: %.2f cyc/instr/SMSP   mixed(32 DADD + 4 DMMA): %lld cyc/iter "
           "(DADD alone %lld, DMMA alone %lld, sum %lld)\n",
           warps, t[1] / (double)n / 32 / wps, t[2] / (double)n / 4 / wps, t[3] / n, t[1] / n, t[2] / n, (t[1] + t[2]) / n);
  }
  printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
