// tools/dmma_probe.cu -- what does mma.sync.m8n8k4.f64 compute, bit for bit?
// Compares D = A*B + C from the tensor core with candidate CPU evaluation orders
// and measures the dependent-issue latency of a DMMA chain.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__global__ void run(const double* A, const double* B, const double* C, double* D, int n) {
  // A: [n][8][4] row-major (m,k); B: [n][4][8] (k,n); C,D: [n][8][8]
  const int lane = threadIdx.x & 31;
  for (int t = blockIdx.x; t < n; t += gridDim.x) {
    const double a = A[t * 32 + (lane >> 2) * 4 + (lane & 3)];      // A[m=lane/4][k=lane%4]
    const double b = B[t * 32 + (lane & 3) * 8 + (lane >> 2)];      // B[k=lane%4][n=lane/4]
    const int m = lane >> 2, nn = (lane & 3) * 2;
    double d0, d1;
    dmma(d0, d1, a, b, C[t * 64 + m * 8 + nn], C[t * 64 + m * 8 + nn + 1]);
    D[t * 64 + m * 8 + nn] = d0;
    D[t * 64 + m * 8 + nn + 1] = d1;
  }
}
__global__ void lat(double* out, long long* cyc, int n) {
  double a = 1.0, b = out[0] + threadIdx.x * 1e-9, c0 = 0.1, c1 = 0.2;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) dmma(c0, c1, a, b, c0, c1);
  long long t1 = clock64();
  // reduction pattern: 2 dependent DMMAs + DADD
  double p = b;
  for (int i = 0; i < n; ++i) {
    double s0, s1, u0, u1;
    dmma(s0, s1, 1.0, p, 0.0, 0.0);
    double tj = s0 + s1;
    dmma(u0, u1, 1.0, tj, 0.0, 0.0);
    p = u0 * 0.03125;
  }
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  out[1 + threadIdx.x] = c0 + c1 + p;
}
static unsigned long long s = 88172645463325252ULL;
static double rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return ((double)(s >> 11) / 9007199254740992.0 - 0.5) * std::ldexp(1.0, (int)(s % 40) - 20); }
int main() {
  const int n = 20000;
  std::vector<double> A(n * 32), B(n * 32), C(n * 64), D(n * 64);
  for (auto& v : A) v = rnd(); for (auto& v : B) v = rnd(); for (auto& v : C) v = rnd();
  double *dA, *dB, *dC, *dD; long long* dc;
  cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8); cudaMalloc(&dD, D.size() * 8); cudaMalloc(&dc, 64);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
  run<<<148, 32>>>(dA, dB, dC, dD, n); cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
  long long ok[6] = {0, 0, 0, 0, 0, 0}, tot = 0;
  for (int t = 0; t < n; ++t) for (int m = 0; m < 8; ++m) for (int j = 0; j < 8; ++j) {
    double a[4], b[4]; for (int k = 0; k < 4; ++k) { a[k] = A[t * 32 + m * 4 + k]; b[k] = B[t * 32 + k * 8 + j]; }
    const double c = C[t * 64 + m * 8 + j], d = D[t * 64 + m * 8 + j];
    double c0 = c; for (int k = 0; k < 4; ++k) c0 = std::fma(a[k], b[k], c0);           // fma chain, k ascending, c first
    double c1 = c; for (int k = 3; k >= 0; --k) c1 = std::fma(a[k], b[k], c1);          // k descending
    double c2 = 0; for (int k = 0; k < 4; ++k) c2 = std::fma(a[k], b[k], c2); c2 += c;  // products first, c last
    double c3 = c; for (int k = 0; k < 4; ++k) c3 = c3 + a[k] * b[k];                   // non-fused, ascending
    double c4 = std::fma(a[0], b[0], c) ; c4 = (c4 + std::fma(a[1], b[1], 0.0)); c4 = c4 + std::fma(a[2], b[2], std::fma(a[3], b[3], 0.0)); // pairwise-ish
    long double e = c; for (int k = 0; k < 4; ++k) e += (long double)a[k] * b[k]; double c5 = (double)e; // extended accumulate
    ok[0] += !memcmp(&c0, &d, 8); ok[1] += !memcmp(&c1, &d, 8); ok[2] += !memcmp(&c2, &d, 8); ok[3] += !memcmp(&c3, &d, 8); ok[4] += !memcmp(&c4, &d, 8); ok[5] += !memcmp(&c5, &d, 8); ++tot;
  }
  printf("elements %lld | fma-chain asc (c first) %lld | fma-chain desc %lld | products then +c %lld | unfused asc %lld | pairwise %lld | extended %lld\n", tot, ok[0], ok[1], ok[2], ok[3], ok[4], ok[5]);
  double* o; cudaMalloc(&o, 4096); double one = 1.0; cudaMemcpy(o, &one, 8, cudaMemcpyHostToDevice);
  lat<<<1, 32>>>(o, dc, 4096); long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
  printf("dependent DMMA latency %.1f cycles; reduction (2 DMMA + DADD + DMUL) %.1f cycles\n", hc[0] / 4096.0, hc[1] / 4096.0);
  for (int w = 4; w <= 32; w *= 2) { lat<<<1, 32 * w>>>(o, dc, 4096); cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost); printf("  %2d warps/SM: DMMA chain %.1f cyc/op/warp, reduction %.1f\n", w, hc[0] / 4096.0, hc[1] / 4096.0); }
  return 0;
}
