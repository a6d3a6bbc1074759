// tools/tmem_probe.cu -- can Tensor Memory serve as per-warp scratch for fp64
// vectors?  16 warps/CTA, each warp stores 10 "slots" (4 doubles per thread = 8
// columns) into its own TMEM window, reads them back, verifies, and times loads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double (&v)[4]) {
  const uint32_t r0 = __double2loint(v[0]), r1 = __double2hiint(v[0]), r2 = __double2loint(v[1]), r3 = __double2hiint(v[1]);
  const uint32_t r4 = __double2loint(v[2]), r5 = __double2hiint(v[2]), r6 = __double2loint(v[3]), r7 = __double2hiint(v[3]);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) probe(int* errors, long long* cyc, double* sink, int reps) {
  __shared__ uint32_t tmem_base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base_s;
  // this warp's window: lanes 32*(warp%4).., columns (warp/4)*80 .. +80
  const uint32_t win = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 80);
  for (int slot = 0; slot < 10; ++slot) {
    double v[4];
    for (int e = 0; e < 4; ++e) v[e] = 1000.0 * blockIdx.x + 100.0 * warp + 10.0 * slot + lane * 0.001 + e * 0.25;
    tmem_st8(win + slot * 8, v);
  }
  tmem_wait_st();
  __syncwarp();
  int bad = 0;
  for (int slot = 9; slot >= 0; --slot) {
    uint32_t r[8];
    tmem_ld8(win + slot * 8, r);
    tmem_wait_ld();
    for (int e = 0; e < 4; ++e) {
      const double got = __hiloint2double(r[2 * e + 1], r[2 * e]);
      const double want = 1000.0 * blockIdx.x + 100.0 * warp + 10.0 * slot + lane * 0.001 + e * 0.25;
      bad += (got != want);
    }
  }
  if (bad) atomicAdd(errors, bad);
  // throughput: stream all 10 slots `reps` times
  double acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < reps; ++i) {
    uint32_t r[10][8];
#pragma unroll
    for (int slot = 0; slot < 10; ++slot) tmem_ld8(win + slot * 8, r[slot]);
    tmem_wait_ld();
#pragma unroll
    for (int slot = 0; slot < 10; ++slot) acc += __hiloint2double(r[slot][1], r[slot][0]) + __hiloint2double(r[slot][7], r[slot][6]);
  }
  const long long t1 = clock64();
  // latency: dependent single loads
  uint32_t a = 0;
  for (int i = 0; i < reps; ++i) {
    uint32_t r[8];
    tmem_ld8(win + (a & 7) * 8, r);
    tmem_wait_ld();
    a = r[0] & 1;
  }
  const long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  sink[blockIdx.x * 512 + threadIdx.x] = acc + a;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}
int main() {
  int* err; long long* cyc; double* sink;
  cudaMalloc(&err, 4); cudaMemset(err, 0, 4); cudaMalloc(&cyc, 16); cudaMalloc(&sink, 148 * 512 * 8);
  const int reps = 2000;
  probe<<<148, 512>>>(err, cyc, sink, reps);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = -1; long long hc[2] = {0, 0};
  cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
  printf("status %s, mismatches %d\n", cudaGetErrorString(e), herr);
  const double bytes = 16.0 * 32 * 10 * 32 * reps;  // per SM: 16 warps x 32 lanes x 10 slots x 32 B
  printf("16 warps/SM streaming: %.1f cycles per 10-slot sweep per warp; TMEM read %.1f B/clk/SM\n", hc[0] / (double)reps, bytes / hc[0]);
  printf("dependent tcgen05.ld.x8 + wait latency: %.1f cycles (16 warps active)\n", hc[1] / (double)reps);
  return herr != 0;
}
