// Micro-benchmark: fp64 tensor-core mma.m8n8k4 vs DFMA throughput / latency on one SM-load pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/bench_dmma scripts/bench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS, bool MMA>
__global__ void k(double* out, long long* clk, int iters, double a, double b) {
  double c[CHAINS][2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (MMA) dmma884(c[i][0], c[i][1], a, b);
      else { c[i][0] = fma(a, b, c[i][0]); c[i][1] = fma(a, b, c[i][1]); }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int CHAINS, bool MMA>
void run(int warps, int sms) {
  double* out; long long* clk;
  cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&clk, 8);
  const int iters = 2000;
  k<CHAINS, MMA><<<sms, warps * 32>>>(out, clk, iters, 1.0000001, 0.9999999);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double per_instr = (double)h / (iters * CHAINS);
  // MMA: 256 FMA per warp instruction; DFMA: 2 x 32 FMA per loop body entry
  const double fma_per_clk_sm = (MMA ? 256.0 : 64.0) * warps / per_instr;
  printf("%s chains %2d warps/SM %2d : %7.1f clk per %s per warp, %6.1f FMA/clk/SM  (%.1f TFLOP/s at 148 SMs x 1.965 GHz)\n",
         MMA ? "DMMA" : "DFMA", CHAINS, warps, per_instr, MMA ? "mma" : "2 dfma", fma_per_clk_sm, fma_per_clk_sm * 2 * 148 * 1.965e9 / 1e12);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) { run<1, true>(w, 148); run<4, true>(w, 148); run<8, true>(w, 148); }
  for (int w : {1, 4, 8, 16, 32}) { run<1, false>(w, 148); run<8, false>(w, 148); }
  return 0;
}
