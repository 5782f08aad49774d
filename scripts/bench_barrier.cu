// Micro-benchmark: grid barrier (+ 2-double all-reduce) variants for a persistent 1-CTA/SM kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/../scripts/bench_barrier scripts/bench_barrier.cu
// Prints clocks per barrier (CTA 0, thread 0) for each variant.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int MAXC = 256;
struct State {
  unsigned flags[MAXC * 32];  // stride selectable
  double slot[2][MAXC][2];
  unsigned counter;
  unsigned release;
  long long clocks;
  double sink;
};

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_rlx(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// V0: all-to-all flags, acquire polls by threads t < nblocks, values read after.
template <int STRIDE, bool ACQ_POLL>
__device__ __forceinline__ void bar_all2all(State* st, unsigned nb, unsigned& gen, double a, double& A, double* red) {
  ++gen;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(~0u, a, o);
  if (lane == 0) red[warp] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    __stcg(&st->slot[gen & 1][blockIdx.x][0], s);
    st_rel(&st->flags[blockIdx.x * STRIDE], gen);
  }
  double v = 0;
  if (threadIdx.x < nb) {
    if (ACQ_POLL) {
      while (ld_acq(&st->flags[threadIdx.x * STRIDE]) != gen) {}
    } else {
      while (ld_rlx(&st->flags[threadIdx.x * STRIDE]) != gen) {}
      __threadfence();
    }
    v = __ldcg(&st->slot[gen & 1][threadIdx.x][0]);
  }
  __syncthreads();
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < (int)((nb + 31) >> 5); ++w) s += red[w];
  A = s;
  __syncthreads();
}

// V1: flag carries the payload: 64-bit word = (gen << 32 | float bits)?  Not exact for doubles; instead
// publish value first and flag second but poll the flag with relaxed loads by ONE warp per 32 CTAs.
// V2: classic counter: atom.add.release arrive, thread 0 polls the counter; values via atomicAdd double.
__device__ __forceinline__ void bar_counter(State* st, unsigned nb, unsigned& gen, double a, double& A, double* red,
                                            double* accum) {
  ++gen;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(~0u, a, o);
  if (lane == 0) red[warp] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    atomicAdd(&accum[gen & 3], s);
    if (blockIdx.x == 0) accum[(gen + 2) & 3] = 0.0;
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(&st->counter) : "memory");
    while (ld_acq(&st->counter) < gen * nb) {}
  }
  __syncthreads();
  A = __ldcg(&accum[gen & 3]);
}

// V3: two-level: CTA 0 gathers (threads poll one flag each), publishes totals + release word; others poll it.
template <int STRIDE>
__device__ __forceinline__ void bar_gather(State* st, unsigned nb, unsigned& gen, double a, double& A, double* red) {
  ++gen;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(~0u, a, o);
  if (lane == 0) red[warp] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    __stcg(&st->slot[gen & 1][blockIdx.x][0], s);
    st_rel(&st->flags[blockIdx.x * STRIDE], gen);
  }
  if (blockIdx.x == 0) {
    double v = 0;
    if (threadIdx.x < nb) {
      while (ld_acq(&st->flags[threadIdx.x * STRIDE]) != gen) {}
      v = __ldcg(&st->slot[gen & 1][threadIdx.x][0]);
    }
    __syncthreads();
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(~0u, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
      for (int w = 0; w < (int)((nb + 31) >> 5); ++w) s += red[w];
      __stcg(&st->slot[gen & 1][MAXC - 1][1], s);
      st_rel(&st->release, gen);
    }
  }
  if (threadIdx.x == 0) {
    while (ld_acq(&st->release) != gen) {}
  }
  __syncthreads();
  A = __ldcg(&st->slot[gen & 1][MAXC - 1][1]);
  __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(512, 1) bench(State* st, double* accum, int iters, int work) {
  __shared__ double red[16];
  unsigned gen = 0;
  double x = threadIdx.x * 1e-3, A = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int w = 0; w < work; ++w) x = x * 1.0000001 + 1e-9;  // stand-in for the phase's arithmetic
    if (V == 0) bar_all2all<1, true>(st, gridDim.x, gen, x, A, red);
    if (V == 1) bar_all2all<32, true>(st, gridDim.x, gen, x, A, red);
    if (V == 2) bar_all2all<1, false>(st, gridDim.x, gen, x, A, red);
    if (V == 3) bar_counter(st, gridDim.x, gen, x, A, red, accum);
    if (V == 4) bar_gather<1>(st, gridDim.x, gen, x, A, red);
    if (V == 5) bar_all2all<32, false>(st, gridDim.x, gen, x, A, red);
    x += A * 1e-30;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { st->clocks = clock64() - t0; st->sink = x; }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  State* st;
  double* accum;
  cudaMalloc(&st, sizeof(State));
  cudaMalloc(&accum, 4 * sizeof(double));
  const int iters = 2000;
  const char* names[] = {"all2all packed acquire-poll", "all2all 128B-stride acquire-poll", "all2all packed relaxed-poll+fence",
                         "atomic counter + atomicAdd", "gather at CTA0 + release word", "all2all 128B-stride relaxed-poll+fence"};
  for (int threads : {512, 256}) {
    for (int v = 0; v < 6; ++v) {
      for (int work : {0, 2000}) {
        cudaMemset(st, 0, sizeof(State));
        cudaMemset(accum, 0, 4 * sizeof(double));
        switch (v) {
          case 0: bench<0><<<sms, threads>>>(st, accum, iters, work); break;
          case 1: bench<1><<<sms, threads>>>(st, accum, iters, work); break;
          case 2: bench<2><<<sms, threads>>>(st, accum, iters, work); break;
          case 3: bench<3><<<sms, threads>>>(st, accum, iters, work); break;
          case 4: bench<4><<<sms, threads>>>(st, accum, iters, work); break;
          case 5: bench<5><<<sms, threads>>>(st, accum, iters, work); break;
        }
        cudaError_t e = cudaDeviceSynchronize();
        State h;
        cudaMemcpy(&h, st, sizeof(State), cudaMemcpyDeviceToHost);
        printf("threads %3d  %-42s work %4d : %8.0f clk / iteration  (%s)\n", threads, names[v], work,
               (double)h.clocks / iters, cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
