// Throughput of global reductions (RED) into an L2-resident array on this GPU: what bounds the flush of the Schur
// kernels (ba_schur_mma / ba_schur_pipe), which add ~87M fp64 values per launch into a 16 MB block-sparse matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bench_atomics scripts/bench_atomics.cu && scripts/bench_atomics
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int PATTERN>   // 0: every lane its own random 8-byte slot; 1: a warp hits 32 consecutive elements
__global__ void red_kernel(T* data, unsigned mask, int reps) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = gtid >> 5;
  unsigned x = (PATTERN == 0 ? gtid : warp) * 2654435761u + 12345u;
  for (int r = 0; r < reps; ++r) {
    x = x * 1664525u + 1013904223u;
    const unsigned idx = PATTERN == 0 ? ((x >> 4) & mask) : ((((x >> 4) & mask) & ~31u) + lane);
    atomicAdd(&data[idx], (T)1);
  }
}

template <typename T, int PATTERN>
void run(const char* name, int sms) {
  const unsigned n = 1u << 21;   // 2M elements (16 MB of doubles): resident in the 126 MB L2
  T* d;
  cudaMalloc(&d, n * sizeof(T));
  cudaMemset(d, 0, n * sizeof(T));
  const int blocks = sms * 8, threads = 256, reps = 256;
  red_kernel<T, PATTERN><<<blocks, threads>>>(d, n - 1, 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  red_kernel<T, PATTERN><<<blocks, threads>>>(d, n - 1, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double total = (double)blocks * threads * reps;
  printf("%-44s %8.1f G red/s  (%.3f per clock per SM at 1.965 GHz)\n", name, total / ms * 1e-6, total / ms * 1e-6 / 1.965 / sms);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<double, 0>("fp64 add, scattered lanes", sms);
  run<double, 1>("fp64 add, 32 consecutive elements per warp", sms);
  run<float, 0>("fp32 add, scattered lanes", sms);
  run<float, 1>("fp32 add, 32 consecutive elements per warp", sms);
  run<unsigned long long, 0>("u64 add, scattered lanes", sms);
  run<unsigned long long, 1>("u64 add, 32 consecutive elements per warp", sms);
  return 0;
}
