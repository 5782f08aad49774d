// tcgen05 tensor-core distance kernel for brute-force L2 matching (sm_100a).
//
// The N x M x 128 contraction of opensfm/matching.py:742-747 (cv2 knnMatch)
// is a dense GEMM: d2(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j.  For descriptors
// whose values are integers in [0,255] (HAHOG / SIFT as OpenSfM stores them,
// opensfm/features.py:526-534, and the synthetic scenes) every product and
// partial sum is an integer below 2^24, so bf16 operands with fp32 accumulation
// reproduce the float32 sum of squared differences of cv2 bit for bit.
//
// Operand layout (built once per descriptor set by prepare_tc):
//   K = 144 = 128 descriptor dims + 16 augmentation columns.
//   A role (queries):  [ a_i            | 1, 1, 1, 0 ... ]
//   B role (trains):   [ -2 b_j         | hi, mid, lo of |b_j|^2, 0 ... ]
//   => accumulator(i,j) = |b_j|^2 - 2 a_i.b_j = d2(i,j) - |a_i|^2   (exact)
//   so the epilogue needs no per-column add: ranking within a query row is
//   the ranking of the accumulator itself.
// Rows are stored in HBM already in the UMMA canonical K-major no-swizzle
// ("interleave") core-matrix order: [row/8][k/8][row%8][k%8] bf16, 128 bytes
// per core matrix, so one tile is a single contiguous range and is staged with
// one cp.async.bulk (TMA engine) per operand tile; LBO = 128 B, SBO = 18*128 B.
//
// Kernel: persistent, 1 CTA / SM, 6 warps:
//   warp 0  bulk-copy producer (Q tile double-buffered per task, T tiles 2 stages)
//   warp 1  TMEM alloc + single-thread tcgen05.mma issue (M=128, N=256, K=16 x 9)
//   warps 2-5  epilogue: tcgen05.ld the 128x256 fp32 accumulator (double-buffered
//           in TMEM, 2 x 256 columns, loads software-pipelined) and keep a running
//           top-2 per query row in d^2 space: group-of-8 min filter (0.75 instruction /
//           element) with exact updates only for elements that beat the row's current
//           second best; the loop body is kept small enough for the instruction cache.
#include <cuda_bf16.h>

#include "common.cuh"
#include "match_common.cuh"

namespace osfm {

constexpr int TC_M = 256;                // query rows per task: two M = 128 MMAs share every train tile
constexpr int TC_MH = 128;               // rows of one MMA (TMEM lanes)
constexpr int TC_N = 128;                // train rows per tile
constexpr int TC_KD = 128;               // descriptor dims carried
constexpr int TC_KP = 144;               // padded K (9 x UMMA_K)
constexpr int TC_KCH = TC_KP / 8;        // 16-byte K chunks per row
constexpr int TC_ROW_BYTES = TC_KP * 2;  // 288
constexpr int TC_Q_BYTES = TC_M * TC_ROW_BYTES;  // 36864
constexpr int TC_T_BYTES = TC_N * TC_ROW_BYTES;  // 73728
constexpr int TC_SBO = TC_KCH * 128;     // bytes between 8-row groups
constexpr int TC_LBO = 128;              // bytes between K-adjacent core matrices
constexpr int TC_STAGES = 2;
constexpr int TC_EPI_WARPS = 8;            // two per TMEM lane quarter, interleaved over 16-column chunks
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SMEM = 2 * TC_Q_BYTES + TC_STAGES * TC_T_BYTES;  // 221184

int tc_tile_m() { return TC_M; }
int tc_tile_n() { return TC_N; }

bool tc_available() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}

// ---------------------------------------------------------------------------
// Operand preparation
// ---------------------------------------------------------------------------
// One warp per padded row, one pass over the uploaded float32 matrix: exactness check (integers in
// [0,255]), |x|^2, max |x|^2, the zero-padded float32 row of the SIMT kernel (when it is not the
// upload itself) and both bf16 operand roles in the UMMA core-matrix order.
// info[0] |= 1 if any value is not bf16-exact; info[1] = max |x|^2 as float bits.
// SrcT = float (the reference's in-memory form, features.py:169-170) or uint8_t (the on-disk form of HAHOG / SIFT
// descriptors, uploaded as bytes and widened here: a quarter of the host->device traffic).
template <class SrcT>
__global__ void __launch_bounds__(256)
    tc_prepare_set(const SrcT* __restrict__ src, int n, int dim, int rows_padded, float* __restrict__ padded, int dim_padded,
                   float* __restrict__ norm, __nv_bfloat16* __restrict__ qa, __nv_bfloat16* __restrict__ tb,
                   int* __restrict__ info) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows_padded) return;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  bool bad = false;
  if (row < n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = lane * 4 + e;
      if (k < dim) {
        v[e] = (float)src[(size_t)row * dim + k];
        bad |= !(v[e] >= 0.0f && v[e] <= 255.0f && v[e] == floorf(v[e]));
      }
    }
    if (padded) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = lane * 4 + e;
        if (k < dim_padded) padded[(size_t)row * dim_padded + k] = v[e];
      }
    }
  }
  float s = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])));
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    norm[row] = s;
    if (bad) atomicOr(&info[0], 1);
    if (row < n) atomicMax(reinterpret_cast<unsigned*>(&info[1]), __float_as_uint(s));
  }
  // data chunks: lanes 2c and 2c+1 hold the two halves of 16-byte chunk c
  {
    const int c = lane >> 1;
    const size_t off = ((size_t)(row >> 3) * TC_KCH + c) * 64 + (row & 7) * 8 + (lane & 1) * 4;
    __align__(8) __nv_bfloat16 a[4], b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      a[e] = __float2bfloat16(v[e]);
      b[e] = __float2bfloat16(-2.0f * v[e]);
    }
    *reinterpret_cast<uint2*>(qa + off) = *reinterpret_cast<const uint2*>(a);
    *reinterpret_cast<uint2*>(tb + off) = *reinterpret_cast<const uint2*>(b);
  }
  // augmentation chunk (lane 0) and the zero chunk that pads K to 144 (lane 1)
  if (lane < TC_KCH - TC_KD / 8) {
    const int c = TC_KD / 8 + lane;
    const size_t off = ((size_t)(row >> 3) * TC_KCH + c) * 64 + (row & 7) * 8;
    __align__(16) __nv_bfloat16 a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      a[e] = __float2bfloat16(0.0f);
      b[e] = __float2bfloat16(0.0f);
    }
    if (lane == 0) {
      if (row < n) {
        a[0] = a[1] = a[2] = __float2bfloat16(1.0f);
        const __nv_bfloat16 hi = __float2bfloat16(s);
        const float r1 = s - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16(r1);
        const float r2 = r1 - __bfloat162float(mid);
        b[0] = hi; b[1] = mid; b[2] = __float2bfloat16(r2);
      } else {
        // padding trains can never be selected: accumulator = +inf for real queries
        b[0] = __float2bfloat16(__builtin_huge_valf());
      }
    }
    *reinterpret_cast<uint4*>(qa + off) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(tb + off) = *reinterpret_cast<const uint4*>(b);
  }
}

// src: the dense n x dim float32 upload on this stream; padded_dst: the SIMT kernel's zero-padded copy
// to fill as well, or null when the upload already is that copy.  Asynchronous: s.tc_ok is decided by
// Matcher::refresh_info() from d_info[s.slot].
void Matcher::prepare_tc(DescSet& s, const void* src, bool src_u8, float* padded_dst) {
  const int rows_padded = s.rows_padded;
  const size_t op_bytes = (size_t)rows_padded * TC_ROW_BYTES;
  __nv_bfloat16* qa = reinterpret_cast<__nv_bfloat16*>(s.tc_data);
  __nv_bfloat16* tb = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(s.tc_data) + op_bytes);
  float* norm = reinterpret_cast<float*>(reinterpret_cast<char*>(s.tc_data) + 2 * op_bytes);
  if (src_u8)
    tc_prepare_set<uint8_t><<<(rows_padded + 7) / 8, 256, 0, stream>>>(static_cast<const uint8_t*>(src), s.n, s.dim, rows_padded,
                                                                      padded_dst, s.dim_padded, norm, qa, tb, d_info.p + 2 * s.slot);
  else
    tc_prepare_set<float><<<(rows_padded + 7) / 8, 256, 0, stream>>>(static_cast<const float*>(src), s.n, s.dim, rows_padded,
                                                                    padded_dst, s.dim_padded, norm, qa, tb, d_info.p + 2 * s.slot);
  OSFM_LAUNCH_CHECK();
  s.tc_q = qa;
  s.tc_t = tb;
  s.tc_norm = norm;
  s.info_pending = true;
}
int tc_rows_padded(int n) { return (n + TC_M - 1) / TC_M * TC_M; }   // whole query tiles (and whole train tiles)
size_t tc_operand_bytes(int rows_padded) { return 2 * (size_t)rows_padded * TC_ROW_BYTES + (size_t)rows_padded * sizeof(float); }
bool tc_capable(int dim, bool u8, int n) { return !u8 && dim <= TC_KD && n > 0 && tc_available(); }

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  // try_wait suspends the thread for a hardware-defined time slice before it reports failure, so the loop is
  // cheap; the clock is only consulted every 1024 failed slices (reading it on every poll cost more issue
  // slots than the epilogue itself, profiles/r01_ncu_tc_v6.txt: 143M TRYWAIT + CS2R pairs per launch)
  for (unsigned spins = 0;; ++spins) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((spins & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {
        atomicExch(err_flag, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle (cute::UMMA::SmemDescriptor: start>>4 @0, LBO>>4 @16, SBO>>4 @32, version=1 @46)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((TC_LBO >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((TC_SBO >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format BF16 (1) @7/@10, K-major both,
// n_dim = N>>3 @17, m_dim = M>>4 @24
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                              ((uint32_t)(TC_MH >> 4) << 24);

#define OSFM_TMEM_LD16(taddr, v)                                                                          \
  asm volatile(                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                           \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                    \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
        "=r"(v[15])                                                                                       \
      : "r"(taddr)                                                                                        \
      : "memory")

#define OSFM_TMEM_LD64(taddr, v)                                                                          \
  asm volatile(                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                           \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                           \
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"                                  \
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"                                  \
      "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"                          \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]),        \
        "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]),        \
        "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),        \
        "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),        \
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])         \
      : "r"(taddr)                                                                                        \
      : "memory")

struct TcTask {
  MatchJob job;
  int q0, t_begin, ntiles, chunk;
};

__device__ __forceinline__ TcTask tc_decode(const MatchJob* jobs, const int* tile_prefix, int njobs, int task) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_prefix[mid] <= task) lo = mid; else hi = mid - 1;
  }
  TcTask t;
  t.job = jobs[lo];
  const int local = task - tile_prefix[lo];
  const int qtile = local / t.job.nchunks;
  t.chunk = local % t.job.nchunks;
  t.q0 = qtile * TC_M;
  t.t_begin = t.chunk * t.job.chunk_len;
  const int t_end = min(t.job.nt, t.t_begin + t.job.chunk_len);
  t.ntiles = (t_end - t.t_begin + TC_N - 1) / TC_N;
  return t;
}

// Epilogue state of one query row: the two smallest accumulator values (= d^2 - |a|^2, exact
// integers) with their train indices.  Ranking in d^2 is the ranking cv2 uses (sqrt'd float32
// distance, ties -> lowest index) as long as float32 sqrt is injective on the integers involved,
// i.e. d^2 < 2^22; the host only selects this kernel for descriptor sets whose norms guarantee
// that bound (Matcher::match_pairs_async), everything else goes to the exact SIMT kernel.
struct RowState {
  float q1, q2;
  int i1, i2;
};

// Branch-free update (used for the first tile of a task, where most elements are records).
__device__ __forceinline__ void row_update(RowState& st, float x, int idx) {
  const bool lt1 = x < st.q1;
  const bool lt2 = x < st.q2;
  st.q2 = lt1 ? st.q1 : (lt2 ? x : st.q2);
  st.i2 = lt1 ? st.i1 : (lt2 ? idx : st.i2);
  st.q1 = lt1 ? x : st.q1;
  st.i1 = lt1 ? idx : st.i1;
}

// 16 accumulator columns of one row: two groups of 8, min filter against the row's current second
// best, exact (predicated) updates only inside a group that beats it.  Deliberately small: the whole
// epilogue loop body must stay resident in the instruction cache (an earlier fully unrolled version was
// 89 KB of SASS and spent most of its time in instruction-fetch stalls, profiles/r01_*).
template <int OFF, int NV, bool MASKED>
__device__ __forceinline__ void row_consume16(RowState& st, const uint32_t (&vv)[NV], int col0, uint32_t bits16) {
  const uint32_t* v = vv + OFF;   // OFF is a compile-time constant: the accesses below stay register-resident
  // minima of four groups of 4 (independent chains), one test for the common case; a triggered chunk
  // re-examines only the group(s) of 4 that beat the threshold (their updates are predicated by ptxas,
  // 7 instructions per element, so small groups matter).  MASKED (guided matching): elements whose bit is
  // clear read as +inf and can never enter the row's two best.
  float x[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    x[e] = __uint_as_float(v[e]);
    if (MASKED) x[e] = ((bits16 >> e) & 1u) ? x[e] : __builtin_huge_valf();
  }
  if (MASKED && bits16 == 0u) return;
  float m[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) m[g] = fminf(fminf(fminf(x[g * 4], x[g * 4 + 1]), x[g * 4 + 2]), x[g * 4 + 3]);
  if (fminf(fminf(fminf(m[0], m[1]), m[2]), m[3]) < st.q2) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (m[g] < st.q2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (x[g * 4 + e] < st.q2) row_update(st, x[g * 4 + e], col0 + g * 4 + e);
        }
      }
    }
  }
}

template <bool MASKED>
__global__ void __launch_bounds__(TC_THREADS, 1)
    bf_top2_tc(const MatchJob* __restrict__ jobs, const int* __restrict__ tile_prefix, int njobs, int ntasks,
               Top2* __restrict__ partial, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_qfull[2], bar_qempty[2], bar_full[TC_STAGES], bar_empty[TC_STAGES],
      bar_accfull[2], bar_accempty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* q_smem[2] = {smem, smem + TC_Q_BYTES};
  uint8_t* t_smem[TC_STAGES];
#pragma unroll
  for (int s = 0; s < TC_STAGES; ++s) t_smem[s] = smem + 2 * TC_Q_BYTES + s * TC_T_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_qfull[i], 1);
      mbar_init(&bar_qempty[i], 1);
      mbar_init(&bar_accfull[i], 1);
      mbar_init(&bar_accempty[i], TC_EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_smem))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===== bulk-copy producer =====
    if (lane == 0) {
      int stage = 0, ph = 0, n = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++n) {
        const TcTask t = tc_decode(jobs, tile_prefix, njobs, task);
        const int b = n & 1, qph = (n >> 1) & 1;
        mbar_wait(&bar_qempty[b], qph ^ 1, err_flag);
        // the last query tile of a job may hold <= 128 rows: only its first half is loaded and multiplied
        const uint32_t qbytes = (t.job.nq - t.q0 > TC_MH) ? TC_Q_BYTES : TC_Q_BYTES / 2;
        mbar_expect_tx(&bar_qfull[b], qbytes);
        bulk_copy_g2s(q_smem[b], reinterpret_cast<const uint8_t*>(t.job.q_tc) + (size_t)t.q0 * TC_ROW_BYTES,
                      qbytes, &bar_qfull[b]);
        for (int i = 0; i < t.ntiles; ++i) {
          mbar_wait(&bar_empty[stage], ph ^ 1, err_flag);
          mbar_expect_tx(&bar_full[stage], TC_T_BYTES);
          bulk_copy_g2s(t_smem[stage],
                        reinterpret_cast<const uint8_t*>(t.job.t_tc) + (size_t)(t.t_begin + i * TC_N) * TC_ROW_BYTES,
                        TC_T_BYTES, &bar_full[stage]);
          if (++stage == TC_STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0, ph = 0, n = 0, tilecount = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++n) {
        const TcTask t = tc_decode(jobs, tile_prefix, njobs, task);
        const int b = n & 1, qph = (n >> 1) & 1;
        mbar_wait(&bar_qfull[b], qph, err_flag);
        const uint64_t adesc0 = make_smem_desc(smem_u32(q_smem[b]));
        const uint64_t adesc1 = make_smem_desc(smem_u32(q_smem[b]) + (TC_MH / 8) * TC_SBO);   // query rows 128..255
        const bool two = t.job.nq - t.q0 > TC_MH;   // a short last tile skips the second MMA (its rows do not exist)
        for (int i = 0; i < t.ntiles; ++i, ++tilecount) {
          const int a = tilecount & 1, aph = (tilecount >> 1) & 1;
          mbar_wait(&bar_accempty[a], aph ^ 1, err_flag);
          mbar_wait(&bar_full[stage], ph, err_flag);
          tc_fence_after();
          const uint64_t bdesc0 = make_smem_desc(smem_u32(t_smem[stage]));
          // accumulator stage a = TMEM columns [256 a, 256 a + 256): query rows 0..127 in the first 128 columns,
          // 128..255 in the second.  Both MMAs read the same train tile from shared memory: every byte the TMA
          // engine brings in feeds 256 query rows (the M = 128 x N = 256 tile moved twice the bytes per output
          // and was bound by the L2 -> SM path, profiles/README.md round 2)
          const uint32_t d_tmem = tmem_base + (uint32_t)a * (2 * TC_N);
#pragma unroll
          for (int k = 0; k < TC_KP / 16; ++k) {
            // one UMMA_K = 16 bf16 = two core matrices = 256 bytes along K
            const uint64_t koff = (uint64_t)((k * 2 * TC_LBO) >> 4);
            tc_mma_bf16(d_tmem, adesc0 + koff, bdesc0 + koff, TC_IDESC, k > 0 ? 1u : 0u);
            if (two) tc_mma_bf16(d_tmem + TC_N, adesc1 + koff, bdesc0 + koff, TC_IDESC, k > 0 ? 1u : 0u);
          }
          tc_commit(&bar_empty[stage]);   // smem stage reusable when these MMAs retire
          tc_commit(&bar_accfull[a]);     // accumulator complete
          if (++stage == TC_STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(&bar_qempty[b]);  // Q buffer reusable after the task's last MMA
      }
    }
  } else {
    // ===== epilogue: warps 2..9; a warp may only touch TMEM lanes 32*(warp%4)..+31.  The two warps of a lane
    // quarter take the two query halves of the task (rows 0..127 / 128..255 = the two accumulators of a stage), so
    // each SM sub-partition always has two independent instruction streams and every thread owns one query row =====
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = half * TC_MH + quarter * 32 + lane;
    int tilecount = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const TcTask t = tc_decode(jobs, tile_prefix, njobs, task);
      const int gq = t.q0 + row_in_tile;
      const float na = gq < t.job.nq ? t.job.q_norm[gq] : 0.0f;
      RowState st;
      st.q1 = st.q2 = __builtin_huge_valf();
      st.i1 = st.i2 = -1;
      for (int i = 0; i < t.ntiles; ++i, ++tilecount) {
        const int a = tilecount & 1, aph = (tilecount >> 1) & 1;
        mbar_wait(&bar_accfull[a], aph, err_flag);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * (2 * TC_N) + (uint32_t)half * TC_N;
        const int col_base = t.t_begin + i * TC_N;
        // this row's 128 columns of the tile in two 64-column loads, both in flight at once; the accumulator is
        // handed back to the MMA issuer as soon as the values are in registers -- the TMEM stage is held for one
        // load latency, not for the consume time (8 dependent 16-column loads per tile made the epilogue the
        // critical path: profiles/r01_ncu_tc_v6.txt, top stall on the accumulator-full wait)
        uint32_t va[64], vb[64];
        uint32_t mw[4] = {0u, 0u, 0u, 0u};   // guided matching: the row's mask bits of these 128 columns
        if (MASKED && gq < t.job.nq) {
          const uint32_t* mrow = t.job.mask_bits + (size_t)gq * t.job.mask_words;
          const int w0 = col_base >> 5;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (w0 + q < t.job.mask_words) mw[q] = mrow[w0 + q];
        }
        OSFM_TMEM_LD64(taddr, va);
        OSFM_TMEM_LD64(taddr + 64, vb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_accempty[a]);
        row_consume16<0, 64, MASKED>(st, va, col_base, mw[0] & 0xffffu);
        row_consume16<16, 64, MASKED>(st, va, col_base + 16, mw[0] >> 16);
        row_consume16<32, 64, MASKED>(st, va, col_base + 32, mw[1] & 0xffffu);
        row_consume16<48, 64, MASKED>(st, va, col_base + 48, mw[1] >> 16);
        row_consume16<0, 64, MASKED>(st, vb, col_base + 64, mw[2] & 0xffffu);
        row_consume16<16, 64, MASKED>(st, vb, col_base + 80, mw[2] >> 16);
        row_consume16<32, 64, MASKED>(st, vb, col_base + 96, mw[3] & 0xffffu);
        row_consume16<48, 64, MASKED>(st, vb, col_base + 112, mw[3] >> 16);
      }
      if (gq < t.job.nq) {
        // partial results of this kernel are squared distances (exact integers in fp32)
        Top2 out;
        out.s1 = st.i1 >= 0 ? fmaxf(st.q1 + na, 0.0f) : __builtin_huge_valf();
        out.i1 = st.i1;
        out.s2 = st.i2 >= 0 ? fmaxf(st.q2 + na, 0.0f) : __builtin_huge_valf();
        out.i2 = st.i2;
        partial[t.job.partial_off + (size_t)t.chunk * t.job.nq + gq] = out;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ===========================================================================================================
// Hamming distance on the tensor cores ("BruteForce-Hamming": AKAZE MLDB 61 bytes, ORB 32 bytes).
//
// Bits become +-1 in fp8 (E4M3: +1 = 0x38, -1 = 0xB8, both exact): for two descriptors a, b of nbits bits
//     sum_i a_i b_i = (#equal bits) - (#different bits) = nbits - 2 H(a, b),
// so with the train operand negated the fp32 accumulator is 2 H - nbits, an exact integer, and its ranking inside
// a query row is cv2's ranking by Hamming distance (ties -> lowest index through the same epilogue as the L2
// kernel).  K = 512 fp8 per row; positions beyond nbits are 0 in real rows of both roles (they add nothing),
// +1 in every query row and +448 in the *padding* rows of a train set, so a padding train scores
// >= 8 * 448 - nbits > any real one and is never selected (needs >= 8 spare positions: nbytes <= 63).
// `tcgen05.mma.kind::f8f6f4`, M = 128, N = 128, K = 32 x 16; operands in the same no-swizzle K-major core-matrix
// order as the bf16 kernel (a core matrix row is 16 bytes = 16 fp8), 512 B per row.  1 CTA / SM, 10 warps:
// bulk-copy producer (Q tile single-buffered: 64 KB, T tiles 2 x 64 KB), MMA issuer (2 accumulator stages of 128
// TMEM columns), 8 epilogue warps = 4 lane quarters x 2 column halves (one 64-column TMEM load per warp and tile).
// The popcount kernel this replaces is bound by the POPC pipe (16 / clk / SM): 2.1e11 pairs/s at 74 % of that roof.
// ===========================================================================================================
constexpr int H8_M = 128, H8_N = 128;
constexpr int H8_ROW_BYTES = 512;
constexpr int H8_KCH = H8_ROW_BYTES / 16;         // 16-byte K chunks per row
constexpr int H8_SBO = H8_KCH * 128;              // bytes between 8-row groups
constexpr int H8_Q_BYTES = H8_M * H8_ROW_BYTES;   // 65536
constexpr int H8_T_BYTES = H8_N * H8_ROW_BYTES;   // 65536
constexpr int H8_STAGES = 2;
constexpr int H8_SMEM = H8_Q_BYTES + H8_STAGES * H8_T_BYTES;   // 196608
// c_format F32 (1) @4, a/b format E4M3 (0) @7/@10, K-major, n_dim = N>>3 @17, m_dim = M>>4 @24
constexpr uint32_t H8_IDESC = (1u << 4) | ((uint32_t)(H8_N >> 3) << 17) | ((uint32_t)(H8_M >> 4) << 24);

int h8_tile_m() { return H8_M; }
int h8_tile_n() { return H8_N; }
size_t h8_operand_bytes(int rows_padded) { return 2 * (size_t)rows_padded * H8_ROW_BYTES; }
bool h8_capable(int nbytes, int n) { return nbytes >= 1 && nbytes <= 63 && n > 0 && tc_available(); }

// one thread per (row, 16-byte K chunk): 16 bits of the source row -> 16 fp8 values in both roles
__global__ void __launch_bounds__(256)
    h8_prepare_set(const uint8_t* __restrict__ src, int n, int nbytes, int src_stride, int rows_padded, uint8_t* __restrict__ qa,
                   uint8_t* __restrict__ tb) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)rows_padded * H8_KCH) return;
  const int row = (int)(idx / H8_KCH), c = (int)(idx % H8_KCH);
  const int nbits = nbytes * 8;
  __align__(16) uint8_t a[16], b[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int bit = c * 16 + e;
    uint8_t va = 0, vb = 0;
    if (row < n) {
      if (bit < nbits) {
        const int on = (src[(size_t)row * src_stride + (bit >> 3)] >> (bit & 7)) & 1;
        va = on ? 0x38 : 0xB8;   // +1 / -1
        vb = on ? 0xB8 : 0x38;   // negated
      } else {
        va = 0x38;               // +1 against the padding trains' markers
      }
    } else if (bit >= nbits) {
      vb = 0x7E;                 // +448: a padding train can never win
    }
    a[e] = va; b[e] = vb;
  }
  const size_t off = ((size_t)(row >> 3) * H8_KCH + c) * 128 + (row & 7) * 16;
  *reinterpret_cast<uint4*>(qa + off) = *reinterpret_cast<const uint4*>(a);
  *reinterpret_cast<uint4*>(tb + off) = *reinterpret_cast<const uint4*>(b);
}

void Matcher::prepare_h8(DescSet& s, const uint8_t* src, int src_stride) {
  const size_t op_bytes = (size_t)s.rows_padded * H8_ROW_BYTES;
  uint8_t* qa = reinterpret_cast<uint8_t*>(s.tc_data);
  uint8_t* tb = qa + op_bytes;
  const size_t total = (size_t)s.rows_padded * H8_KCH;
  h8_prepare_set<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, s.n, s.dim, src_stride, s.rows_padded, qa, tb);
  OSFM_LAUNCH_CHECK();
  s.tc_q = reinterpret_cast<const __nv_bfloat16*>(qa);
  s.tc_t = reinterpret_cast<const __nv_bfloat16*>(tb);
  s.tc_norm = nullptr;
  s.tc_ok = true;
}

__device__ __forceinline__ void tc_mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t h8_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((TC_LBO >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((H8_SBO >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

struct H8Task {
  MatchJob job;
  int q0, t_begin, ntiles, chunk;
};
__device__ __forceinline__ H8Task h8_decode(const MatchJob* jobs, const int* tile_prefix, int njobs, int task) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_prefix[mid] <= task) lo = mid; else hi = mid - 1;
  }
  H8Task t;
  t.job = jobs[lo];
  const int local = task - tile_prefix[lo];
  const int qtile = local / t.job.nchunks;
  t.chunk = local % t.job.nchunks;
  t.q0 = qtile * H8_M;
  t.t_begin = t.chunk * t.job.chunk_len;
  const int t_end = min(t.job.nt, t.t_begin + t.job.chunk_len);
  t.ntiles = (t_end - t.t_begin + H8_N - 1) / H8_N;
  return t;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    bf_top2_tc_h8(const MatchJob* __restrict__ jobs, const int* __restrict__ tile_prefix, int njobs, int ntasks,
                  Top2* __restrict__ partial, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_qfull, bar_qempty, bar_full[H8_STAGES], bar_empty[H8_STAGES], bar_accfull[2],
      bar_accempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float4 merge_buf[H8_M];
  uint8_t* q_smem = smem;
  uint8_t* t_smem[H8_STAGES];
#pragma unroll
  for (int st = 0; st < H8_STAGES; ++st) t_smem[st] = smem + H8_Q_BYTES + st * H8_T_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_qfull, 1);
    mbar_init(&bar_qempty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_accfull[i], 1);
      mbar_init(&bar_accempty[i], TC_EPI_WARPS);
    }
    for (int i = 0; i < H8_STAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_smem))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, ph = 0, n = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++n) {
        const H8Task t = h8_decode(jobs, tile_prefix, njobs, task);
        mbar_wait(&bar_qempty, (n & 1) ^ 1, err_flag);
        mbar_expect_tx(&bar_qfull, H8_Q_BYTES);
        bulk_copy_g2s(q_smem, reinterpret_cast<const uint8_t*>(t.job.q_tc) + (size_t)t.q0 * H8_ROW_BYTES, H8_Q_BYTES, &bar_qfull);
        for (int i = 0; i < t.ntiles; ++i) {
          mbar_wait(&bar_empty[stage], ph ^ 1, err_flag);
          mbar_expect_tx(&bar_full[stage], H8_T_BYTES);
          bulk_copy_g2s(t_smem[stage],
                        reinterpret_cast<const uint8_t*>(t.job.t_tc) + (size_t)(t.t_begin + i * H8_N) * H8_ROW_BYTES,
                        H8_T_BYTES, &bar_full[stage]);
          if (++stage == H8_STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0, ph = 0, n = 0, tilecount = 0;
      for (int task = blockIdx.x; task < ntasks; task += gridDim.x, ++n) {
        const H8Task t = h8_decode(jobs, tile_prefix, njobs, task);
        mbar_wait(&bar_qfull, n & 1, err_flag);
        const uint64_t adesc0 = h8_smem_desc(smem_u32(q_smem));
        for (int i = 0; i < t.ntiles; ++i, ++tilecount) {
          const int a = tilecount & 1, aph = (tilecount >> 1) & 1;
          mbar_wait(&bar_accempty[a], aph ^ 1, err_flag);
          mbar_wait(&bar_full[stage], ph, err_flag);
          tc_fence_after();
          const uint64_t bdesc0 = h8_smem_desc(smem_u32(t_smem[stage]));
          const uint32_t d_tmem = tmem_base + (uint32_t)a * H8_N;
#pragma unroll
          for (int k = 0; k < H8_ROW_BYTES / 32; ++k) {
            // one UMMA_K = 32 fp8 = two core matrices = 256 bytes along K
            const uint64_t koff = (uint64_t)((k * 2 * TC_LBO) >> 4);
            tc_mma_f8(d_tmem, adesc0 + koff, bdesc0 + koff, H8_IDESC, k > 0 ? 1u : 0u);
          }
          tc_commit(&bar_empty[stage]);
          tc_commit(&bar_accfull[a]);
          if (++stage == H8_STAGES) { stage = 0; ph ^= 1; }
        }
        tc_commit(&bar_qempty);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;          // column half of the 128-wide tile
    const int row_in_tile = quarter * 32 + lane;
    int tilecount = 0;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      const H8Task t = h8_decode(jobs, tile_prefix, njobs, task);
      const int gq = t.q0 + row_in_tile;
      const float nbits = (float)(t.job.dim * 8);
      RowState st;
      st.q1 = st.q2 = __builtin_huge_valf();
      st.i1 = st.i2 = -1;
      for (int i = 0; i < t.ntiles; ++i, ++tilecount) {
        const int a = tilecount & 1, aph = (tilecount >> 1) & 1;
        mbar_wait(&bar_accfull[a], aph, err_flag);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)a * H8_N + (uint32_t)half * 64;
        const int col_base = t.t_begin + i * H8_N + half * 64;
        uint32_t va[64];
        OSFM_TMEM_LD64(taddr, va);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_accempty[a]);
        row_consume16<0, 64, false>(st, va, col_base, 0u);
        row_consume16<16, 64, false>(st, va, col_base + 16, 0u);
        row_consume16<32, 64, false>(st, va, col_base + 32, 0u);
        row_consume16<48, 64, false>(st, va, col_base + 48, 0u);
      }
      // merge the two column halves of a row (lexicographic (distance, index), like cv2's insertion order)
      if (half == 1) merge_buf[row_in_tile] = make_float4(st.q1, __int_as_float(st.i1), st.q2, __int_as_float(st.i2));
      asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");
      if (half == 0) {
        const float4 o = merge_buf[row_in_tile];
        Top2 a2, b2;
        a2.s1 = st.q1; a2.i1 = st.i1; a2.s2 = st.q2; a2.i2 = st.i2;
        b2.s1 = o.x; b2.i1 = __float_as_int(o.y); b2.s2 = o.z; b2.i2 = __float_as_int(o.w);
        top2_merge(a2, b2);
        if (gq < t.job.nq) {
          Top2 out;   // accumulator = 2 H - nbits  ->  the Hamming distance cv2 reports (an integer as float)
          out.s1 = a2.i1 >= 0 ? (a2.s1 + nbits) * 0.5f : __builtin_huge_valf();
          out.i1 = a2.i1;
          out.s2 = a2.i2 >= 0 ? (a2.s2 + nbits) * 0.5f : __builtin_huge_valf();
          out.i2 = a2.i2;
          partial[t.job.partial_off + (size_t)t.chunk * t.job.nq + gq] = out;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(32 * TC_EPI_WARPS) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

void launch_tc_h8(Matcher& m, int njobs, int ntasks) {
  if (!m.h8_attr_set) {
    OSFM_CUDA(cudaFuncSetAttribute(bf_top2_tc_h8, cudaFuncAttributeMaxDynamicSharedMemorySize, H8_SMEM));
    m.h8_attr_set = true;
  }
  m.d_flags.reserve(4);
  OSFM_CUDA(cudaMemsetAsync(m.d_flags.p + 1, 0, sizeof(int), m.stream));
  const int grid = std::min(ntasks, m.num_sms);
  bf_top2_tc_h8<<<grid, TC_THREADS, H8_SMEM, m.stream>>>(m.d_jobs.p, m.d_prefix.p, njobs, ntasks, m.d_partial.p, m.d_flags.p + 1);
  OSFM_LAUNCH_CHECK();
}

void launch_tc(Matcher& m, int njobs, int ntasks, bool masked) {
  if (!m.tc_attr_set) {   // per matcher (= per device)
    OSFM_CUDA(cudaFuncSetAttribute(bf_top2_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    OSFM_CUDA(cudaFuncSetAttribute(bf_top2_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    m.tc_attr_set = true;
  }
  m.d_flags.reserve(4);
  OSFM_CUDA(cudaMemsetAsync(m.d_flags.p + 1, 0, sizeof(int), m.stream));
  const int grid = std::min(ntasks, m.num_sms);
  if (masked)
    bf_top2_tc<true><<<grid, TC_THREADS, TC_SMEM, m.stream>>>(m.d_jobs.p, m.d_prefix.p, njobs, ntasks, m.d_partial.p,
                                                             m.d_flags.p + 1);
  else
    bf_top2_tc<false><<<grid, TC_THREADS, TC_SMEM, m.stream>>>(m.d_jobs.p, m.d_prefix.p, njobs, ntasks, m.d_partial.p,
                                                              m.d_flags.p + 1);
  OSFM_LAUNCH_CHECK();
}

}  // namespace osfm
