// Types shared by the SIMT and tcgen05 matching kernels.
#pragma once
#include <cuda_bf16.h>

#include <map>
#include <vector>

#include "common.cuh"

namespace osfm {

// Running two-nearest list of one query, in cv2's ranking space:
// s = sqrt(float32 d^2) for L2, bit count for Hamming; i = train index or -1.
struct Top2 {
  float s1;
  int i1;
  float s2;
  int i2;
};

__host__ __device__ inline Top2 top2_empty() {
  Top2 t;
  t.s1 = __builtin_huge_valf();
  t.s2 = __builtin_huge_valf();
  t.i1 = -1;
  t.i2 = -1;
  return t;
}

// Insert candidate (s, i); order is lexicographic (distance, index), which is
// what cv2's stable insertion with a strict `<` produces when trains are
// visited in increasing index order.
__host__ __device__ inline void top2_insert(Top2& t, float s, int i) {
  if (s < t.s1 || (s == t.s1 && (unsigned)i < (unsigned)t.i1)) {
    t.s2 = t.s1; t.i2 = t.i1;
    t.s1 = s; t.i1 = i;
  } else if (s < t.s2 || (s == t.s2 && (unsigned)i < (unsigned)t.i2)) {
    t.s2 = s; t.i2 = i;
  }
}

__host__ __device__ inline void top2_merge(Top2& t, const Top2& o) {
  if (o.i1 >= 0) top2_insert(t, o.s1, o.i1);
  if (o.i2 >= 0) top2_insert(t, o.s2, o.i2);
}

// One direction of one image pair.
struct MatchJob {
  const void* q;  // queries, padded rows (float32 or packed uint8 words)
  const void* t;  // trains
  const __nv_bfloat16* q_tc;  // tcgen05 operand copies (blocked core-matrix layout), or null
  const __nv_bfloat16* t_tc;
  const float* q_norm;        // |q_i|^2 (tcgen05 path)
  const float* t_norm;
  int nq, nt;
  int dim, dim_padded;  // dim_padded: 4-byte elements per padded row
  int qtiles;
  int nchunks, chunk_len;
  long long partial_off;  // Top2 partial[partial_off + chunk * nq + q]
  long long match_off;    // int32 match[match_off + q]
  const uint8_t* mask;
  long long mask_sq, mask_st;
  // guided matching: one bit per (query, train), row-major over the queries, mask_words 32-bit words per row
  const uint32_t* mask_bits;
  int mask_words;
};
__device__ __forceinline__ bool job_allows(const MatchJob& job, int gq, int gt) {
  if (job.mask && job.mask[(size_t)gq * job.mask_sq + (size_t)gt * job.mask_st] == 0) return false;
  if (job.mask_bits && !((job.mask_bits[(size_t)gq * job.mask_words + (gt >> 5)] >> (gt & 31)) & 1u)) return false;
  return true;
}

struct DescSet {
  void* data = nullptr;
  int n = 0, dim = 0, dim_padded = 0, row_bytes = 0;
  bool u8 = false;
  // tcgen05 operands (float32 sets with integer values in [0,255] and dim <= 128 only)
  void* tc_data = nullptr;  // one allocation: [A-role | B-role | norms]
  const __nv_bfloat16* tc_q = nullptr;
  const __nv_bfloat16* tc_t = nullptr;
  const float* tc_norm = nullptr;
  bool tc_ok = false;
  float tc_max_norm = 0.0f;  // max_i |x_i|^2
  int rows_padded = 0;
  // device memory comes from the matcher's slabs; the exactness flag / max norm of a freshly added set
  // live in d_info[slot] until refresh_info() reads them back (no host sync per add)
  int slab = -1, slot = -1;
  size_t slab_bytes = 0, bear_bytes = 0;   // sizes of the two slab allocations (data + operands; bearings)
  bool info_pending = false;
  // unit bearing vectors of the features (n x 3 float32), for guided matching; null until set
  float* bearings = nullptr;
  int bear_slab = -1;
};

struct Slab {
  char* base = nullptr;
  size_t cap = 0, used = 0;   // bump pointer
  int live = 0;
  // ranges below `used` that were released while neighbours stayed live: (offset, bytes), sorted, coalesced.
  // A long-lived matcher that replaces descriptor sets key by key reuses them instead of growing.
  std::vector<std::pair<size_t, size_t>> free_ranges;
};

struct Matcher {
  int device;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4];
  int num_sms = 148;
  int next_id = 1;
  int kernel_choice = 0;
  int last_kernel = 0;
  long long last_total_results = 0;
  int last_npairs = 0;
  bool results_in_match_buf = true;
  bool tc_attr_set = false, tcm_attr_set = false, fx_attr_set = false, h8_attr_set = false;
  std::map<int, DescSet> sets;
  std::vector<MatchJob> h_jobs;
  std::vector<int> h_prefix;
  std::vector<long long> h_out_off;
  DevBuf<MatchJob> d_jobs;
  DevBuf<int> d_prefix;
  DevBuf<long long> d_out_off;
  DevBuf<Top2> d_partial;
  DevBuf<int32_t> d_match, d_out;
  DevBuf<uint8_t> staging, mask_buf;
  DevBuf<int> d_flags;
  DevBuf<uint32_t> d_mask_bits;
  DevBuf<int> d_pair_counts;
  DevBuf<long long> d_pair_off;
  DevBuf<int32_t> d_pairs;
  DevBuf<double> d_epi_vec, d_epi_pose;
  PinnedBuf<double> p_epi_pose;
  PinnedBuf<MatchJob> p_jobs;
  PinnedBuf<int> p_prefix;
  PinnedBuf<long long> p_out_off;

  explicit Matcher(int dev);
  ~Matcher();
  Matcher(const Matcher&) = delete;
  Matcher& operator=(const Matcher&) = delete;

  std::vector<Slab> slabs;
  DevBuf<int> d_info;                 // [MAX_SLOTS][2]: not-exact flag, max |x|^2 (float bits)
  std::vector<int> h_info;
  std::vector<int> free_slots, pending;
  int next_slot = 0;
  static constexpr int MAX_SLOTS = 1 << 16;
  void* slab_alloc(size_t bytes, int* slab_idx);
  void slab_release(int idx, void* ptr, size_t bytes);
  void refresh_info();
  // u8: Hamming descriptors.  u8_as_l2: uint8 storage of an L2 descriptor (widened to float32 on the device).
  int add_async(const void* host, int n, int dim, bool u8, bool u8_as_l2 = false);  // no host sync; the host buffer must stay valid
  int add(const void* host, int n, int dim, bool u8, bool u8_as_l2 = false);
  void remove(int id);
  void clear();
  void free_set(DescSet& s);
  // guided: pose12 = npairs x 12 doubles [R cam2->cam1 row-major | origin of camera 2 in camera 1] and the angle
  // threshold; the epipolar mask is built on the device as a bitmask (matching.py:847-868)
  void match_pairs_async(int npairs, const int* ids_a, const int* ids_b, double ratio, bool symmetric,
                         const uint8_t* dmask, const double* pose12 = nullptr, double epi_threshold = 0.0);
  void set_bearings(int id, const float* host_n_by_3);
  void sync();
  void fetch(int32_t* out, int64_t capacity);
  long long fetch_pairs(long long* offsets_out, int32_t* pairs_out, long long capacity_rows);
  void last_ms(float* total, float* kernel);
  void one_shot(const void* f1, int n1, const void* f2, int n2, int dim, bool u8, double ratio,
                const uint8_t* mask, bool symmetric, int32_t* out);
  // match_tc.cu
  void prepare_tc(DescSet& s, const void* src, bool src_u8, float* padded_dst);
  void prepare_h8(DescSet& s, const uint8_t* src, int src_stride);   // +-1 fp8 operands of a Hamming set
};

// match_tc.cu
bool tc_available();
int tc_tile_m();
int tc_tile_n();
void launch_tc(Matcher& m, int njobs, int ntiles, bool masked);
int tc_rows_padded(int n);
size_t tc_operand_bytes(int rows_padded);
bool tc_capable(int dim, bool u8, int n);
// tensor-core Hamming (match_tc.cu)
int h8_tile_m();
int h8_tile_n();
size_t h8_operand_bytes(int rows_padded);
bool h8_capable(int nbytes, int n);
void launch_tc_h8(Matcher& m, int njobs, int ntiles);

}  // namespace osfm
