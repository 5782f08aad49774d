// Bundle adjustment on B200 (sm_100a): Levenberg-Marquardt with Schur
// elimination of the points and PCG on the reduced camera system, all in fp64.
//
// Replaces bundle::BundleAdjuster::Run (opensfm/src/bundle/src/bundle_adjuster.cc:595-1121)
// = problem assembly + ceres::Solve(SPARSE_SCHUR) + ComputeReprojectionErrors, for the
// residual blocks the reference's default pipeline builds (sfm/src/ba_helpers.cc:581-763):
// point projections with a shared robust loss, camera-intrinsics priors (log-scale focal /
// aspect ratio), rig-instance position priors.  The LM rules are Ceres' (SURVEY.md §8c).
//
// HBM layout (SoA, observations sorted by point so that V_p, g_p, the Schur update
// and the back-substitution of a point touch one contiguous range):
//   obs_{shot,point,x,y,isig}[N]          observation records
//   r[nres][N], Jc[nres*wc][N], Jp[nres*3][N]   robustified residual / Jacobian planes
//   S[nc*nc], rhs[nc]                     reduced camera system (dense, symmetric)
//   Vinv[6][npf], vectors[nc + 3 npf]     per-point inverse blocks, LM vectors
//
// Kernels: ba_linearize (per observation), ba_colnorm_grad, ba_schur (CTA per point:
// U, g_c, V^-1 and W V^-1 W^T scattered to S with fp64 atomics), ba_finish_system,
// PCG kernels (block-Jacobi), ba_backsub (warp per point), ba_model_change, ba_update.
#include <algorithm>
#include <dlfcn.h>

#include <chrono>
#include <cstring>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "ba_models.cuh"
#include "common.cuh"

// ncclUniqueId is 128 opaque bytes passed BY VALUE to ncclCommInitRank (nccl.h: NCCL_UNIQUE_ID_BYTES)
struct OsfmNcclId { char internal[128]; };

namespace osfm {

// ---------------------------------------------------------------------------
// Device-side problem view
// ---------------------------------------------------------------------------
struct BAView {
  int K, NI, NR, S, P;
  long long N;
  int nc, npf;      // reduced camera-side dimension, free local points
  int wc, nres;     // Jacobian plane counts
  int loss;
  double loss_a;
  const int *cam_type, *cam_off, *cam_np, *cam_poff, *inst_poff, *rc_poff, *pt_poff;
  const int *shot_inst, *shot_cam, *shot_rc, *shot_use_rc;
  const int *obs_shot, *obs_point;
  const double *obs_x, *obs_y, *obs_isig;
  const long long* obs_orig;
  const long long* pt_start;
  double *r, *Jc, *Jp;
};

struct Params {
  double *cam, *inst, *rc, *pts, *ext;
};

// scalar accumulators (device)
struct Scalars {
  double cost;
  double model_change;
  double step_norm2;
  double x_norm2;
  double grad_max_bits;  // max |g| via atomicMax on the bit pattern (non-negative doubles)
  double gdot;           // gradient . step (projected line search of bounded problems)
};

__device__ __forceinline__ double block_reduce_sum(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = threadIdx.x < nw ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// ---------------------------------------------------------------------------
// Per-observation kernels
// ---------------------------------------------------------------------------
// MODE 0: cost only (candidate evaluation).  MODE 1: residual + Jacobian planes (robustified).
// MODE 2: unscaled reprojection errors (bundle_adjuster.cc:531-566), written in original order.
// NB = minimum resident CTAs per SM the register allocation is held to (the kernel is latency-bound: at its
// natural 156 registers only 12 warps fit on an SM).
// TYPE >= 0: every camera of the problem has this projection type and no shot goes through a rig camera: the
// model dispatch, the parameter count and the Jacobian block sizes become compile-time constants, the blocks
// live in registers instead of a local-memory frame.
template <int MODE, int NB = 3, int TYPE = -1>
__global__ void __launch_bounds__(128, NB) ba_linearize(BAView v, Params p, Scalars* sc, double* reproj) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0.0;
  if (i < v.N) {
    const int shot = v.obs_shot[i];
    const int cam = v.shot_cam[shot];
    const int type = TYPE >= 0 ? TYPE : v.cam_type[cam];
    const int C = TYPE >= 0 ? model_num_params(TYPE >= 0 ? TYPE : 0) : v.cam_np[cam];
    const bool use_rc = TYPE >= 0 ? false : v.shot_use_rc[shot] != 0;
    const int pt = v.obs_point[i];
    double camp[MAX_CAM_PARAMS], ri[6], rc[6], X[3];
#pragma unroll
    for (int j = 0; j < MAX_CAM_PARAMS; ++j)
      if (j < C) camp[j] = p.cam[v.cam_off[cam] + j];
#pragma unroll
    for (int j = 0; j < 6; ++j) ri[j] = p.inst[6 * (size_t)v.shot_inst[shot] + j];
    if (use_rc) {
#pragma unroll
      for (int j = 0; j < 6; ++j) rc[j] = p.rc[6 * (size_t)v.shot_rc[shot] + j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) X[j] = p.pts[3 * (size_t)pt + j];
    const double ox = v.obs_x[i], oy = v.obs_y[i];
    if (MODE == 2) {
      double r[3] = {0.0, 0.0, 0.0};
      observation_eval(type, camp, ri, rc, use_rc, X, ox, oy, 1.0, r, nullptr, nullptr, nullptr, nullptr);
      const long long o = v.obs_orig[i];
      reproj[3 * o] = r[0]; reproj[3 * o + 1] = r[1]; reproj[3 * o + 2] = r[2];
    } else if (MODE == 0) {
      double r[3];
      const int nres = observation_eval(type, camp, ri, rc, use_rc, X, ox, oy, v.obs_isig[i], r, nullptr, nullptr,
                                        nullptr, nullptr);
      double s = r[0] * r[0] + r[1] * r[1];
      if (nres == 3) s += r[2] * r[2];
      double w;
      cost = 0.5 * robust_loss(v.loss, v.loss_a, s, &w);
      // a residual block whose parameter blocks are all constant is not part of the minimised cost (ceres removes
      // it from the reduced program; its value only enters Summary::fixed_cost)
      if (v.cam_poff[cam] < 0 && v.inst_poff[v.shot_inst[shot]] < 0 && (!use_rc || v.rc_poff[v.shot_rc[shot]] < 0) &&
          v.pt_poff[pt] < 0)
        cost = 0.0;
    } else {
      double r[3], jc[3 * MAX_CAM_PARAMS], jri[18], jrc[18], jp[9];
      const int nres = observation_eval(type, camp, ri, rc, use_rc, X, ox, oy, v.obs_isig[i], r, jc, jri, jrc, jp);
      double s = r[0] * r[0] + r[1] * r[1];
      if (nres == 3) s += r[2] * r[2];
      double w;
      cost = 0.5 * robust_loss(v.loss, v.loss_a, s, &w);
      const bool pfree = v.pt_poff[pt] >= 0;
      if (!pfree && v.cam_poff[cam] < 0 && v.inst_poff[v.shot_inst[shot]] < 0 && (!use_rc || v.rc_poff[v.shot_rc[shot]] < 0))
        cost = 0.0;   // all blocks constant: not part of the minimised cost (see MODE 0)
      const size_t N = (size_t)v.N;
      if (TYPE >= 0) {
        // uniform projection type, no rig cameras: nres = 2 rows, wc = C + 6 columns, all indices compile-time
        constexpr int CC = TYPE == PT_PERSPECTIVE ? 3 : TYPE == PT_FISHEYE ? 3 : TYPE == PT_BROWN ? 9 : 1;
        const int wcl = v.wc;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          v.r[k * N + i] = w * r[k];
          double* jcrow = v.Jc + (size_t)k * wcl * N + i;
#pragma unroll
          for (int j = 0; j < CC; ++j) jcrow[(size_t)j * N] = w * jc[k * CC + j];
#pragma unroll
          for (int j = 0; j < 6; ++j) jcrow[(size_t)(CC + j) * N] = w * jri[k * 6 + j];
          for (int j = CC + 6; j < wcl; ++j) jcrow[(size_t)j * N] = 0.0;
#pragma unroll
          for (int j = 0; j < 3; ++j) v.Jp[((size_t)k * 3 + j) * N + i] = pfree ? w * jp[k * 3 + j] : 0.0;
        }
        for (int k = 2; k < v.nres; ++k) {
          v.r[k * N + i] = 0.0;
          for (int j = 0; j < wcl; ++j) v.Jc[((size_t)k * wcl + j) * N + i] = 0.0;
          for (int j = 0; j < 3; ++j) v.Jp[((size_t)k * 3 + j) * N + i] = 0.0;
        }
      } else
      for (int k = 0; k < v.nres; ++k) {
        const bool live = k < nres;
        v.r[k * N + i] = live ? w * r[k] : 0.0;
        double* jcrow = v.Jc + (size_t)k * v.wc * N + i;
        for (int j = 0; j < v.wc; ++j) {
          double val = 0.0;
          if (live) {
            if (j < C) val = jc[k * C + j];
            else if (j < C + 6) val = jri[k * 6 + (j - C)];
            else if (use_rc && j < C + 12) val = jrc[k * 6 + (j - C - 6)];
          }
          jcrow[(size_t)j * N] = w * val;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) v.Jp[((size_t)k * 3 + j) * N + i] = (live && pfree) ? w * jp[k * 3 + j] : 0.0;
      }
    }
  }
  if (MODE != 2) {
    const double tot = block_reduce_sum(cost);
    if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&sc->cost, tot);
  }
}

// Global column of local camera-side column j of an observation, or -1.
struct ObsCols {
  int g_cam, C, g_inst, g_rc;
  __device__ __forceinline__ int col(int j) const {
    if (j < C) return g_cam >= 0 ? g_cam + j : -1;
    if (j < C + 6) return g_inst >= 0 ? g_inst + (j - C) : -1;
    if (j < C + 12) return g_rc >= 0 ? g_rc + (j - C - 6) : -1;
    return -1;
  }
};
__device__ __forceinline__ ObsCols obs_cols(const BAView& v, long long i) {
  const int shot = v.obs_shot[i];
  const int cam = v.shot_cam[shot];
  ObsCols oc;
  oc.g_cam = v.cam_poff[cam];
  oc.C = v.cam_np[cam];
  oc.g_inst = v.inst_poff[v.shot_inst[shot]];
  oc.g_rc = v.shot_use_rc[shot] ? v.rc_poff[v.shot_rc[shot]] : -2;  // -2: no rig-camera columns at all
  return oc;
}

// Prior residual rows (camera prior bundle_adjuster.cc:568-593 / prior_error.h:78-95,
// position prior bundle_adjuster.cc:745-778): row value, column, derivative.
struct PriorView {
  int n_cam_rows, n_pos_rows;
  const int *cam_row_param, *cam_row_col, *cam_row_log;  // index into cam params / reduced column
  const double *cam_row_prior, *cam_row_scale;
  // linear rows on one component of a rig instance (kind 1: GPS position prior) or of a rig camera
  // (kind 2: DataPriorError<Pose>, bundle_adjuster.cc:779-790)
  const int *pos_row_kind, *pos_row_inst, *pos_row_axis, *pos_row_col;
  const double *pos_row_prior, *pos_row_scale;
};
__device__ __forceinline__ void prior_row(const PriorView& pv, const Params& p, int row, double* r, int* col,
                                          double* d) {
  if (row < pv.n_cam_rows) {
    const double val = p.cam[pv.cam_row_param[row]];
    const double sc = pv.cam_row_scale[row];
    *col = pv.cam_row_col[row];
    if (pv.cam_row_log[row]) {
      *r = sc * log(val / pv.cam_row_prior[row]);
      *d = sc / val;
    } else {
      *r = sc * (val - pv.cam_row_prior[row]);
      *d = sc;
    }
  } else {
    const int q = row - pv.n_cam_rows;
    const double sc = pv.pos_row_scale[q];
    *col = pv.pos_row_col[q];
    const double* base = pv.pos_row_kind[q] == 2 ? p.rc : p.inst;
    *r = sc * (base[6 * (size_t)pv.pos_row_inst[q] + pv.pos_row_axis[q]] - pv.pos_row_prior[q]);
    *d = sc;
  }
}

__global__ void ba_prior_cost(PriorView pv, Params p, Scalars* sc) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  double c = 0.0;
  if (row < pv.n_cam_rows + pv.n_pos_rows) {
    double r, d;
    int col;
    prior_row(pv, p, row, &r, &col, &d);
    c = 0.5 * r * r;
  }
  const double tot = block_reduce_sum(c);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&sc->cost, tot);
}

// Point priors (AddPointPrior, bundle_adjuster.cc:224-236; residual block :688-708, DataPriorError<Vec3d>):
// r_j = d_j (X_j - x0_j), d_j = 1 / max(sigma_j, eps), j over x, y (and z with an altitude prior).  Dense arrays in
// the caller's point order (d = 0: no row); local point np -> global_of[np].  Points with a prior are kept off the
// segment path (ba_order.cuh), so only ba_schur has to add them to V_p and g_p.
struct PointPriorView {
  const double* d;   // [3 * Pfull], nullptr when no point has a prior
  const double* x0;
  const int* global_of;
};
// MODE 0: cost.  1: cost + squared column norms + gradient.  2: model cost change.  3: adds J_p^T r to the
// right-hand side t[3][npf] of the back-substitution.  One thread per local point.
template <int MODE>
__global__ void ba_point_prior(PointPriorView pp, BAView v, Params p, const double* __restrict__ scale,
                               const double* __restrict__ y, double* colnorm2, double* grad, double* t, Scalars* sc) {
  const int np = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (np < v.P) {
    const int pf = v.pt_poff[np];
    const size_t g = (size_t)pp.global_of[np];
    if (pf >= 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double d = pp.d[3 * g + j];
        if (d == 0.0) continue;
        const double r = d * (p.pts[3 * (size_t)np + j] - pp.x0[3 * g + j]);
        const int col = v.nc + 3 * pf + j;
        if (MODE <= 1) acc += 0.5 * r * r;
        if (MODE == 1) { colnorm2[col] += d * d; grad[col] += d * r; }   // only this thread touches the point's columns here
        if (MODE == 2) { const double m = -d * scale[col] * y[col]; acc += -m * (r + 0.5 * m); }
        if (MODE == 3) t[(size_t)j * v.npf + pf] += d * r;
      }
    }
  }
  if (MODE <= 2) {
    const double tot = block_reduce_sum(acc);
    if (threadIdx.x == 0 && tot != 0.0) atomicAdd(MODE == 2 ? &sc->model_change : &sc->cost, tot);
  }
}

// Squared column norms and gradient of the (unscaled, robustified) Jacobian.
// (camera-side columns of the observations i >= i0: the ones no segment covers)
__global__ void __launch_bounds__(256) ba_colnorm_grad(BAView v, long long i0, double* colnorm2, double* grad) {
  const long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.N) return;
  const size_t N = (size_t)v.N;
  const ObsCols oc = obs_cols(v, i);
  double r[3];
  for (int k = 0; k < v.nres; ++k) r[k] = v.r[k * N + i];
  for (int j = 0; j < v.wc; ++j) {
    const int g = oc.col(j);
    if (g < 0) continue;
    double n2 = 0.0, gr = 0.0;
    for (int k = 0; k < v.nres; ++k) {
      const double a = v.Jc[((size_t)k * v.wc + j) * N + i];
      n2 += a * a;
      gr += a * r[k];
    }
    atomicAdd(&colnorm2[g], n2);
    atomicAdd(&grad[g], gr);
  }
}

// Segmented sum over lanes holding consecutive observations of the same point (observations are
// sorted by point): after the call the LAST lane of every run holds the run's total.
template <int NV>
__device__ __forceinline__ void seg_scan_by_point(int pt, double (&val)[NV]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int other = __shfl_up_sync(0xffffffffu, pt, o);
    const bool take = lane >= o && other == pt;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double up = __shfl_up_sync(0xffffffffu, val[k], o);
      if (take) val[k] += up;
    }
  }
}

// Point-side columns: one thread per observation (coalesced plane reads), run totals by shuffles,
// one atomic per (run, column) -- a run is cut only at warp boundaries.
__global__ void __launch_bounds__(256) ba_colnorm_grad_points(BAView v, double* colnorm2, double* grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t N = (size_t)v.N;
  const bool in = i < v.N;
  const int pt = in ? v.obs_point[i] : -1;
  double val[6] = {0, 0, 0, 0, 0, 0};
  if (in) {
    for (int k = 0; k < v.nres; ++k) {
      const double rk = v.r[k * N + i];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double a = v.Jp[((size_t)k * 3 + j) * N + i];
        val[j] += a * a;
        val[3 + j] += a * rk;
      }
    }
  }
  seg_scan_by_point<6>(pt, val);
  const int next = __shfl_down_sync(0xffffffffu, pt, 1);
  const bool tail = (threadIdx.x & 31) == 31 || next != pt;
  if (in && tail) {
    const int pf = v.pt_poff[pt];
    if (pf >= 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        atomicAdd(&colnorm2[v.nc + 3 * pf + j], val[j]);
        atomicAdd(&grad[v.nc + 3 * pf + j], val[3 + j]);
      }
    }
  }
}

// Camera-side columns of the segments (points seen by exactly the same k shots, ba_order.cuh): one
// warp per segment, lane = (point slot g, shot c) so that a lane always works on the same shot; the
// running sums stay in registers and a segment issues k * wc atomics instead of points * k * wc.
template <int WCT>
__global__ void __launch_bounds__(256, WCT ? 3 : 2) ba_colnorm_grad_seg(BAView v, const int* __restrict__ seg_start, int nseg,
                                                           double* colnorm2, double* grad) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= nseg) return;
  const int lane = threadIdx.x & 31;
  const int p0 = seg_start[s], np = seg_start[s + 1] - p0;
  const long long base = v.pt_start[p0];
  const int k = (int)(v.pt_start[p0 + 1] - base);
  const int G = 32 / k;
  const int g = lane / k, c = lane - g * k;
  const bool active = g < G;
  const size_t N = (size_t)v.N;
  const int wc = WCT ? WCT : v.wc;
  constexpr int WCU = WCT ? WCT : 16;
  double n2[WCU], gr[WCU];
#pragma unroll
  for (int j = 0; j < WCU; ++j) { n2[j] = 0.0; gr[j] = 0.0; }
  if (active) {
    for (int pi = g; pi < np; pi += G) {
      const size_t i = (size_t)(base + (long long)pi * k + c);
      for (int q = 0; q < v.nres; ++q) {
        const double rq = v.r[q * N + i];
#pragma unroll
        for (int j = 0; j < WCU; ++j) {
          if (j < wc) {
            const double a = v.Jc[((size_t)q * wc + j) * N + i];
            n2[j] += a * a;
            gr[j] += a * rq;
          }
        }
      }
    }
  }
  // lanes c, c + k, c + 2k, ... hold partial sums of the same shot
  for (int gg = 1; gg < G; ++gg) {
#pragma unroll
    for (int j = 0; j < WCU; ++j) {
      if (j < wc) {
        const double a = __shfl_down_sync(0xffffffffu, n2[j], gg * k);
        const double b = __shfl_down_sync(0xffffffffu, gr[j], gg * k);
        if (lane < k) { n2[j] += a; gr[j] += b; }
      }
    }
  }
  if (lane < k) {
    const ObsCols oc = obs_cols(v, base + c);
#pragma unroll
    for (int j = 0; j < WCU; ++j) {
      if (j < wc) {
        const int col = oc.col(j);
        if (col >= 0) {
          atomicAdd(&colnorm2[col], n2[j]);
          atomicAdd(&grad[col], gr[j]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same camera-side sums with the observation tiles staged by the TMA engine (cp.async.bulk + mbarrier).
// The residual / Jacobian planes of a chunk of a segment (<= 8 points seen by the same k shots) are, per plane,
// one contiguous run of 8 * np * k bytes: a warp stages the 2 + 2 * 9 runs of a chunk into its shared-memory
// tile with one bulk copy each (lane 0 issues them, all complete on one mbarrier), double-buffered so that the
// copy of the next chunk is in flight while the current one is reduced.  Lane l owns the accumulators
// (shot c, column j) = l, l + 32, l + 64 and walks the points of the chunk in shared memory: no shuffles, no
// atomics until the segment's k * 9 sums are flushed.  wc == 9, nres == 2 (one 3-parameter camera + pose per
// shot: BASELINE configs[1..3]); chunks whose runs are not 16-byte aligned are staged with plain loads.
// ---------------------------------------------------------------------------------------------------------
constexpr int CG_WARPS = 4;
constexpr int CG_ROWS = 20;                         // r[2] + Jc[2 * 9]
constexpr int CG_PCHUNK = 8, CG_KMAX = 16;          // points per chunk, shots per segment (= SEG_KMAX, ba_reduced.cuh)
constexpr int CG_OBS = CG_PCHUNK * CG_KMAX;         // observations per chunk
constexpr int CG_STAGE_DOUBLES = CG_ROWS * CG_OBS;
constexpr int CG_SMEM = CG_WARPS * 2 * CG_STAGE_DOUBLES * (int)sizeof(double);   // 163,840 bytes

__device__ __forceinline__ uint32_t cg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cg_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = cg_smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
}

__global__ void __launch_bounds__(32 * CG_WARPS, 1)
    ba_colnorm_grad_tma(BAView v, const int* __restrict__ seg_start, int nseg, const long long* __restrict__ tab_off,
                        const int* __restrict__ tab, double* colnorm2, double* grad) {
  extern __shared__ __align__(128) double cg_tiles[];
  __shared__ __align__(8) uint64_t bars[CG_WARPS][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int w = 0; w < CG_WARPS; ++w)
      for (int st = 0; st < 2; ++st)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cg_smem_u32(&bars[w][st])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double* tile[2] = {cg_tiles + (size_t)(warp * 2) * CG_STAGE_DOUBLES, cg_tiles + (size_t)(warp * 2 + 1) * CG_STAGE_DOUBLES};
  const size_t N = (size_t)v.N;
  const int gw = blockIdx.x * CG_WARPS + warp, nw = gridDim.x * CG_WARPS;
  uint32_t phase[2] = {0u, 0u};

  // stage chunk [pc0, pc0 + np) of a segment with k shots into tile[st]; returns whether the TMA path was used
  auto stage = [&](int st, int pc0, int np, int k) -> bool {
    const long long ibase = v.pt_start[pc0];
    const int run = np * k;
    const bool aligned = ((ibase | (long long)run | (long long)N) & 1LL) == 0;   // 16-byte aligned runs in every plane
    if (aligned) {
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)run * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cg_smem_u32(&bars[warp][st])),
                     "r"(bytes * CG_ROWS)
                     : "memory");
        for (int row = 0; row < CG_ROWS; ++row) {
          const double* src = (row < 2 ? v.r + (size_t)row * N : v.Jc + (size_t)(row - 2) * N) + ibase;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           cg_smem_u32(tile[st] + row * CG_OBS)),
                       "l"(src), "r"(bytes), "r"(cg_smem_u32(&bars[warp][st]))
                       : "memory");
        }
      }
    } else {
      for (int idx = lane; idx < CG_ROWS * run; idx += 32) {
        const int row = idx / run, e = idx - row * run;
        tile[st][row * CG_OBS + e] = (row < 2 ? v.r + (size_t)row * N : v.Jc + (size_t)(row - 2) * N)[ibase + e];
      }
    }
    return aligned;
  };

  // (segment, chunk) walk of this warp, one step ahead for the prefetch
  int s = gw, pc = 0, p_end = 0, k = 0;
  auto open_segment = [&]() {
    while (s < nseg) {
      pc = seg_start[s]; p_end = seg_start[s + 1];
      if (pc < p_end) { k = (int)(v.pt_start[pc + 1] - v.pt_start[pc]); return true; }
      s += nw;
    }
    return false;
  };
  if (!open_segment()) return;
  int st = 0;
  bool cur_tma = stage(0, pc, min(CG_PCHUNK, p_end - pc), k);
  double n2[3] = {0.0, 0.0, 0.0}, gr[3] = {0.0, 0.0, 0.0};
  while (true) {
    const int np = min(CG_PCHUNK, p_end - pc);
    const int cur_k = k, cur_pc = pc;
    const bool last_chunk = pc + np >= p_end;
    // next (segment, chunk)
    int ns = s, npc = pc + np, np_end = p_end, nk = k;
    bool have_next = true;
    if (last_chunk) {
      ns = s + nw;
      have_next = false;
      while (ns < nseg) {
        npc = seg_start[ns]; np_end = seg_start[ns + 1];
        if (npc < np_end) { nk = (int)(v.pt_start[npc + 1] - v.pt_start[npc]); have_next = true; break; }
        ns += nw;
      }
    }
    bool next_tma = false;
    if (have_next) next_tma = stage(st ^ 1, npc, min(CG_PCHUNK, np_end - npc), nk);
    if (cur_tma) { cg_mbar_wait(&bars[warp][st], phase[st]); phase[st] ^= 1u; }
    else __syncwarp();
    // reduce the chunk: accumulator a = c * 9 + j
    const double* T = tile[st];
    const int nacc = cur_k * 9;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int a = lane + 32 * u;
      if (a < nacc) {
        const int c = a / 9, j = a - c * 9;
        for (int pi = 0; pi < np; ++pi) {
          const int e = pi * cur_k + c;
          const double j0 = T[(2 + j) * CG_OBS + e], j1 = T[(2 + 9 + j) * CG_OBS + e];
          n2[u] += j0 * j0 + j1 * j1;
          gr[u] += j0 * T[e] + j1 * T[CG_OBS + e];
        }
      }
    }
    __syncwarp();   // the tile may be overwritten by the copy issued two steps from now
    if (last_chunk) {
      const long long base = v.pt_start[seg_start[s]];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int a = lane + 32 * u;
        if (a < nacc) {
          int col;
          if (tab) {   // per-segment column table (ba_seg_tables): one load instead of the shot -> camera -> offset chain
            col = tab[tab_off[s] + a];
          } else {
            const int c = a / 9, j = a - c * 9;
            const ObsCols oc = obs_cols(v, base + c);
            col = oc.col(j);
          }
          if (col >= 0) { atomicAdd(&colnorm2[col], n2[u]); atomicAdd(&grad[col], gr[u]); }
        }
        n2[u] = 0.0; gr[u] = 0.0;
      }
    }
    (void)cur_pc;
    if (!have_next) break;
    s = ns; pc = npc; p_end = np_end; k = nk;
    cur_tma = next_tma;
    st ^= 1;
  }
}

__global__ void ba_prior_colnorm_grad(PriorView pv, Params p, double* colnorm2, double* grad) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= pv.n_cam_rows + pv.n_pos_rows) return;
  double r, d;
  int col;
  prior_row(pv, p, row, &r, &col, &d);
  atomicAdd(&colnorm2[col], d * d);
  atomicAdd(&grad[col], d * r);
}

// scale = 1 / (1 + sqrt(colnorm2))   (Ceres jacobi_scaling)
__global__ void ba_make_scale(const double* colnorm2, double* scale, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scale[i] = 1.0 / (1.0 + sqrt(colnorm2[i]));
}
// diag2 = clamp(colnorm2 * scale^2, 1e-6, 1e32) / radius ; also gradient max-norm
__global__ void ba_make_diag(const double* colnorm2, const double* scale, double* diag, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) diag[i] = fmin(fmax(colnorm2[i] * scale[i] * scale[i], 1e-6), 1e32);
}
// Multi-GPU: everything rank-summed after a linearisation travels in ONE buffer
//   pack = [ colnorm2 (nc) | grad (nc) | cost | per-rank max |point gradient| (world) ]
__global__ void ba_pack_lin(const double* __restrict__ colnorm2, const double* __restrict__ grad, const Scalars* sc, int nc,
                            int rank, int world, double* __restrict__ pack) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nc) { pack[i] = colnorm2[i]; pack[nc + i] = grad[i]; }
  if (i == 0) pack[2 * nc] = sc->cost;
  if (i < world) pack[2 * nc + 1 + i] = i == rank ? sc->grad_max_bits : 0.0;
}
__global__ void ba_unpack_lin(const double* __restrict__ pack, int nc, int world, double* __restrict__ colnorm2,
                              double* __restrict__ grad, Scalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < nc) { colnorm2[i] = pack[i]; grad[i] = pack[nc + i]; v = fabs(pack[nc + i]); }
  if (i < world) v = fmax(v, pack[2 * nc + 1 + i]);
  if (i == 0) sc->cost = pack[2 * nc];
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomic_max_nonneg(&sc->grad_max_bits, v);
}
__global__ void ba_grad_max(const double* grad, int n, Scalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = i < n ? fabs(grad[i]) : 0.0;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomic_max_nonneg(&sc->grad_max_bits, v);
}

}  // namespace osfm
#include "ba_reduced.cuh"
#include "ba_schur_pipe.cuh"

#include <tuple>
#include <utility>

namespace osfm {
// The PCG kernels synchronise the whole grid through their own barrier (flags in global memory): every CTA must be
// resident at the same time.  A cooperative launch makes the runtime check that (it fails with
// cudaErrorCooperativeLaunchTooLarge instead of deadlocking when the grid cannot be co-resident).
template <typename... KArgs, size_t... I>
inline void launch_cooperative_impl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st,
                                    std::tuple<KArgs...>& a, std::index_sequence<I...>) {
  void* ptrs[] = {static_cast<void*>(&std::get<I>(a))...};
  OSFM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(block), ptrs, smem, st));
}
template <typename... KArgs, typename... Args>
inline void launch_cooperative(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  std::tuple<KArgs...> a(std::forward<Args>(args)...);
  launch_cooperative_impl(kern, grid, block, smem, st, a, std::index_sequence_for<KArgs...>{});
}
}  // namespace osfm
static_assert(osfm::CG_KMAX == osfm::SEG_KMAX, "colnorm tiles hold the widest segment");
#include "ba_side.cuh"
#include "ba_order.cuh"
namespace osfm {

// ---------------------------------------------------------------------------
// Back-substitution: y_p = V^-1 (g_p - W^T y_c)  (scaled system).
// ba_backsub_rows: one thread per observation (coalesced plane reads) computes its share of
// g_p - W^T y_c, run totals by shuffles, one atomic per (run, coordinate) into t[3][npf];
// ba_backsub_points: one thread per point applies V^-1.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_backsub_rows(BAView v, const double* __restrict__ scale,
                                                       const double* __restrict__ y, double* __restrict__ t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t N = (size_t)v.N;
  const bool in = i < v.N;
  const int pt = in ? v.obs_point[i] : -1;
  const int pf = in ? v.pt_poff[pt] : -1;
  double val[3] = {0.0, 0.0, 0.0};
  if (pf >= 0) {
    const int wc = v.wc;
    const ObsCols oc = obs_cols(v, i);
    for (int q = 0; q < v.nres; ++q) {
      double e = v.r[q * N + i];
      for (int j = 0; j < wc; ++j) {
        const int g = oc.col(j);
        if (g >= 0) e -= v.Jc[((size_t)q * wc + j) * N + i] * scale[g] * y[g];
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) val[j] += v.Jp[((size_t)q * 3 + j) * N + i] * e;
    }
  }
  seg_scan_by_point<3>(pt, val);
  const int next = __shfl_down_sync(0xffffffffu, pt, 1);
  const bool tail = (threadIdx.x & 31) == 31 || next != pt;
  if (pf >= 0 && tail) {
    const size_t NP = (size_t)v.npf;
#pragma unroll
    for (int j = 0; j < 3; ++j) atomicAdd(&t[j * NP + pf], val[j]);
  }
}
__global__ void __launch_bounds__(256) ba_backsub_points(BAView v, const double* __restrict__ scale,
                                                         const double* __restrict__ Vinv, const double* __restrict__ t,
                                                         double* __restrict__ y) {
  const int pf = blockIdx.x * blockDim.x + threadIdx.x;
  if (pf >= v.npf) return;
  const size_t NP = (size_t)v.npf;
  const int nc = v.nc;
  const double t0 = t[pf] * scale[nc + 3 * pf], t1 = t[NP + pf] * scale[nc + 3 * pf + 1],
               t2 = t[2 * NP + pf] * scale[nc + 3 * pf + 2];
  const double a = Vinv[0 * NP + pf], b = Vinv[1 * NP + pf], c = Vinv[2 * NP + pf];
  const double d = Vinv[3 * NP + pf], e = Vinv[4 * NP + pf], f = Vinv[5 * NP + pf];
  y[nc + 3 * pf + 0] = a * t0 + b * t1 + c * t2;
  y[nc + 3 * pf + 1] = b * t0 + d * t1 + e * t2;
  y[nc + 3 * pf + 2] = c * t0 + e * t1 + f * t2;
}

// Ceres: model_cost_change = -sum m (r + m/2), m = Js step ; step = -y
__global__ void __launch_bounds__(256) ba_model_change(BAView v, const double* __restrict__ scale,
                                                       const double* __restrict__ y, Scalars* sc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double tot = 0.0;
  if (i < v.N) {
    const size_t N = (size_t)v.N;
    const int nc = v.nc, wc = v.wc;
    const ObsCols oc = obs_cols(v, i);
    const int pf = v.pt_poff[v.obs_point[i]];
    for (int q = 0; q < v.nres; ++q) {
      double m = 0.0;
      for (int j = 0; j < wc; ++j) {
        const int g = oc.col(j);
        if (g >= 0) m -= v.Jc[((size_t)q * wc + j) * N + i] * scale[g] * y[g];
      }
      if (pf >= 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) m -= v.Jp[((size_t)q * 3 + j) * N + i] * scale[nc + 3 * pf + j] * y[nc + 3 * pf + j];
      }
      tot += -m * (v.r[q * N + i] + 0.5 * m);
    }
  }
  const double t = block_reduce_sum(tot);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(&sc->model_change, t);
}
__global__ void ba_prior_model_change(PriorView pv, Params p, const double* scale, const double* y, Scalars* sc) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  double tot = 0.0;
  if (row < pv.n_cam_rows + pv.n_pos_rows) {
    double r, d;
    int col;
    prior_row(pv, p, row, &r, &col, &d);
    const double m = -d * scale[col] * y[col];
    tot = -m * (r + 0.5 * m);
  }
  const double t = block_reduce_sum(tot);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(&sc->model_change, t);
}

// candidate = x - scale * y ; accumulates |delta|^2 and |x|^2 (free parameters only).
// which: 0 cameras, 1 instances, 2 rig cameras, 3 points.
__global__ void ba_update(int which, int count, const int* __restrict__ poff, const int* __restrict__ off,
                          const int* __restrict__ np, int stride, int base, const double* __restrict__ src,
                          double* __restrict__ dst, const double* __restrict__ scale, const double* __restrict__ y,
                          Scalars* sc, int accumulate_norms, const double* __restrict__ lower = nullptr,
                          double alpha = 1.0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  double sn = 0.0, xn = 0.0;
  if (b < count) {
    const int o = which == 0 ? off[b] : b * stride;
    const int n = which == 0 ? np[b] : stride;
    const int g = poff[b];
    for (int j = 0; j < n; ++j) {
      double val = src[o + j];
      if (g >= 0) {
        const int gi = which == 3 ? base + 3 * g + j : g + j;
        double d = -alpha * scale[gi] * y[gi];
        if (lower && val + d < lower[o + j]) d = lower[o + j] - val;   // projection onto the bound (ceres bounded LM)
        xn += val * val;
        sn += d * d;
        val += d;
      }
      dst[o + j] = val;
    }
  }
  if (accumulate_norms) {
    const double t1 = block_reduce_sum(sn);
    const double t2 = block_reduce_sum(xn);
    if (threadIdx.x == 0) {
      if (t1 != 0.0) atomicAdd(&sc->step_norm2, t1);
      if (t2 != 0.0) atomicAdd(&sc->x_norm2, t2);
    }
  }
}

// gradient . delta with delta = -scale * y: camera side counted when cam_side != 0 (one rank), points always
__global__ void ba_grad_dot(int n, int nc, int cam_side, const double* __restrict__ grad, const double* __restrict__ scale,
                            const double* __restrict__ y, Scalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n && (i >= nc || cam_side)) v = -grad[i] * scale[i] * y[i];
  const double t = block_reduce_sum(v);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(&sc->gdot, t);
}

// Model cost change of the step without another pass over the Jacobian.  Ceres evaluates
//   model_cost_change = -step^T (g + H step / 2),   H = J^T J, g = J^T r   (trust_region_minimizer.cc)
// from J step; the step solves (H + D) step = -g (D = LM diagonal / radius), so H step = -g - D step and
//   model_cost_change = (y^T g_s + y^T D y) / 2,   step = -y, g_s = scale * grad   (scaled variables)
// exactly when the linear system is solved exactly, and to the solver's 1e-8 relative residual here (every term of H
// and g -- observations, priors, side terms -- is in the system that was solved).  Replaces ba_model_change +
// ba_prior_model_change + side_model_change + ba_point_prior<2>: 447 MB of plane reads per LM iteration.
__global__ void ba_model_change_alg(int n, int nc, int cam_side, const double* __restrict__ grad, const double* __restrict__ scale,
                                    const double* __restrict__ diag, double inv_radius, const double* __restrict__ y, Scalars* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n && (i >= nc || cam_side)) v = 0.5 * y[i] * (grad[i] * scale[i] + diag[i] * inv_radius * y[i]);
  const double t = block_reduce_sum(v);
  if (threadIdx.x == 0 && t != 0.0) atomicAdd(&sc->model_change, t);
}

// single-observation device evaluation (test hook)
__global__ void ba_eval_one(int type, const double* in, int use_rc, double* out, int* nres_out) {
  // in: cam[16] ri[6] rc[6] X[3] obs[2] isig[1]
  double r[3] = {0, 0, 0}, jc[3 * MAX_CAM_PARAMS], jri[18], jrc[18], jp[9];
  for (int i = 0; i < 18; ++i) jrc[i] = 0.0;
  const int nres = observation_eval(type, in, in + 16, in + 22, use_rc != 0, in + 28, in[31], in[32], in[33], r, jc,
                                    jri, jrc, jp);
  *nres_out = nres;
  for (int i = 0; i < 3; ++i) out[i] = r[i];
  for (int i = 0; i < 48; ++i) out[3 + i] = jc[i];
  for (int i = 0; i < 18; ++i) out[51 + i] = jri[i];
  for (int i = 0; i < 18; ++i) out[69 + i] = jrc[i];
  for (int i = 0; i < 9; ++i) out[87 + i] = jp[i];
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
// CUDA-event stopwatch on the launching stream; collect() after a stream synchronize.
struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  double total_ms = 0.0;
  long long count = 0;
  bool pending = false;
  void init() { OSFM_CUDA(cudaEventCreate(&a)); OSFM_CUDA(cudaEventCreate(&b)); }
  void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); a = b = nullptr; }
  void start(cudaStream_t st) { collect(); OSFM_CUDA(cudaEventRecord(a, st)); }
  void stop(cudaStream_t st) { OSFM_CUDA(cudaEventRecord(b, st)); pending = true; }
  void collect() {
    if (!pending) return;
    OSFM_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    OSFM_CUDA(cudaEventElapsedTime(&ms, a, b));
    total_ms += ms; ++count; pending = false;
  }
};

template <class T>
static void upload(DevBuf<T>& d, const std::vector<T>& h, cudaStream_t st) {
  d.reserve(std::max<size_t>(h.size(), 1));
  if (!h.empty()) OSFM_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

// ---------------------------------------------------------------------------
// NCCL, loaded at run time (libnccl.so.2: the copy torch already mapped if there is one, so that one
// process never mixes two NCCL versions).  Only the five entry points the all-reduce needs; the
// constants are nccl.h's (ncclFloat64 = 8, ncclSum = 0, ncclUniqueId = 128 bytes).
// ---------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, OsfmNcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return;
    api.lib = h;
    api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<int (*)(void**, int, OsfmNcclId, int)>(dlsym(h, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy)
    throw std::runtime_error("NCCL (libnccl.so.2) is not available in this process");
  return api;
}
static void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    NcclApi& a = nccl_api();
    throw std::runtime_error(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "NCCL error"));
  }
}

struct BA {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // osfm_ba_set_observations_async: the measurement arrays travel on their own stream while run() already sorts the indices
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_obs = nullptr;
  bool obs_pending = false;
  // host copies of the problem
  std::vector<int> cam_type, cam_const, cam_prior_log;
  std::vector<double> cam_params, cam_prior, cam_prior_sigma;
  std::vector<double> inst, inst_prior_pos, inst_prior_std;
  std::vector<int> inst_const, inst_has_prior;
  std::vector<double> rc, rc_prior, rc_prior_sigma;   // rig-camera pose priors: empty = none
  std::vector<int> rc_const;
  // ext blocks (biases, reconstruction scales, std-deviation scales), side terms, point priors
  std::vector<int> ext_size, ext_const;
  std::vector<double> ext_values, ext_lower;
  std::vector<osfm_side_term> side_terms;
  std::vector<double> side_consts;
  std::vector<int> pp_point, pp_alt;
  std::vector<double> pp_prior, pp_sigma;
  std::vector<int> shot_inst, shot_cam, shot_rc, shot_use_rc;
  std::vector<double> pts;
  std::vector<int> pt_const;
  long long n_obs_full = 0;  // observations live on the device only (d_raw_*)
  // options (defaults of bundle::BundleAdjuster(), bundle_adjuster.cc:24-44)
  int loss = OSFM_LOSS_CAUCHY;
  double loss_a = 1.0;
  int max_iterations = 500;
  bool compute_reproj = true;
  int rank = 0, world = 1;
  osfm_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  void* nccl_comm = nullptr;   // own communicator (osfm_ba_set_nccl); used when no callback is set
  int nccl_rank = -1, nccl_world = 0;
  // results
  bool reproj_valid = false;
  osfm_ba_summary summary{};
  bool has_run = false;

  // device state
  DevBuf<int> d_cam_type, d_cam_off, d_cam_np, d_cam_poff, d_inst_poff, d_rc_poff, d_pt_poff;
  DevBuf<int> d_shot_inst, d_shot_cam, d_shot_rc, d_shot_use_rc, d_obs_shot, d_obs_point;
  DevBuf<double> d_obs_x, d_obs_y, d_obs_isig;
  DevBuf<long long> d_obs_orig, d_pt_start;
  DevBuf<double> d_cam[2], d_inst[2], d_rc[2], d_pts[2];
  DevBuf<double> d_r, d_Jc, d_Jp, d_Sbuf, d_Vinv, d_gp, d_slots;
  DevBuf<double> d_scale, d_colnorm2, d_grad, d_diag, d_y, d_bs_t;
  DevBuf<double> d_px, d_pr, d_pz, d_pp, d_pAp, d_Minv, d_reproj, d_full_pts;
  DevBuf<double> d_Wdef;                   // [PCG_ND][nc] deflation vectors of the pipelined PCG (pcg_gauge_vectors)
  DevBuf<int> d_blk_off, d_blk_sz, d_cam_blk, d_inst_blk, d_rc_blk;
  // block-sparse reduced system (ba_reduced.cuh)
  DevBuf<unsigned long long> d_tkeys, d_skeys, d_skeys2, d_rkeys, d_rkeys2;
  DevBuf<int> d_tvals, d_area, d_offs, d_row_ptr, d_row_col, d_row_off, d_diag_off, d_prior_diag_off;
  DevBuf<int> d_pr_blk, d_pr_local, d_g_obs_shot, d_g_obs_point;
  DevBuf<long long> d_g_pt_start;
  DevBuf<int4> d_upper;
  DevBuf<unsigned> d_count;
  DevBuf<char> d_cub;
  DevBuf<PcgState> d_pcg;
  DevBuf<int> d_seg_start;
  // raw observations as the caller gave them + scratch of the device-side ordering (ba_order.cuh)
  DevBuf<int> d_raw_shot, d_raw_point, d_ptc_full;
  DevBuf<double> d_raw_xy, d_raw_sigma, d_pts_in;
  DevBuf<unsigned long long> d_okeys, d_okeys2, d_pkey, d_pkey2;
  DevBuf<int> d_ovals, d_ovals2, d_pval, d_order, d_inv_order, d_global_of, d_free_flag, d_free_scan, d_head, d_run_head, d_nsel;
  DevBuf<long long> d_kk;
  DevBuf<char> d_seg_flags;
  DevBuf<OrderCounts> d_oc;
  PinnedBuf<OrderCounts> h_oc;
  DevBuf<double> d_rowsJ, d_rowsW, d_rowsY, d_Vig;
  DevBuf<int> d_row_M, d_qoff, d_blk_row, d_cbase, d_colidx, d_row_of, d_grp_b1, d_grp_b2;
  DevBuf<long long> d_rowbase;
  DevBuf<double> d_Spcg, d_Ap;
  PinnedBuf<PcgState> h_pcg;
  DevBuf<int> d_pcg_rowlo;
  PcgResident pcg_res{};
  bool pcg_resident = false;
  int pcg_smem = 0;
  DevBuf<int> d_pcg_grplo;
  DevBuf<unsigned long long> d_prof;
  bool seg_attr = false, mma_attr = false, cg_attr = false;   // per handle (= per device): dynamic shared-memory opt-in done
  DevBuf<long long> d_tab_off, d_tab_sizes;
  DevBuf<int> d_sp_nch, d_sp_chunk0;       // chunks per segment / first chunk of every segment (ba_schur_pipe)
  DevBuf<SchurChunk> d_sp_chunks;
  DevBuf<int> d_sp_ftab;                   // flush destinations per segment (sp_flush_tables)
  int sp_nchunks = 0;
  bool cc_attr = false;
  bool have_seg_tab = false;               // ba_seg_tables has run for this problem (gcol of every segment column)
  bool sp_attr = false;
  DevBuf<int> d_tab;
  PcgPipe pcg_pipe{};
  bool pcg_pipe_ok = false;
  int pcg_pipe_smem = 0, pcg_pipe_its = 0, pcg_fallbacks = 0;
  int num_sms = 148;
  DevBuf<int> d_pr_cam_param, d_pr_cam_col, d_pr_cam_log, d_pr_pos_inst, d_pr_pos_axis, d_pr_pos_col;
  DevBuf<double> d_pr_cam_prior, d_pr_cam_scale, d_pr_pos_prior, d_pr_pos_scale;
  DevBuf<Scalars> d_sc;
  PinnedBuf<Scalars> h_sc;
  DevBuf<double> d_eval;
  DevBuf<int> d_pr_pos_kind, d_ext_off, d_ext_np, d_ext_poff, d_ext_blk, d_side_jofs, d_side_rofs;
  DevBuf<double> d_ext[2], d_ext_lower, d_side_consts, d_side_J, d_side_r, d_pp_d, d_pp_x0;
  DevBuf<SideTerm> d_side_terms;

  explicit BA(int dev) : device(dev) {
    OSFM_CUDA(cudaSetDevice(device));
    OSFM_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
    stream = own_stream;
    OSFM_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    OSFM_CUDA(cudaEventCreateWithFlags(&ev_obs, cudaEventDisableTiming));
    h_sc.reserve(1);
    h_pcg.reserve(1);
    OSFM_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  }
  ~BA() {
    cudaSetDevice(device);
    if (nccl_comm) { try { nccl_api().CommDestroy(nccl_comm); } catch (...) {} }
    if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
    if (ev_obs) cudaEventDestroy(ev_obs);
    if (own_stream) cudaStreamDestroy(own_stream);
  }

  Scalars read_scalars() {
    OSFM_CUDA(cudaMemcpyAsync(h_sc.p, d_sc.p, sizeof(Scalars), cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaStreamSynchronize(stream));
    return *h_sc.p;
  }
  // OSFM_BA_TRACE: number of all-reduces and the host time spent issuing them (+ device time when traced)
  int ar_calls = 0;
  double ar_host_ms = 0.0, ar_dev_ms = 0.0;
  bool ar_trace = false;
  void allreduce_dev(double* buf, long long count) {
    if (world > 1) {
      const auto t0 = std::chrono::high_resolution_clock::now();
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (ar_trace) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, stream); }
      struct Done {
        BA* self; std::chrono::high_resolution_clock::time_point t0; cudaEvent_t e0, e1;
        ~Done() {
          if (self->ar_trace) {
            cudaEventRecord(e1, self->stream); cudaEventSynchronize(e1);
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); self->ar_dev_ms += ms;
            cudaEventDestroy(e0); cudaEventDestroy(e1);
          }
          ++self->ar_calls;
          self->ar_host_ms += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        }
      } done{this, t0, e0, e1};
      if (allreduce) {
        if (allreduce(buf, count, stream, allreduce_user) != 0) throw std::runtime_error("all-reduce callback failed");
      } else if (nccl_comm && nccl_world == world && nccl_rank == rank) {
        nccl_check(nccl_api().AllReduce(buf, buf, (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, nccl_comm, stream), "ncclAllReduce");
      } else {
        throw ArgError("world > 1 but neither an all-reduce callback nor an NCCL communicator is set");
      }
    }
  }
  void run();
};

constexpr int LIN_NB_DEFAULT = 3;
static int grid_for(long long n, int threads) { return (int)std::max<long long>(1, (n + threads - 1) / threads); }

void BA::run() {
  OSFM_CUDA(cudaSetDevice(device));
  const auto t_start = std::chrono::high_resolution_clock::now();
  // OSFM_BA_TRACE=1: host wall-clock per phase of run() on stderr (diagnostics only)
  static const bool trace_on = []() { const char* e = getenv("OSFM_BA_TRACE"); return e && e[0] == '1'; }();
  auto t_prev = t_start;
  ar_trace = trace_on; ar_calls = 0; ar_host_ms = 0.0; ar_dev_ms = 0.0;
  auto trace = [&](const char* what) {
    if (!trace_on) return;
    const auto now = std::chrono::high_resolution_clock::now();
    fprintf(stderr, "[osfm_ba] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  const int K = (int)cam_type.size(), NI = (int)inst_const.size(), NR = (int)rc_const.size();
  const int S = (int)shot_inst.size(), Pfull = (int)pt_const.size();
  const long long Nfull = n_obs_full;
  if (K == 0 && Nfull > 0) throw ArgError("observations but no cameras");
  int64_t launches0 = g_kernel_launches.load();

  // ---- validation (errors mirror the reference's: missing ids -> runtime_error) ----
  for (int s = 0; s < S; ++s) {
    if (shot_inst[s] < 0 || shot_inst[s] >= NI) throw ArgError("shot references a rig instance that doesn't exist");
    if (shot_cam[s] < 0 || shot_cam[s] >= K) throw ArgError("shot references a camera that doesn't exist");
    if (shot_use_rc[s] && (shot_rc[s] < 0 || shot_rc[s] >= NR))
      throw ArgError("shot references a rig camera that doesn't exist");
  }
  // (observation indices are checked on the device, ord_make_keys)

  trace("validate");
  // ---- layout of the reduced vector: [free cameras | free instances | free rig cameras] ----
  std::vector<int> cam_off(K + 1, 0), cam_np(K), cam_poff(K), inst_poff(NI), rc_poff(std::max(NR, 1), -1);
  std::vector<int> blk_off, blk_sz;
  int off = 0;
  for (int k = 0; k < K; ++k) {
    cam_np[k] = model_num_params(cam_type[k]);
    cam_off[k + 1] = cam_off[k] + cam_np[k];
  }
  for (int k = 0; k < K; ++k) {
    cam_poff[k] = cam_const[k] ? -1 : off;
    if (!cam_const[k]) { blk_off.push_back(off); blk_sz.push_back(cam_np[k]); off += cam_np[k]; }
  }
  for (int i = 0; i < NI; ++i) {
    inst_poff[i] = inst_const[i] ? -1 : off;
    if (!inst_const[i]) { blk_off.push_back(off); blk_sz.push_back(6); off += 6; }
  }
  for (int i = 0; i < NR; ++i) {
    rc_poff[i] = rc_const[i] ? -1 : off;
    if (!rc_const[i]) { blk_off.push_back(off); blk_sz.push_back(6); off += 6; }
  }
  const int NE = (int)ext_size.size();
  std::vector<int> ext_off(NE + 1, 0), ext_poff(std::max(NE, 1), -1);
  for (int i = 0; i < NE; ++i) {
    if (ext_size[i] < 1 || ext_size[i] > MAXB) throw ArgError("ext block size must be in [1, 16]");
    ext_off[i + 1] = ext_off[i] + ext_size[i];
    ext_poff[i] = ext_const[i] ? -1 : off;
    if (!ext_const[i]) { blk_off.push_back(off); blk_sz.push_back(ext_size[i]); off += ext_size[i]; }
  }
  const int nc = off;
  const int nblk = (int)blk_off.size();
  const int nc_pad = (nc + 15) / 16 * 16;  // rhs sits in front of the reduced system in one buffer
  // parameter-block id of every camera / rig instance / rig camera / ext block (-1 = constant)
  std::vector<int> cam_blk(std::max(K, 1), -1), inst_blk(std::max(NI, 1), -1), rc_blk(std::max(NR, 1), -1),
      ext_blk(std::max(NE, 1), -1);
  {
    int b = 0;
    for (int k = 0; k < K; ++k) if (!cam_const[k]) cam_blk[k] = b++;
    for (int i = 0; i < NI; ++i) if (!inst_const[i]) inst_blk[i] = b++;
    for (int i = 0; i < NR; ++i) if (!rc_const[i]) rc_blk[i] = b++;
    for (int i = 0; i < NE; ++i) if (!ext_const[i]) ext_blk[i] = b++;
  }
  // side terms: block references checked here (the reference's std::map::at / "doesn't exist" errors)
  bool constrained = false;   // a free ext parameter with a finite lower bound (ceres: Problem::IsConstrained)
  for (int i = 0; i < NE; ++i)
    for (int j = 0; j < ext_size[i] && !ext_const[i]; ++j) constrained |= std::isfinite(ext_lower[ext_off[i] + j]);
  const int NT = (int)side_terms.size();
  std::vector<int> side_jofs(NT + 1, 0), side_rofs(NT + 1, 0);
  for (int t = 0; t < NT; ++t) {
    const osfm_side_term& st = side_terms[t];
    if (st.type < 0 || st.type >= OSFM_SIDE_NUM_TYPES) throw ArgError("unknown side term type");
    if (st.nblocks < 1 || st.nblocks > SIDE_MAX_BLOCKS || st.nres < 1 || st.nres > SIDE_MAX_RES)
      throw ArgError("side term with a bad block / residual count");
    int np = 0;
    for (int b = 0; b < st.nblocks; ++b) {
      const int kd = st.kind[b], ix = st.idx[b];
      const int cnt = kd == SB_CAM ? K : kd == SB_INST ? NI : kd == SB_RIGCAM ? NR : kd == SB_EXT ? NE : -1;
      if (ix < 0 || ix >= cnt) throw ArgError("side term references a parameter block that doesn't exist");
      np += kd == SB_CAM ? model_num_params(cam_type[ix]) : kd == SB_EXT ? ext_size[ix] : 6;
    }
    if (np > SIDE_MAX_PARAMS) throw ArgError("side term with too many parameters");
    if (st.cofs < 0 || (size_t)st.cofs > side_consts.size()) throw ArgError("side term constants out of range");
    side_jofs[t + 1] = side_jofs[t] + st.nres * np;
    side_rofs[t + 1] = side_rofs[t] + st.nres;
  }

  // preconditioner groups: a camera and the rig instance that is its only user (and vice versa) are
  // merged into one diagonal block when they fit (C + 6 <= 16); everything else stays on its own
  std::vector<int> grp_b1, grp_b2;
  {
    std::vector<int> cam_user(std::max(K, 1), -1), inst_cam(std::max(NI, 1), -1);  // -1 none, -2 several
    for (int s2 = 0; s2 < S; ++s2) {
      const int k = shot_cam[s2], i = shot_inst[s2];
      cam_user[k] = cam_user[k] == -1 || cam_user[k] == i ? i : -2;
      inst_cam[i] = inst_cam[i] == -1 || inst_cam[i] == k ? k : -2;
    }
    std::vector<char> inst_done(std::max(NI, 1), 0);
    for (int k = 0; k < K; ++k) {
      if (cam_blk[k] < 0) continue;
      const int i = cam_user[k];
      if (i >= 0 && inst_blk[i] >= 0 && inst_cam[i] == k && cam_np[k] + 6 <= 16) {
        grp_b1.push_back(cam_blk[k]); grp_b2.push_back(inst_blk[i]); inst_done[i] = 1;
      } else {
        grp_b1.push_back(cam_blk[k]); grp_b2.push_back(-1);
      }
    }
    for (int i = 0; i < NI; ++i)
      if (inst_blk[i] >= 0 && !inst_done[i]) { grp_b1.push_back(inst_blk[i]); grp_b2.push_back(-1); }
    for (int i = 0; i < NR; ++i)
      if (rc_blk[i] >= 0) { grp_b1.push_back(rc_blk[i]); grp_b2.push_back(-1); }
    for (int i = 0; i < NE; ++i)
      if (ext_blk[i] >= 0) { grp_b1.push_back(ext_blk[i]); grp_b2.push_back(-1); }
  }
  const int ngroups = (int)grp_b1.size();

  int wc = 0, nres = 2;
  for (int s = 0; s < S; ++s) {
    wc = std::max(wc, cam_np[shot_cam[s]] + 6 + (shot_use_rc[s] ? 6 : 0));
    if (cam_type[shot_cam[s]] == PT_SPHERICAL) nres = 3;
  }
  wc = std::max(wc, 1);

  trace("layout");
  // ---- order the observations on the device (ba_order.cuh): shard points over ranks
  //      (p % world == rank), sort by (point, shot), put points seen by exactly the same shots next to
  //      each other (segments of the fast Schur path) ----
  // The segmented Schur path (ba_point_blocks + ba_obs_rows + ba_schur_seg) is the default: 3.6 ms vs 5.3 ms
  // per launch for the per-point kernel on the 2M-observation scene (profiles/README.md).
  // OSFM_BA_SEGMENT_SCHUR=0 forces every point through ba_schur (kept for A/B runs and tests).
  static const bool use_seg = []() { const char* e = getenv("OSFM_BA_SEGMENT_SCHUR"); return !(e && e[0] == '0'); }();
  // register budget of ba_linearize<1> (A/B switch): 3, 4 or 5 resident CTAs per SM
  static const int lin_nb = []() { const char* e = getenv("OSFM_BA_LIN_NB"); return e ? atoi(e) : LIN_NB_DEFAULT; }();
  // one projection type for all cameras and no rig-camera shots -> specialised linearisation kernels
  static const bool lin_special = []() { const char* e = getenv("OSFM_BA_LIN_SPECIAL"); return !(e && e[0] == '0'); }();
  int uniform_type = -1;
  if (lin_special && K > 0) {
    uniform_type = cam_type[0];
    for (int k = 1; k < K; ++k) if (cam_type[k] != uniform_type) uniform_type = -1;
    for (int s = 0; s < S && uniform_type >= 0; ++s) if (shot_use_rc[s]) uniform_type = -1;
    if (uniform_type != PT_PERSPECTIVE && uniform_type != PT_BROWN && uniform_type != PT_FISHEYE) uniform_type = -1;
  }
  if (Nfull >= (1LL << 31)) throw ArgError("too many observations");
  const int P = Pfull > rank ? (Pfull - rank + world - 1) / world : 0;
  const size_t Nfz = (size_t)std::max<long long>(Nfull, 1), Pz = (size_t)std::max(P, 1);
  const bool have_pp = !pp_point.empty();
  {
    std::vector<int> ptc = pt_const;
    for (int& c : ptc) c = c ? 1 : 0;
    for (int q : pp_point) {
      if (q < 0 || q >= Pfull) throw ArgError("point prior on a point that doesn't exist");
      ptc[q] |= 2;
    }
    upload(d_ptc_full, ptc, stream);
    OSFM_CUDA(cudaStreamSynchronize(stream));  // ptc goes out of scope
  }
  upload(d_pts_in, pts, stream);
  d_okeys.reserve(Nfz); d_okeys2.reserve(Nfz); d_ovals.reserve(Nfz); d_ovals2.reserve(Nfz);
  d_g_pt_start.reserve((size_t)Pfull + 1);
  d_pkey.reserve(Pz); d_pkey2.reserve(Pz); d_pval.reserve(Pz); d_order.reserve(Pz); d_inv_order.reserve(Pz);
  d_global_of.reserve(Pz); d_free_flag.reserve(Pz + 1); d_free_scan.reserve(Pz + 1); d_kk.reserve(Pz + 1);
  d_pt_start.reserve(Pz + 1); d_pt_poff.reserve(Pz); d_head.reserve(Pz); d_run_head.reserve(Pz);
  d_seg_flags.reserve(Pz); d_seg_start.reserve(Pz + 1); d_nsel.reserve(1); d_oc.reserve(1); h_oc.reserve(1);
  d_pts[0].reserve(3 * Pz); d_pts[1].reserve(3 * Pz);
  int pbits = 1;
  while ((1LL << pbits) <= (long long)Pfull) ++pbits;
  {
    size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, d_okeys.p, d_okeys2.p, d_ovals.p, d_ovals2.p, (int)Nfull, 0, 32 + pbits, stream);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, d_pkey.p, d_pkey2.p, d_pval.p, d_order.p, P, 0, 64, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t3, d_kk.p, d_pt_start.p, P + 1, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, t4, d_free_flag.p, d_free_scan.p, P + 1, stream);
    cub::DeviceScan::InclusiveScan(nullptr, t5, d_head.p, d_run_head.p, OrdMax(), P, stream);
    cub::DeviceSelect::Flagged(nullptr, t6, thrust::counting_iterator<int>(0), d_seg_flags.p, d_seg_start.p, d_nsel.p, P, stream);
    d_cub.reserve(std::max({t1, t2, t3, t4, t5, t6}) + 256);
  }
  OSFM_CUDA(cudaMemsetAsync(d_oc.p, 0, sizeof(OrderCounts), stream));
  ord_make_keys<<<grid_for(Nfull, 256), 256, 0, stream>>>(Nfull, d_raw_shot.p, d_raw_point.p, S, Pfull, d_okeys.p, d_ovals.p, d_oc.p);
  OSFM_LAUNCH_CHECK();
  size_t tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceRadixSort::SortPairs(d_cub.p, tmpb, d_okeys.p, d_okeys2.p, d_ovals.p, d_ovals2.p, (int)Nfull, 0, 32 + pbits, stream));
  const unsigned long long* okeys = d_okeys2.p;  // sorted keys / positions in the caller's list
  const int* ovals = d_ovals2.p;
  ord_point_starts<<<grid_for(Pfull + 1, 256), 256, 0, stream>>>(okeys, Nfull, Pfull, d_g_pt_start.p);
  OSFM_LAUNCH_CHECK();
  ord_pair_bound<<<grid_for(Pfull, 256), 256, 0, stream>>>(d_g_pt_start.p, Pfull, d_oc.p);
  OSFM_LAUNCH_CHECK();
  ord_signatures<<<grid_for(P, 256), 256, 0, stream>>>(okeys, d_g_pt_start.p, d_ptc_full.p, P, world, rank, wc, use_seg ? 1 : 0,
                                                       SEG_KMAX, SEG_NA, SEG_WCMAX, d_pkey.p, d_pval.p);
  OSFM_LAUNCH_CHECK();
  tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceRadixSort::SortPairs(d_cub.p, tmpb, d_pkey.p, d_pkey2.p, d_pval.p, d_order.p, P, 0, 64, stream));
  ord_counts<<<grid_for(P + 1, 256), 256, 0, stream>>>(d_order.p, d_g_pt_start.p, d_ptc_full.p, P, world, rank, d_kk.p,
                                                       d_free_flag.p, d_inv_order.p, d_global_of.p);
  OSFM_LAUNCH_CHECK();
  tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tmpb, d_kk.p, d_pt_start.p, P + 1, stream));
  tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tmpb, d_free_flag.p, d_free_scan.p, P + 1, stream));
  ord_finish_points<<<grid_for(P + 1, 256), 256, 0, stream>>>(P, d_free_flag.p, d_free_scan.p, d_pt_start.p, d_global_of.p,
                                                              d_pts_in.p, d_pt_poff.p, d_pts[0].p, d_pts[1].p, d_oc.p);
  OSFM_LAUNCH_CHECK();
  OSFM_CUDA(cudaMemcpyAsync(h_oc.p, d_oc.p, sizeof(OrderCounts), cudaMemcpyDeviceToHost, stream));
  OSFM_CUDA(cudaStreamSynchronize(stream));
  if (h_oc.p->err & 1) throw ArgError("observation references a shot that doesn't exist");
  if (h_oc.p->err & 2) throw ArgError("observation references a point that doesn't exist");
  const long long N = h_oc.p->n_local;
  const int npf = h_oc.p->npf;
  const int n = nc + 3 * npf;
  {
    const size_t Nz0 = (size_t)std::max<long long>(N, 1);
    d_obs_orig.reserve(Nz0); d_obs_shot.reserve(Nz0); d_obs_point.reserve(Nz0);
    d_obs_x.reserve(Nz0); d_obs_y.reserve(Nz0); d_obs_isig.reserve(Nz0);
  }
  if (obs_pending) {   // image coordinates / standard deviations uploaded by osfm_ba_set_observations_async
    OSFM_CUDA(cudaStreamWaitEvent(stream, ev_obs, 0));
    obs_pending = false;
  }
  ord_gather_obs<<<grid_for(Nfull, 256), 256, 0, stream>>>(okeys, ovals, Nfull, d_g_pt_start.p, d_inv_order.p, d_pt_start.p, world,
                                                           rank, d_raw_xy.p, d_raw_sigma.p, d_obs_orig.p, d_obs_shot.p,
                                                           d_obs_point.p, d_obs_x.p, d_obs_y.p, d_obs_isig.p);
  OSFM_LAUNCH_CHECK();
  ord_seg_heads<<<grid_for(P, 256), 256, 0, stream>>>(P, d_pkey2.p, d_order.p, d_g_pt_start.p, okeys, d_ptc_full.p, world, rank,
                                                      d_head.p);
  OSFM_LAUNCH_CHECK();
  tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceScan::InclusiveScan(d_cub.p, tmpb, d_head.p, d_run_head.p, OrdMax(), P, stream));
  ord_seg_flags<<<grid_for(P, 256), 256, 0, stream>>>(P, d_pkey2.p, d_run_head.p, d_seg_flags.p);
  OSFM_LAUNCH_CHECK();
  tmpb = d_cub.cap;
  OSFM_CUDA(cub::DeviceSelect::Flagged(d_cub.p, tmpb, thrust::counting_iterator<int>(0), d_seg_flags.p, d_seg_start.p, d_nsel.p, P,
                                       stream));
  ord_seg_finish<<<1, 32, 0, stream>>>(P, d_pkey2.p, d_pt_start.p, d_seg_start.p, d_nsel.p, d_oc.p);
  OSFM_LAUNCH_CHECK();
  if (world > 1) {  // the structure of the reduced system needs every point's shots on every rank
    d_g_obs_shot.reserve(Nfz); d_g_obs_point.reserve(Nfz);
    ord_split_keys<<<grid_for(Nfull, 256), 256, 0, stream>>>(okeys, Nfull, d_g_obs_shot.p, d_g_obs_point.p);
    OSFM_LAUNCH_CHECK();
  }
  OSFM_CUDA(cudaMemcpyAsync(h_oc.p, d_oc.p, sizeof(OrderCounts), cudaMemcpyDeviceToHost, stream));
  OSFM_CUDA(cudaStreamSynchronize(stream));
  const int nseg = h_oc.p->nseg, P_fast = h_oc.p->p_fast;
  const long long n_fast_obs = h_oc.p->n_fast;
  const long long pair_bound_all = (long long)h_oc.p->pair_bound;
  if (trace_on)
    fprintf(stderr, "[osfm_ba] points %d (fast path %d in %d segments), observations %lld (fast path %lld), free points %d\n", P,
            P_fast, nseg, N, n_fast_obs, npf);

  trace("sort");
  // ---- prior rows (rank 0 adds them; Ceres drops residuals of constant blocks) ----
  std::vector<int> pr_cam_param, pr_cam_col, pr_cam_log, pr_pos_kind, pr_pos_inst, pr_pos_axis, pr_pos_col;
  std::vector<double> pr_cam_prior, pr_cam_scale, pr_pos_prior, pr_pos_scale;
  for (int k = 0; k < K; ++k) {
    if (cam_poff[k] < 0) continue;
    for (int j = 0; j < cam_np[k]; ++j) {
      const int idx = cam_off[k] + j;
      pr_cam_param.push_back(idx);
      pr_cam_col.push_back(cam_poff[k] + j);
      pr_cam_log.push_back(cam_prior_log[idx]);
      pr_cam_prior.push_back(cam_prior[idx]);
      pr_cam_scale.push_back(1.0 / std::max(cam_prior_sigma[idx], DBL_EPSILON));  // prior_error.h:31-35
    }
  }
  for (int i = 0; i < NI; ++i) {
    if (!inst_has_prior[i] || inst_poff[i] < 0) continue;
    for (int j = 0; j < 3; ++j) {
      pr_pos_kind.push_back(1); pr_pos_inst.push_back(i); pr_pos_axis.push_back(3 + j); pr_pos_col.push_back(inst_poff[i] + 3 + j);
      pr_pos_prior.push_back(inst_prior_pos[3 * (size_t)i + j]);
      pr_pos_scale.push_back(1.0 / std::max(inst_prior_std[3 * (size_t)i + j], DBL_EPSILON));
    }
  }
  const bool have_rc_prior = !rc_prior.empty();
  for (int i = 0; i < NR && have_rc_prior; ++i) {   // DataPriorError<Pose> on every free rig camera (:779-790)
    if (rc_poff[i] < 0) continue;
    for (int j = 0; j < 6; ++j) {
      pr_pos_kind.push_back(2); pr_pos_inst.push_back(i); pr_pos_axis.push_back(j); pr_pos_col.push_back(rc_poff[i] + j);
      pr_pos_prior.push_back(rc_prior[6 * (size_t)i + j]);
      pr_pos_scale.push_back(1.0 / std::max(rc_prior_sigma[6 * (size_t)i + j], DBL_EPSILON));
    }
  }
  const int npr = (int)(pr_cam_param.size() + pr_pos_inst.size());
  const bool add_priors = rank == 0;
  std::vector<int> pr_blk, pr_local;  // diagonal entry of every prior row inside its diagonal block
  for (int k = 0; k < K; ++k) {
    if (cam_poff[k] < 0) continue;
    for (int j = 0; j < cam_np[k]; ++j) { pr_blk.push_back(cam_blk[k]); pr_local.push_back(j); }
  }
  for (int i = 0; i < NI; ++i) {
    if (!inst_has_prior[i] || inst_poff[i] < 0) continue;
    for (int j = 0; j < 3; ++j) { pr_blk.push_back(inst_blk[i]); pr_local.push_back(3 + j); }
  }
  for (int i = 0; i < NR && have_rc_prior; ++i) {
    if (rc_poff[i] < 0) continue;
    for (int j = 0; j < 6; ++j) { pr_blk.push_back(rc_blk[i]); pr_local.push_back(j); }
  }

  trace("priors");
  // ---- upload ----
  upload(d_cam_type, cam_type, stream); upload(d_cam_off, cam_off, stream); upload(d_cam_np, cam_np, stream);
  upload(d_cam_poff, cam_poff, stream); upload(d_inst_poff, inst_poff, stream); upload(d_rc_poff, rc_poff, stream);
  upload(d_shot_inst, shot_inst, stream); upload(d_shot_cam, shot_cam, stream); upload(d_shot_rc, shot_rc, stream);
  upload(d_shot_use_rc, shot_use_rc, stream);
  std::vector<double> rc_h = rc;
  if (rc_h.empty()) rc_h.assign(6, 0.0);
  for (int b = 0; b < 2; ++b) {
    upload(d_cam[b], cam_params, stream); upload(d_inst[b], inst, stream); upload(d_rc[b], rc_h, stream);
  }
  upload(d_blk_off, blk_off, stream); upload(d_blk_sz, blk_sz, stream);
  upload(d_cam_blk, cam_blk, stream); upload(d_inst_blk, inst_blk, stream); upload(d_rc_blk, rc_blk, stream);
  upload(d_pr_blk, pr_blk, stream); upload(d_pr_local, pr_local, stream);
  upload(d_grp_b1, grp_b1, stream); upload(d_grp_b2, grp_b2, stream);
  upload(d_pr_cam_param, pr_cam_param, stream); upload(d_pr_cam_col, pr_cam_col, stream);
  upload(d_pr_cam_log, pr_cam_log, stream); upload(d_pr_cam_prior, pr_cam_prior, stream);
  upload(d_pr_cam_scale, pr_cam_scale, stream); upload(d_pr_pos_inst, pr_pos_inst, stream);
  upload(d_pr_pos_axis, pr_pos_axis, stream); upload(d_pr_pos_col, pr_pos_col, stream);
  upload(d_pr_pos_kind, pr_pos_kind, stream);
  // ext blocks, side terms, point priors
  {
    std::vector<double> ev = ext_values, el = ext_lower;
    if (ev.empty()) { ev.assign(1, 0.0); el.assign(1, 0.0); }
    upload(d_ext[0], ev, stream); upload(d_ext[1], ev, stream); upload(d_ext_lower, el, stream);
    upload(d_ext_off, ext_off, stream); upload(d_ext_np, ext_size, stream); upload(d_ext_poff, ext_poff, stream);
    upload(d_ext_blk, ext_blk, stream);
    upload(d_side_terms, side_terms, stream); upload(d_side_consts, side_consts, stream);
    upload(d_side_jofs, side_jofs, stream); upload(d_side_rofs, side_rofs, stream);
    d_side_J.reserve((size_t)side_jofs[NT] + 1); d_side_r.reserve((size_t)side_rofs[NT] + 1);
    if (have_pp) {
      std::vector<double> ppd(3 * (size_t)Pfull, 0.0), ppx(3 * (size_t)Pfull, 0.0);
      for (size_t q = 0; q < pp_point.size(); ++q) {
        const size_t g = (size_t)pp_point[q];
        for (int j = 0; j < (pp_alt[q] ? 3 : 2); ++j) {
          ppd[3 * g + j] = 1.0 / std::max(pp_sigma[3 * q + j], DBL_EPSILON);   // prior_error.h:31-35
          ppx[3 * g + j] = pp_prior[3 * q + j];
        }
      }
      upload(d_pp_d, ppd, stream); upload(d_pp_x0, ppx, stream);
    }
    OSFM_CUDA(cudaStreamSynchronize(stream));  // local vectors go out of scope
  }
  upload(d_pr_pos_prior, pr_pos_prior, stream); upload(d_pr_pos_scale, pr_pos_scale, stream);
  const size_t Nz = (size_t)std::max<long long>(N, 1);
  d_r.reserve(nres * Nz); d_Jc.reserve((size_t)nres * wc * Nz); d_Jp.reserve((size_t)nres * 3 * Nz);
  d_Vinv.reserve(6 * (size_t)std::max(npf, 1)); d_gp.reserve(3 * (size_t)std::max(npf, 1));
  const size_t nz = (size_t)std::max(n, 1);
  d_scale.reserve(nz); d_colnorm2.reserve(nz); d_grad.reserve(nz); d_diag.reserve(nz); d_y.reserve(nz);
  d_px.reserve(std::max(nc, 1)); d_pr.reserve(std::max(nc, 1)); d_pz.reserve(std::max(nc, 1));
  d_pp.reserve(std::max(nc, 1)); d_pAp.reserve(std::max(nc, 1));
  d_pcg.reserve(1);
  d_Minv.reserve((size_t)std::max(ngroups, 1) * MAXB * MAXB);
  d_Ap.reserve(std::max(nc, 1));
  d_sc.reserve(1);

  BAView v{};
  v.K = K; v.NI = NI; v.NR = NR; v.S = S; v.P = P; v.N = N; v.nc = nc; v.npf = npf; v.wc = wc; v.nres = nres;
  v.loss = loss; v.loss_a = loss_a;
  v.cam_type = d_cam_type.p; v.cam_off = d_cam_off.p; v.cam_np = d_cam_np.p; v.cam_poff = d_cam_poff.p;
  v.inst_poff = d_inst_poff.p; v.rc_poff = d_rc_poff.p; v.pt_poff = d_pt_poff.p;
  v.shot_inst = d_shot_inst.p; v.shot_cam = d_shot_cam.p; v.shot_rc = d_shot_rc.p; v.shot_use_rc = d_shot_use_rc.p;
  v.obs_shot = d_obs_shot.p; v.obs_point = d_obs_point.p; v.obs_x = d_obs_x.p; v.obs_y = d_obs_y.p;
  v.obs_isig = d_obs_isig.p; v.obs_orig = d_obs_orig.p; v.pt_start = d_pt_start.p;
  v.r = d_r.p; v.Jc = d_Jc.p; v.Jp = d_Jp.p;
  PriorView pv{};
  pv.n_cam_rows = add_priors ? (int)pr_cam_param.size() : 0;
  pv.n_pos_rows = add_priors ? (int)pr_pos_inst.size() : 0;
  pv.cam_row_param = d_pr_cam_param.p; pv.cam_row_col = d_pr_cam_col.p; pv.cam_row_log = d_pr_cam_log.p;
  pv.cam_row_prior = d_pr_cam_prior.p; pv.cam_row_scale = d_pr_cam_scale.p;
  pv.pos_row_kind = d_pr_pos_kind.p;
  pv.pos_row_inst = d_pr_pos_inst.p; pv.pos_row_axis = d_pr_pos_axis.p; pv.pos_row_col = d_pr_pos_col.p;
  pv.pos_row_prior = d_pr_pos_prior.p; pv.pos_row_scale = d_pr_pos_scale.p;
  const int npr_local = pv.n_cam_rows + pv.n_pos_rows;
  (void)npr;

  int cur = 0;  // index of the accepted parameter set
  auto params_of = [&](int b) { return Params{d_cam[b].p, d_inst[b].p, d_rc[b].p, d_pts[b].p, d_ext[b].p}; };
  SideView sv{};
  sv.n = NT; sv.terms = d_side_terms.p; sv.consts = d_side_consts.p; sv.jofs = d_side_jofs.p; sv.rofs = d_side_rofs.p;
  sv.J = d_side_J.p; sv.r = d_side_r.p;
  sv.ext_off = d_ext_off.p; sv.ext_np = d_ext_np.p; sv.ext_poff = d_ext_poff.p; sv.ext_blk = d_ext_blk.p;
  PointPriorView ppv{have_pp ? d_pp_d.p : nullptr, d_pp_x0.p, d_global_of.p};
  BlkMaps bm{d_cam_blk.p, d_inst_blk.p, d_rc_blk.p};

  cudaEvent_t ev0, ev1;
  OSFM_CUDA(cudaEventCreate(&ev0)); OSFM_CUDA(cudaEventCreate(&ev1));
  trace("upload");
  EventTimer tm_lin, tm_schur, tm_pcg, tm_back;
  tm_lin.init(); tm_schur.init(); tm_pcg.init(); tm_back.init();

  // cost at parameter set b (sum over ranks)
  auto eval_cost = [&](int b) -> double {
    OSFM_CUDA(cudaMemsetAsync(&d_sc.p->cost, 0, sizeof(double), stream));
    if (N > 0) {
      if (uniform_type == PT_PERSPECTIVE) ba_linearize<0, 3, PT_PERSPECTIVE><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (uniform_type == PT_BROWN) ba_linearize<0, 3, PT_BROWN><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (uniform_type == PT_FISHEYE) ba_linearize<0, 3, PT_FISHEYE><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else ba_linearize<0><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      OSFM_LAUNCH_CHECK();
    }
    if (npr_local > 0) {
      ba_prior_cost<<<grid_for(npr_local, 128), 128, 0, stream>>>(pv, params_of(b), d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (NT > 0 && add_priors) {
      side_cost<<<grid_for(NT, 128), 128, 0, stream>>>(sv, v, bm, params_of(b), d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (have_pp && P > 0) {
      ba_point_prior<0><<<grid_for(P, 128), 128, 0, stream>>>(ppv, v, params_of(b), nullptr, nullptr, nullptr, nullptr, nullptr, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    allreduce_dev(&d_sc.p->cost, 1);
    return read_scalars().cost;
  };
  // residual + Jacobian planes, column norms, gradient at parameter set b
  auto linearize = [&](int b, double* grad_max) -> double {
    OSFM_CUDA(cudaMemsetAsync(d_sc.p, 0, sizeof(Scalars), stream));
    OSFM_CUDA(cudaMemsetAsync(d_colnorm2.p, 0, sizeof(double) * nz, stream));
    OSFM_CUDA(cudaMemsetAsync(d_grad.p, 0, sizeof(double) * nz, stream));
    if (N > 0) {
      tm_lin.start(stream);
      if (uniform_type == PT_PERSPECTIVE) ba_linearize<1, 5, PT_PERSPECTIVE><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (uniform_type == PT_BROWN) ba_linearize<1, 4, PT_BROWN><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (uniform_type == PT_FISHEYE) ba_linearize<1, 5, PT_FISHEYE><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (lin_nb == 5) ba_linearize<1, 5><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else if (lin_nb == 4) ba_linearize<1, 4><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      else ba_linearize<1, 3><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(b), d_sc.p, nullptr);
      OSFM_LAUNCH_CHECK();
      tm_lin.stop(stream);
      ba_colnorm_grad_points<<<grid_for(N, 256), 256, 0, stream>>>(v, d_colnorm2.p, d_grad.p);
      OSFM_LAUNCH_CHECK();
      if (nseg > 0) {
        // OSFM_BA_COLNORM_TMA=0: the shuffle kernel instead of the bulk-copy staged one (A/B switch)
        static const bool use_tma = []() { const char* e = getenv("OSFM_BA_COLNORM_TMA"); return !(e && e[0] == '0'); }();
        if (wc == 9 && nres == 2 && use_tma && have_seg_tab && sp_nchunks > 0) {
          // chunk list + segment tables exist (every call but the first of a run): no dependent index loads
          if (!cc_attr) {
            OSFM_CUDA(cudaFuncSetAttribute(ba_colnorm_grad_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, CC_SMEM));
            cc_attr = true;
          }
          const int grid = std::max(1, std::min(num_sms, (sp_nchunks + CC_WARPS - 1) / CC_WARPS));
          ba_colnorm_grad_chunks<<<grid, 32 * CC_WARPS, CC_SMEM, stream>>>(v, d_sp_chunks.p, sp_nchunks, d_tab.p, d_colnorm2.p, d_grad.p);
        } else if (wc == 9 && nres == 2 && use_tma) {
          if (!cg_attr) {
            OSFM_CUDA(cudaFuncSetAttribute(ba_colnorm_grad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, CG_SMEM));
            cg_attr = true;
          }
          const int grid = std::max(1, std::min(num_sms, (nseg + CG_WARPS - 1) / CG_WARPS));
          ba_colnorm_grad_tma<<<grid, 32 * CG_WARPS, CG_SMEM, stream>>>(v, d_seg_start.p, nseg, have_seg_tab ? d_tab_off.p : nullptr,
                                                                        have_seg_tab ? d_tab.p : nullptr, d_colnorm2.p, d_grad.p);
        } else if (wc == 9)
          ba_colnorm_grad_seg<9><<<grid_for((long long)nseg * 32, 256), 256, 0, stream>>>(v, d_seg_start.p, nseg, d_colnorm2.p, d_grad.p);
        else
          ba_colnorm_grad_seg<0><<<grid_for((long long)nseg * 32, 256), 256, 0, stream>>>(v, d_seg_start.p, nseg, d_colnorm2.p, d_grad.p);
        OSFM_LAUNCH_CHECK();
      }
      if (N > n_fast_obs) {
        ba_colnorm_grad<<<grid_for(N - n_fast_obs, 256), 256, 0, stream>>>(v, n_fast_obs, d_colnorm2.p, d_grad.p);
        OSFM_LAUNCH_CHECK();
      }
    }
    if (npr_local > 0) {
      ba_prior_cost<<<grid_for(npr_local, 128), 128, 0, stream>>>(pv, params_of(b), d_sc.p);
      OSFM_LAUNCH_CHECK();
      ba_prior_colnorm_grad<<<grid_for(npr_local, 128), 128, 0, stream>>>(pv, params_of(b), d_colnorm2.p, d_grad.p);
      OSFM_LAUNCH_CHECK();
    }
    if (NT > 0) {   // every rank keeps the terms' Jacobians (the system part is added after the all-reduce)
      side_linearize<<<NT, SIDE_THREADS, 0, stream>>>(sv, v, bm, params_of(b), d_sc.p, add_priors ? 1 : 0);
      OSFM_LAUNCH_CHECK();
      if (add_priors) {
        side_colnorm_grad<<<NT, SIDE_THREADS, 0, stream>>>(sv, v, bm, params_of(b), d_colnorm2.p, d_grad.p);
        OSFM_LAUNCH_CHECK();
      }
    }
    if (have_pp && P > 0) {
      ba_point_prior<1><<<grid_for(P, 128), 128, 0, stream>>>(ppv, v, params_of(b), nullptr, nullptr, d_colnorm2.p, d_grad.p, nullptr, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (world > 1) {
      // local max over the point part, then one all-reduce for cost, camera-side sums and the per-rank maxima
      if (n > nc) {
        ba_grad_max<<<grid_for(n - nc, 256), 256, 0, stream>>>(d_grad.p + nc, n - nc, d_sc.p);
        OSFM_LAUNCH_CHECK();
      }
      const int npack = 2 * nc + 1 + world;
      d_slots.reserve(npack);
      const int gsz = std::max(nc, world);
      ba_pack_lin<<<grid_for(gsz, 256), 256, 0, stream>>>(d_colnorm2.p, d_grad.p, d_sc.p, nc, rank, world, d_slots.p);
      OSFM_LAUNCH_CHECK();
      allreduce_dev(d_slots.p, npack);
      ba_unpack_lin<<<grid_for(gsz, 256), 256, 0, stream>>>(d_slots.p, nc, world, d_colnorm2.p, d_grad.p, d_sc.p);
      OSFM_LAUNCH_CHECK();
    } else if (n > 0) {
      ba_grad_max<<<grid_for(n, 256), 256, 0, stream>>>(d_grad.p, n, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    const Scalars s = read_scalars();
    const double gm = s.grad_max_bits;
    *grad_max = gm;
    return s.cost;
  };

  // persistent PCG grid: every CTA must be resident (1 CTA / SM) for the grid barrier
  const int pcg_grid = std::max(1, std::min(std::min(num_sms, PCG_MAX_CTAS), (nc + PCG_THREADS / 32 - 1) / (PCG_THREADS / 32)));
  trace("pre-struct");
  // ---- block-sparse structure of the reduced camera system (identical on every rank) ----
  BsrView bsr{};
  int n_upper = 0, n_blocks_all = 0;
  long long s_upper_total = 0, s_total = 0;
  if (nblk > 0) {
    // global CSR by point (every rank needs the same structure, not only its shard)
    const int* g_shot = d_obs_shot.p;
    const int* g_point = d_obs_point.p;
    const long long* g_start = d_pt_start.p;
    const long long pair_bound = pair_bound_all;
    if (world > 1) { g_shot = d_g_obs_shot.p; g_point = d_g_obs_point.p; g_start = d_g_pt_start.p; }
    const long long bound = std::min<long long>((long long)nblk * (nblk + 1) / 2, 9 * pair_bound + nblk + 21LL * NT);
    unsigned tsize = 1024;
    while ((long long)tsize < 4 * bound) {
      if (tsize >= (1u << 28)) throw std::runtime_error("reduced camera system has too many block pairs");
      tsize <<= 1;
    }
    d_tkeys.reserve(tsize); d_tvals.reserve(tsize);
    OSFM_CUDA(cudaMemsetAsync(d_tkeys.p, 0xff, sizeof(unsigned long long) * tsize, stream));
    BAView vg = v;  // only the shot tables are used by the enumeration
    const long long n_enum = world > 1 ? Nfull : N;
    if (n_enum > 0) {
      bsr_enum_pairs<<<grid_for(n_enum, 128), 128, 0, stream>>>(vg, bm, g_shot, g_start, g_point, n_enum, d_tkeys.p,
                                                              tsize - 1, nblk);
      OSFM_LAUNCH_CHECK();
    }
    if (NT > 0) {
      side_enum_pairs<<<grid_for(NT, 128), 128, 0, stream>>>(sv, v, bm, params_of(0), d_tkeys.p, tsize - 1, nblk);
      OSFM_LAUNCH_CHECK();
    }
    bsr_insert_diagonal<<<grid_for(nblk, 128), 128, 0, stream>>>(d_tkeys.p, tsize - 1, nblk);
    OSFM_LAUNCH_CHECK();
    const size_t cap = (size_t)(2 * bound + 16);
    d_skeys.reserve(cap); d_skeys2.reserve(cap); d_rkeys.reserve(cap); d_rkeys2.reserve(cap);
    d_area.reserve(cap); d_offs.reserve(cap); d_count.reserve(4);
    OSFM_CUDA(cudaMemsetAsync(d_count.p, 0, sizeof(unsigned), stream));
    bsr_compact<<<grid_for(tsize, 256), 256, 0, stream>>>(d_tkeys.p, tsize, nblk, d_skeys.p, d_count.p);
    OSFM_LAUNCH_CHECK();
    unsigned n_all_u = 0;
    OSFM_CUDA(cudaMemcpyAsync(&n_all_u, d_count.p, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaStreamSynchronize(stream));
    const int n_all = (int)n_all_u;
    n_blocks_all = n_all;
    n_upper = nblk + (n_all - nblk) / 2;
    // sort: upper keys (bit 63 clear) first, each group ordered by (bi, bj)
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp1, d_skeys.p, d_skeys2.p, n_all, 0, 64, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_area.p, d_offs.p, n_all, stream);
    d_cub.reserve(std::max(tmp1, tmp2) + 256);
    size_t tmp = d_cub.cap;
    OSFM_CUDA(cub::DeviceRadixSort::SortKeys(d_cub.p, tmp, d_skeys.p, d_skeys2.p, n_all, 0, 64, stream));
    g_kernel_launches.fetch_add(1);
    bsr_block_areas<<<grid_for(n_all, 256), 256, 0, stream>>>(d_skeys2.p, n_all, nblk, d_blk_sz.p, d_area.p);
    OSFM_LAUNCH_CHECK();
    tmp = d_cub.cap;
    OSFM_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tmp, d_area.p, d_offs.p, n_all, stream));
    g_kernel_launches.fetch_add(1);
    int last_off[2] = {0, 0}, last_area = 0, upper_end = 0;
    OSFM_CUDA(cudaMemcpyAsync(&last_off[0], d_offs.p + n_all - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaMemcpyAsync(&last_area, d_area.p + n_all - 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (n_upper < n_all)
      OSFM_CUDA(cudaMemcpyAsync(&upper_end, d_offs.p + n_upper, sizeof(int), cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaStreamSynchronize(stream));
    s_total = (long long)last_off[0] + last_area;
    s_upper_total = n_upper < n_all ? upper_end : s_total;
    // hash table values (inserting the mirrored keys) + plain key list, then block-row lists
    bsr_fill_table<<<grid_for(n_all, 256), 256, 0, stream>>>(d_skeys2.p, d_offs.p, n_all, d_tkeys.p, d_tvals.p, tsize - 1,
                                                            d_rkeys.p);
    OSFM_LAUNCH_CHECK();
    tmp = d_cub.cap;
    OSFM_CUDA(cub::DeviceRadixSort::SortKeys(d_cub.p, tmp, d_rkeys.p, d_rkeys2.p, n_all, 0, 64, stream));
    g_kernel_launches.fetch_add(1);
    bsr.tkeys = d_tkeys.p; bsr.tvals = d_tvals.p; bsr.tmask = tsize - 1; bsr.nblk = nblk;
    bsr.blk_off = d_blk_off.p; bsr.blk_sz = d_blk_sz.p;
    d_row_ptr.reserve(nblk + 2); d_row_col.reserve(n_all + 1); d_row_off.reserve(n_all + 1);
    bsr_rows<<<grid_for(n_all + 1, 256), 256, 0, stream>>>(d_rkeys2.p, n_all, bsr, d_row_ptr.p, d_row_col.p, d_row_off.p);
    OSFM_LAUNCH_CHECK();
    d_upper.reserve(n_upper + 1);
    bsr_upper_list<<<grid_for(n_upper, 256), 256, 0, stream>>>(d_skeys2.p, d_offs.p, n_upper, bsr, d_upper.p);
    OSFM_LAUNCH_CHECK();
    d_diag_off.reserve(nblk + 1);
    bsr_diag_offsets<<<grid_for(nblk, 128), 128, 0, stream>>>(bsr, d_diag_off.p);
    OSFM_LAUNCH_CHECK();
    d_prior_diag_off.reserve(pr_blk.size() + 1);
    if (!pr_blk.empty()) {
      bsr_prior_offsets<<<grid_for((long long)pr_blk.size(), 128), 128, 0, stream>>>(
          d_pr_blk.p, d_pr_local.p, (int)pr_blk.size(), d_diag_off.p, d_blk_sz.p, d_prior_diag_off.p);
      OSFM_LAUNCH_CHECK();
    }
    d_Sbuf.reserve((size_t)std::max<long long>(s_total, 1) + nc_pad);
    // block-row ELL layout of the PCG mat-vec
    d_row_M.reserve(nblk + 1); d_qoff.reserve(n_all + 1); d_blk_row.reserve(n_all + 1);
    pcg_row_sizes<<<grid_for(nblk, 128), 128, 0, stream>>>(d_row_ptr.p, d_row_col.p, bsr, d_row_M.p, d_qoff.p, d_blk_row.p);
    OSFM_LAUNCH_CHECK();
    std::vector<int> row_M(nblk);
    OSFM_CUDA(cudaMemcpyAsync(row_M.data(), d_row_M.p, sizeof(int) * nblk, cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> rowbase(nblk);
    std::vector<int> cbase(nblk);
    long long vb = 0, cbt = 0;
    for (int b = 0; b < nblk; ++b) {
      rowbase[b] = vb; cbase[b] = (int)cbt;
      vb += (long long)blk_sz[b] * row_M[b];
      cbt += row_M[b];
    }
    upload(d_rowbase, rowbase, stream); upload(d_cbase, cbase, stream);
    d_colidx.reserve((size_t)cbt + 1); d_row_of.reserve(nc + 1); d_Spcg.reserve((size_t)vb + 1);
    pcg_fill_colidx<<<grid_for(n_all, 128), 128, 0, stream>>>(d_row_col.p, d_qoff.p, d_blk_row.p, n_all, bsr, d_cbase.p,
                                                             d_colidx.p, d_row_of.p);
    OSFM_LAUNCH_CHECK();
    // resident PCG: contiguous scalar-row ranges per CTA, balanced by stored entries; usable when every CTA's
    // slice of S + its column indices (uint16) + p fit in shared memory
    {
      std::vector<long long> off(nc + 1, 0);
      std::vector<int> row_blk(nc, 0);
      bool monotone = true;
      for (int b = 0; b < nblk; ++b) {
        for (int r = 0; r < blk_sz[b]; ++r) {
          off[blk_off[b] + r] = rowbase[b] + (long long)r * row_M[b];
          row_blk[blk_off[b] + r] = b;
        }
        if (b > 0 && blk_off[b] != blk_off[b - 1] + blk_sz[b - 1]) monotone = false;
      }
      off[nc] = vb;
      const int G = pcg_grid;
      std::vector<int> row_lo(G + 1, nc);
      row_lo[0] = 0;
      for (int c = 1, i = 0; c < G; ++c) {
        const long long want = vb * c / G;
        while (i < nc && off[i] < want) ++i;
        row_lo[c] = i;
      }
      long long ent_max = 0, col_max = 0;
      int rows_max = 0;
      for (int c = 0; c < G; ++c) {
        const int lo = row_lo[c], hi = row_lo[c + 1];
        if (hi <= lo) continue;
        ent_max = std::max(ent_max, off[hi] - off[lo]);
        const int b_lo = row_blk[lo], b_hi = row_blk[hi - 1];
        col_max = std::max<long long>(col_max, (long long)cbase[b_hi] + row_M[b_hi] - cbase[b_lo]);
        rows_max = std::max(rows_max, hi - lo);
      }
      auto up16 = [](long long x) { return (x + 15) / 16 * 16; };
      const long long off_S = up16(8LL * nc), off_cols = off_S + up16(8 * ent_max), off_rows = off_cols + up16(2 * col_max);
      const long long total = off_rows + 12LL * rows_max;
      static const bool allow_res = []() { const char* e = getenv("OSFM_BA_PCG_RESIDENT"); return !(e && e[0] == '0'); }();
      int max_smem = 0;
      OSFM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
      pcg_resident = allow_res && monotone && nc <= 65535 && total + 1024 <= max_smem;
      pcg_smem = pcg_resident ? (int)total : 0;
      pcg_res = PcgResident{};
      if (pcg_resident) {
        upload(d_pcg_rowlo, row_lo, stream);
        pcg_res.row_lo = d_pcg_rowlo.p; pcg_res.off_S = (int)off_S; pcg_res.off_cols = (int)off_cols;
        pcg_res.off_rows = (int)off_rows; pcg_res.max_rows = rows_max;
        OSFM_CUDA(cudaFuncSetAttribute(pcg_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pcg_smem));
      }
    }
    // pipelined PCG: whole preconditioner groups per CTA, balanced by stored entries
    {
      const int G = pcg_grid;
      std::vector<long long> wsum(ngroups + 1, 0);
      auto gw_of = [&](int g, long long* cols, int* rows) {
        long long w = 0;
        for (int k = 0; k < 2; ++k) {
          const int b = k ? grp_b2[g] : grp_b1[g];
          if (b < 0) continue;
          w += (long long)blk_sz[b] * row_M[b];
          if (cols) *cols += row_M[b];
          if (rows) *rows += blk_sz[b];
        }
        return w;
      };
      for (int g = 0; g < ngroups; ++g) wsum[g + 1] = wsum[g] + gw_of(g, nullptr, nullptr);
      std::vector<int> grp_lo(G + 1, ngroups);
      grp_lo[0] = 0;
      for (int c = 1, g = 0; c < G; ++c) {
        const long long want = wsum[ngroups] * c / G;
        while (g < ngroups && wsum[g] < want) ++g;
        grp_lo[c] = g;
      }
      long long ent_max = 0, col_max = 0;
      int rows_max = 0, grp_max = 0;
      for (int c = 0; c < G; ++c) {
        long long cols = 0;
        int rows = 0;
        for (int g = grp_lo[c]; g < grp_lo[c + 1]; ++g) gw_of(g, &cols, &rows);
        ent_max = std::max(ent_max, wsum[grp_lo[c + 1]] - wsum[grp_lo[c]]);
        col_max = std::max(col_max, cols);
        rows_max = std::max(rows_max, rows);
        grp_max = std::max(grp_max, grp_lo[c + 1] - grp_lo[c]);
      }
      auto up16 = [](long long x) { return (x + 15) / 16 * 16; };
      const long long off_S = up16(8 * col_max), off_Minv = off_S + up16(8 * ent_max);
      const long long off_vec = off_Minv + 8LL * grp_max * MAXB * MAXB, off_cols = off_vec + up16(24LL * rows_max);
      const long long off_rows = off_cols + up16(2 * col_max);
      const long long off_defl = up16(off_rows + 36LL * rows_max + 4LL * (grp_max + 1));
      // own rows of the deflation vectors W and of S W, then the gather buffer of the wide barrier
      const long long total = off_defl + 2LL * PCG_ND * 8 * rows_max + 8LL * PCG_NW * G + 8LL * PCG_NW * rows_max;
      cudaFuncAttributes pipe_attr{};
      OSFM_CUDA(cudaFuncGetAttributes(&pipe_attr, pcg_pipelined));
      static const bool allow_pipe = []() { const char* e = getenv("OSFM_BA_PCG_PIPELINED"); return !(e && e[0] == '0'); }();
      int max_smem = 0;
      OSFM_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
      pcg_pipe_ok = allow_pipe && nc <= 65535 && rows_max <= PCG_THREADS && ent_max < (1LL << 30) &&
                    total + (long long)pipe_attr.sharedSizeBytes + 1024 <= max_smem;
      pcg_pipe_smem = pcg_pipe_ok ? (int)total : 0;
      pcg_pipe = PcgPipe{};
      if (pcg_pipe_ok) {
        upload(d_pcg_grplo, grp_lo, stream);
        pcg_pipe.grp_lo = d_pcg_grplo.p; pcg_pipe.off_S = (int)off_S; pcg_pipe.off_Minv = (int)off_Minv;
        pcg_pipe.off_vec = (int)off_vec; pcg_pipe.off_cols = (int)off_cols; pcg_pipe.off_rows = (int)off_rows;
        pcg_pipe.max_rows = rows_max; pcg_pipe.max_groups = grp_max; pcg_pipe.max_cols = (int)col_max;
        pcg_pipe.off_defl = (int)off_defl; pcg_pipe.Wdef = nullptr;
        // 128-bit barrier words (value + generation in one strong 16-byte access): measured SLOWER than flags + slots
        // on B200 (PCG 8.69 vs 8.05 ms at C4), so it is opt-in: OSFM_BA_PCG_B128=1
        static const bool allow_b128 = []() { const char* e = getenv("OSFM_BA_PCG_B128"); return e && e[0] == '1'; }();
        pcg_pipe.b128 = (allow_b128 && pcg_grid <= PCG_B128_GROUP && PCG_THREADS >= 3 * PCG_B128_GROUP) ? 1 : 0;
        OSFM_CUDA(cudaFuncSetAttribute(pcg_pipelined, cudaFuncAttributeMaxDynamicSharedMemorySize, pcg_pipe_smem));
      }
    }
    OSFM_CUDA(cudaStreamSynchronize(stream));  // rowbase / cbase host vectors go out of scope
  }
  if (nblk == 0) d_Sbuf.reserve(nc_pad + 16);
  double* const d_rhs_p = d_Sbuf.p;          // [nc] right-hand side
  double* const d_S_p = d_Sbuf.p + nc_pad;   // block values: upper blocks first
  PcgLayout lay{};
  lay.row_of = d_row_of.p; lay.row_M = d_row_M.p; lay.rowbase = d_rowbase.p; lay.cbase = d_cbase.p;
  lay.colidx = d_colidx.p; lay.ngroups = ngroups; lay.grp_b1 = d_grp_b1.p; lay.grp_b2 = d_grp_b2.p;

  trace("structure");
  // ---- Levenberg-Marquardt (Ceres trust_region_minimizer / levenberg_marquardt_strategy) ----
  if (world > 1) {  // all ranks enter the timed region together (their set-up times differ)
    OSFM_CUDA(cudaMemsetAsync(d_sc.p, 0, sizeof(Scalars), stream));
    allreduce_dev(&d_sc.p->cost, 1);
    OSFM_CUDA(cudaStreamSynchronize(stream));
  }
  OSFM_CUDA(cudaEventRecord(ev0, stream));
  double radius = 1e4;
  const double max_radius = 1e16, min_radius = 1e-32, min_rel_decrease = 1e-3;
  const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int n_invalid = 0, it = 0, n_success = 0, n_solves = 0, pcg_total = 0;
  pcg_fallbacks = 0;
  int termination = 1;
  std::string message = "Maximum number of iterations reached.";

  double grad_max = 0.0;
  have_seg_tab = false;   // the tables / chunk list of a previous run() do not describe this problem
  sp_nchunks = 0;
  double cost = linearize(cur, &grad_max);
  const double initial_cost = cost;
  if (n > 0) {
    ba_make_scale<<<grid_for(n, 256), 256, 0, stream>>>(d_colnorm2.p, d_scale.p, n);
    OSFM_LAUNCH_CHECK();
  }
  // deflation vectors of the reduced solve: the similarity gauge at the initial poses, in the scaled variables
  static const bool deflate_on = []() { const char* e = getenv("OSFM_BA_PCG_DEFLATE"); return !(e && e[0] == '0'); }();
  pcg_pipe.Wdef = nullptr;
  if (deflate_on && pcg_pipe_ok && !pcg_pipe.b128 && NI > 0 && nc > 0) {
    d_Wdef.reserve((size_t)PCG_ND * nc);
    OSFM_CUDA(cudaMemsetAsync(d_Wdef.p, 0, sizeof(double) * PCG_ND * (size_t)nc, stream));
    pcg_gauge_vectors<<<grid_for(NI, 128), 128, 0, stream>>>(NI, d_inst_poff.p, params_of(cur).inst, d_scale.p, nc, d_Wdef.p);
    OSFM_LAUNCH_CHECK();
    pcg_pipe.Wdef = d_Wdef.p;
  }
  // per-segment tables of the tensor-core Schur kernel (columns, block offsets, Jacobi scales): constant from here on
  static const bool mma_on = []() { const char* e = getenv("OSFM_BA_SCHUR_MMA"); return !(e && e[0] == '0'); }();
  // The tensor-core kernels add the same-shot blocks J^T J only in the tiles (t, t) and (t, t + 1): a shot's wc
  // columns must not span three 8-wide tiles, i.e. wc <= 9.  Wider camera sides (Brown: 9 + 6, rig cameras: + 6)
  // use the SIMT segment kernels.
  const bool use_mma = mma_on && wc <= 9;
  bool use_pipe = false;
  sp_nchunks = 0;
  have_seg_tab = false;
  if (nseg > 0 && use_mma && nblk > 0) {
    d_tab_off.reserve((size_t)nseg + 1); d_tab_sizes.reserve((size_t)nseg + 1);
    ba_seg_table_sizes<<<grid_for(nseg + 1, 256), 256, 0, stream>>>(v, d_seg_start.p, nseg, d_tab_sizes.p);
    OSFM_LAUNCH_CHECK();
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, d_tab_sizes.p, d_tab_off.p, nseg + 1, stream);
    d_cub.reserve(tb + 256);
    tb = d_cub.cap;
    OSFM_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tb, d_tab_sizes.p, d_tab_off.p, nseg + 1, stream));
    long long total_ints = 0;
    OSFM_CUDA(cudaMemcpyAsync(&total_ints, d_tab_off.p + nseg, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    OSFM_CUDA(cudaStreamSynchronize(stream));
    d_tab.reserve((size_t)total_ints + 2);
    ba_seg_tables<<<nseg, 128, 0, stream>>>(v, bm, bsr, d_seg_start.p, d_scale.p, d_tab_off.p, d_tab.p);
    OSFM_LAUNCH_CHECK();
    have_seg_tab = true;
    // chunk list of the persistent Schur kernel (ba_schur_pipe.cuh)
    static const bool pipe_on = []() { const char* e = getenv("OSFM_BA_SCHUR_PIPE"); return !(e && e[0] == '0'); }();
    // (the flush table holds offset << 2: the reduced system must stay below 2^29 doubles; 20 KB of table per segment)
    use_pipe = pipe_on && v.nres * (wc + 4) <= SP_ROWS && s_upper_total + (long long)nc_pad < (1LL << 29) &&
               (long long)nseg * SP_FT_SEG * (long long)sizeof(int) <= (8LL << 30);
    if (use_pipe) {
      d_sp_nch.reserve((size_t)nseg + 1); d_sp_chunk0.reserve((size_t)nseg + 1);
      sp_chunk_counts<<<grid_for(nseg + 1, 256), 256, 0, stream>>>(v, d_seg_start.p, nseg, d_sp_nch.p);
      OSFM_LAUNCH_CHECK();
      size_t tb2 = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_sp_nch.p, d_sp_chunk0.p, nseg + 1, stream);
      d_cub.reserve(tb2 + 256);
      tb2 = d_cub.cap;
      OSFM_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tb2, d_sp_nch.p, d_sp_chunk0.p, nseg + 1, stream));
      OSFM_CUDA(cudaMemcpyAsync(&sp_nchunks, d_sp_chunk0.p + nseg, sizeof(int), cudaMemcpyDeviceToHost, stream));
      OSFM_CUDA(cudaStreamSynchronize(stream));
      d_sp_chunks.reserve((size_t)sp_nchunks + 1);
      sp_fill_chunks<<<grid_for((long long)nseg * 32, 256), 256, 0, stream>>>(v, d_seg_start.p, nseg, d_sp_chunk0.p, d_tab_off.p,
                                                                            d_sp_chunks.p);
      OSFM_LAUNCH_CHECK();
      d_sp_ftab.reserve((size_t)nseg * SP_FT_SEG);
      sp_flush_tables<<<nseg, SP_CONS_THREADS, 0, stream>>>(v, d_seg_start.p, d_tab_off.p, d_tab.p, d_sp_ftab.p);
      OSFM_LAUNCH_CHECK();
    }
  }
  // |x| of the free parameters
  auto x_norm_of = [&](int b) -> double {
    OSFM_CUDA(cudaMemsetAsync(&d_sc.p->x_norm2, 0, sizeof(double), stream));
    OSFM_CUDA(cudaMemsetAsync(d_y.p, 0, sizeof(double) * nz, stream));
    Params pp = params_of(b);
    if (K) { ba_update<<<grid_for(K, 128), 128, 0, stream>>>(0, K, d_cam_poff.p, d_cam_off.p, d_cam_np.p, 0, 0, pp.cam, pp.cam, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
    if (NI) { ba_update<<<grid_for(NI, 128), 128, 0, stream>>>(1, NI, d_inst_poff.p, nullptr, nullptr, 6, 0, pp.inst, pp.inst, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
    if (NR) { ba_update<<<grid_for(NR, 128), 128, 0, stream>>>(2, NR, d_rc_poff.p, nullptr, nullptr, 6, 0, pp.rc, pp.rc, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
    if (P) { ba_update<<<grid_for(P, 128), 128, 0, stream>>>(3, P, d_pt_poff.p, nullptr, nullptr, 3, nc, pp.pts, pp.pts, d_scale.p, d_y.p, d_sc.p, 1); OSFM_LAUNCH_CHECK(); }
    if (NE) { ba_update<<<grid_for(NE, 128), 128, 0, stream>>>(0, NE, d_ext_poff.p, d_ext_off.p, d_ext_np.p, 0, 0, pp.ext, pp.ext, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
    allreduce_dev(&d_sc.p->x_norm2, 1);
    return std::sqrt(read_scalars().x_norm2);
  };
  double x_norm = n > 0 ? x_norm_of(cur) : 0.0;

  if (grad_max <= gtol || n == 0) {
    termination = 0;
    message = n == 0 ? "No free parameters." : "Gradient tolerance reached.";
  }
  while (termination == 1) {
    if (it >= max_iterations) break;
    if (radius < min_radius) { termination = 0; message = "Minimum trust region radius reached."; break; }
    ++it;
    if (!reuse_diagonal) {
      ba_make_diag<<<grid_for(n, 256), 256, 0, stream>>>(d_colnorm2.p, d_scale.p, d_diag.p, n);
      OSFM_LAUNCH_CHECK();
    }
    const double inv_radius = 1.0 / radius;
    // --- reduced camera system (upper blocks accumulated with L2 atomics) ---
    if (nc > 0) {
      OSFM_CUDA(cudaMemsetAsync(d_Sbuf.p, 0, sizeof(double) * ((size_t)nc_pad + (size_t)s_upper_total), stream));
    }
    if (P > 0) {
      const size_t smem = (size_t)SCHUR_KC * wc * (2 * 3 * sizeof(double) + 2 * sizeof(int)) +
                          (size_t)SCHUR_KC * 8 * sizeof(int) + (size_t)SCHUR_KC * SCHUR_KC * 9 * sizeof(int);
      tm_schur.start(stream);
      if (nseg > 0) {
        if (!seg_attr) {
          OSFM_CUDA(cudaFuncSetAttribute(ba_schur_seg<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SegSmem)));
          OSFM_CUDA(cudaFuncSetAttribute(ba_schur_seg<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SegSmem)));
          seg_attr = true;
        }
        // default: fused fp64 tensor-core kernel; OSFM_BA_SCHUR_MMA=0 -> the older ba_obs_rows + ba_schur_seg pair
        d_Vig.reserve(3 * (size_t)std::max(npf, 1));
        ba_point_blocks<<<grid_for(P_fast, 128), 128, 0, stream>>>(v, P_fast, d_scale.p, d_diag.p, inv_radius, d_Vinv.p,
                                                                 d_gp.p, d_Vig.p);
        OSFM_LAUNCH_CHECK();
        if (use_mma) {
          if (!mma_attr) {
            OSFM_CUDA(cudaFuncSetAttribute(ba_schur_mma<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SegMmaSmem)));
            OSFM_CUDA(cudaFuncSetAttribute(ba_schur_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SegMmaSmem)));
            mma_attr = true;
          }
          unsigned long long* prof = nullptr;
          if (trace_on) {
            d_prof.reserve(16);
            OSFM_CUDA(cudaMemsetAsync(d_prof.p, 0, 16 * sizeof(unsigned long long), stream));
            prof = d_prof.p;
          }
          if (use_pipe && sp_nchunks > 0) {
            if (!sp_attr) {
              OSFM_CUDA(cudaFuncSetAttribute(ba_schur_pipe<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem)));
              OSFM_CUDA(cudaFuncSetAttribute(ba_schur_pipe<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem)));
              OSFM_CUDA(cudaFuncSetAttribute(ba_schur_pipe<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem)));
              OSFM_CUDA(cudaFuncSetAttribute(ba_schur_pipe<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpSmem)));
              sp_attr = true;
            }
            const int grid = std::max(1, std::min(num_sms, sp_nchunks));
            auto kern = wc == 9 ? (prof ? ba_schur_pipe<9, true> : ba_schur_pipe<9, false>)
                                : (prof ? ba_schur_pipe<0, true> : ba_schur_pipe<0, false>);
            kern<<<grid, SP_THREADS, sizeof(SpSmem), stream>>>(v, d_sp_chunks.p, sp_nchunks, d_tab.p, d_scale.p, d_Vinv.p, d_Vig.p,
                                                              d_sp_ftab.p, d_S_p, d_rhs_p, prof);
            if (trace_on) {
              OSFM_LAUNCH_CHECK();
              unsigned long long hp[16];
              OSFM_CUDA(cudaMemcpyAsync(hp, d_prof.p, sizeof(hp), cudaMemcpyDeviceToHost, stream));
              OSFM_CUDA(cudaStreamSynchronize(stream));
              const unsigned long long nch = std::max<unsigned long long>(hp[3], 1);
              fprintf(stderr, "[osfm_ba] ba_schur_pipe (%d CTAs, %llu chunks) producer clocks / chunk: copy wait %llu buffer wait %llu build %llu\n",
                      grid, hp[3], hp[0] / nch, hp[1] / nch, hp[2] / nch);
              for (int gI = 0; gI < 2; ++gI) {
                const unsigned long long ns = std::max<unsigned long long>(hp[8 + 5 * gI], 1);
                fprintf(stderr, "[osfm_ba]   consumer group %d clocks / segment (%llu segments): tables %llu operand wait %llu mma %llu flush %llu\n",
                        gI, hp[8 + 5 * gI], hp[4 + 5 * gI] / ns, hp[5 + 5 * gI] / ns, hp[6 + 5 * gI] / ns, hp[7 + 5 * gI] / ns);
              }
              prof = nullptr;
            }
          } else if (wc == 9)
            ba_schur_mma<9><<<nseg, SM_THREADS, sizeof(SegMmaSmem), stream>>>(v, d_seg_start.p, d_tab_off.p, d_tab.p, d_scale.p,
                                                                             d_Vinv.p, d_Vig.p, d_S_p, d_rhs_p, prof);
          else
            ba_schur_mma<0><<<nseg, SM_THREADS, sizeof(SegMmaSmem), stream>>>(v, d_seg_start.p, d_tab_off.p, d_tab.p, d_scale.p,
                                                                             d_Vinv.p, d_Vig.p, d_S_p, d_rhs_p, prof);
          OSFM_LAUNCH_CHECK();
          if (trace_on && prof) {
            unsigned long long hp[8];
            OSFM_CUDA(cudaMemcpyAsync(hp, d_prof.p, sizeof(hp), cudaMemcpyDeviceToHost, stream));
            OSFM_CUDA(cudaStreamSynchronize(stream));
            fprintf(stderr, "[osfm_ba] ba_schur_mma clocks / segment (thread 0): structure %llu offsets %llu loads %llu rows %llu mma %llu flush %llu\n",
                    hp[0] / nseg, hp[1] / nseg, hp[2] / nseg, hp[3] / nseg, hp[4] / nseg, hp[5] / nseg);
          }
        } else {
          const long long n_fast = n_fast_obs;
          d_rowsJ.reserve((size_t)n_fast * wc * 3 + 8); d_rowsW.reserve((size_t)n_fast * wc * 3 + 8);
          d_rowsY.reserve((size_t)n_fast * wc * 3 + 8);
          ba_obs_rows<<<grid_for(n_fast * wc, 256), 256, 0, stream>>>(v, bm, bsr, n_fast, d_scale.p, d_Vinv.p, d_Vig.p,
                                                                    d_rowsJ.p, d_rowsW.p, d_rowsY.p, d_rhs_p);
          OSFM_LAUNCH_CHECK();
          if (wc == 9)
            ba_schur_seg<9><<<nseg, SEG_THREADS, sizeof(SegSmem), stream>>>(v, bm, bsr, d_seg_start.p, n_fast, d_rowsJ.p,
                                                                           d_rowsW.p, d_rowsY.p, d_S_p);
          else
            ba_schur_seg<0><<<nseg, SEG_THREADS, sizeof(SegSmem), stream>>>(v, bm, bsr, d_seg_start.p, n_fast, d_rowsJ.p,
                                                                           d_rowsW.p, d_rowsY.p, d_S_p);
          OSFM_LAUNCH_CHECK();
        }
      }
      if (P > P_fast) {
        ba_schur<<<P - P_fast, SCHUR_THREADS, smem, stream>>>(v, bm, bsr, d_scale.p, d_diag.p, inv_radius, d_S_p,
                                                             d_rhs_p, d_Vinv.p, d_gp.p, P_fast, ppv, d_pts[cur].p);
        OSFM_LAUNCH_CHECK();
      }
      tm_schur.stop(stream);
    }
    bool ok = true;
    int pcg_it = 0;
    OSFM_CUDA(cudaMemsetAsync(d_y.p, 0, sizeof(double) * nz, stream));
    if (nc > 0) {
      // the one exchange step of the LM iteration: sum of the partial reduced systems over ranks
      if (world > 1) allreduce_dev(d_Sbuf.p, (long long)nc_pad + s_upper_total);  // rhs + upper blocks, one call
      {
        // S_g stays a pure partial sum through the all-reduce; every rank then adds the (replicated)
        // prior rows and the damping to its copy of the reduced system.
        PriorView pall = pv;
        pall.n_cam_rows = (int)pr_cam_param.size();
        pall.n_pos_rows = (int)pr_pos_inst.size();
        const int nall = pall.n_cam_rows + pall.n_pos_rows;
        if (nall > 0) {
          ba_prior_system<<<grid_for(nall, 128), 128, 0, stream>>>(pall, params_of(cur), d_scale.p, d_prior_diag_off.p,
                                                                  d_S_p, d_rhs_p);
          OSFM_LAUNCH_CHECK();
        }
        if (NT > 0) {
          side_system<<<NT, SIDE_THREADS, 0, stream>>>(sv, v, bm, params_of(cur), bsr, d_scale.p, d_S_p, d_rhs_p);
          OSFM_LAUNCH_CHECK();
        }
      }
      ba_finish_system<<<grid_for((long long)n_upper * 32, 256), 256, 0, stream>>>(d_upper.p, n_upper, bsr, d_S_p,
                                                                                 d_diag.p, inv_radius);
      OSFM_LAUNCH_CHECK();
      // --- PCG: one persistent kernel, |r| <= 1e-8 |b| ---
      tm_pcg.start(stream);
      pcg_convert<<<grid_for((long long)n_blocks_all * 32, 256), 256, 0, stream>>>(
          d_S_p, d_row_col.p, d_row_off.p, d_qoff.p, d_blk_row.p, n_blocks_all, bsr, d_row_M.p, d_rowbase.p, d_Spcg.p);
      OSFM_LAUNCH_CHECK();
      pcg_factor_groups<<<grid_for(ngroups, 64), 64, 0, stream>>>(d_S_p, bsr, d_diag_off.p, d_grp_b1.p, d_grp_b2.p,
                                                                 ngroups, d_Minv.p);
      OSFM_LAUNCH_CHECK();
      OSFM_CUDA(cudaMemsetAsync(d_pcg.p, 0, sizeof(PcgState), stream));
      const int max_pcg = std::min(2 * nc + 100, 5000);
      bool solved = false;
      if (pcg_pipe_ok) {
        launch_cooperative(pcg_pipelined, pcg_grid, PCG_THREADS, pcg_pipe_smem, stream, d_Spcg.p, lay, bsr, d_Minv.p, d_rhs_p,
                           d_px.p, d_pz.p, d_pp.p, d_pcg.p, nc, max_pcg, 1e-16, pcg_pipe);
        OSFM_LAUNCH_CHECK();
        OSFM_CUDA(cudaMemcpyAsync(h_pcg.p, d_pcg.p, PCG_STATE_HEADER, cudaMemcpyDeviceToHost, stream));
        OSFM_CUDA(cudaStreamSynchronize(stream));
        solved = h_pcg.p->converged != 0;
        pcg_pipe_its = h_pcg.p->iterations;
        if (trace_on)
          fprintf(stderr, "[osfm_ba] pipelined pcg %d its converged %d, CTA0 clocks/it: stage %lld matvec %lld reduce %lld update %lld"
                  " | wide barrier (all calls / its): own sums %lld release %lld collect %lld column sums %lld\n",
                  h_pcg.p->iterations, h_pcg.p->converged, h_pcg.p->prof[0] / std::max(pcg_pipe_its, 1),
                  h_pcg.p->prof[1] / std::max(pcg_pipe_its, 1), h_pcg.p->prof[2] / std::max(pcg_pipe_its, 1),
                  h_pcg.p->prof[3] / std::max(pcg_pipe_its, 1), h_pcg.p->prof[4] / std::max(pcg_pipe_its, 1),
                  h_pcg.p->prof[5] / std::max(pcg_pipe_its, 1), h_pcg.p->prof[6] / std::max(pcg_pipe_its, 1),
                  h_pcg.p->prof[7] / std::max(pcg_pipe_its, 1));
        if (!solved) {  // stagnation / breakdown of the pipelined recurrences: classic PCG from scratch
          pcg_total += pcg_pipe_its;
          ++pcg_fallbacks;
          OSFM_CUDA(cudaMemsetAsync(d_pcg.p, 0, sizeof(PcgState), stream));
        }
      }
      if (!solved) {
        if (pcg_resident)
          launch_cooperative(pcg_persistent<true>, pcg_grid, PCG_THREADS, pcg_smem, stream, d_Spcg.p, lay, bsr, d_Minv.p,
                             d_rhs_p, d_px.p, d_pr.p, d_pz.p, d_pp.p, d_pAp.p, d_Ap.p, d_pcg.p, nc, max_pcg, 1e-16, pcg_res);
        else
          launch_cooperative(pcg_persistent<false>, pcg_grid, PCG_THREADS, 0, stream, d_Spcg.p, lay, bsr, d_Minv.p, d_rhs_p,
                             d_px.p, d_pr.p, d_pz.p, d_pp.p, d_pAp.p, d_Ap.p, d_pcg.p, nc, max_pcg, 1e-16, pcg_res);
        OSFM_LAUNCH_CHECK();
      }
      OSFM_CUDA(cudaMemcpyAsync(d_y.p, d_px.p, sizeof(double) * nc, cudaMemcpyDeviceToDevice, stream));
      OSFM_CUDA(cudaMemcpyAsync(h_pcg.p, d_pcg.p, PCG_STATE_HEADER, cudaMemcpyDeviceToHost, stream));
      tm_pcg.stop(stream);
    }
    ++n_solves;
    // --- back-substitution, model cost change ---
    if (P > 0 && npf > 0) {
      tm_back.start(stream);
      d_bs_t.reserve(3 * (size_t)npf);
      OSFM_CUDA(cudaMemsetAsync(d_bs_t.p, 0, sizeof(double) * 3 * (size_t)npf, stream));
      if (N > 0) {
        ba_backsub_rows<<<grid_for(N, 256), 256, 0, stream>>>(v, d_scale.p, d_y.p, d_bs_t.p);
        OSFM_LAUNCH_CHECK();
      }
      if (have_pp) {
        ba_point_prior<3><<<grid_for(P, 128), 128, 0, stream>>>(ppv, v, params_of(cur), nullptr, nullptr, nullptr, nullptr, d_bs_t.p, d_sc.p);
        OSFM_LAUNCH_CHECK();
      }
      ba_backsub_points<<<grid_for(npf, 256), 256, 0, stream>>>(v, d_scale.p, d_Vinv.p, d_bs_t.p, d_y.p);
      OSFM_LAUNCH_CHECK();
      tm_back.stop(stream);
    }
    OSFM_CUDA(cudaMemsetAsync(&d_sc.p->model_change, 0, sizeof(double) * 3, stream));  // model_change, step_norm2, x_norm2
    static const bool mc_explicit = []() { const char* e = getenv("OSFM_BA_MODEL_CHANGE_EXPLICIT"); return e && e[0] == '1'; }();
    if (!mc_explicit) {
      if (n > 0) {
        ba_model_change_alg<<<grid_for(n, 256), 256, 0, stream>>>(n, nc, rank == 0, d_grad.p, d_scale.p, d_diag.p, inv_radius, d_y.p, d_sc.p);
        OSFM_LAUNCH_CHECK();
      }
    } else {
    if (N > 0) {
      ba_model_change<<<grid_for(N, 256), 256, 0, stream>>>(v, d_scale.p, d_y.p, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (npr_local > 0) {
      ba_prior_model_change<<<grid_for(npr_local, 128), 128, 0, stream>>>(pv, params_of(cur), d_scale.p, d_y.p, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (NT > 0 && add_priors) {
      side_model_change<<<grid_for(NT, 128), 128, 0, stream>>>(sv, v, bm, params_of(cur), d_scale.p, d_y.p, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    if (have_pp && P > 0) {
      ba_point_prior<2><<<grid_for(P, 128), 128, 0, stream>>>(ppv, v, params_of(cur), d_scale.p, d_y.p, nullptr, nullptr, nullptr, d_sc.p);
      OSFM_LAUNCH_CHECK();
    }
    }
    // --- candidate point ---
    const int cand = cur ^ 1;
    auto update_candidate = [&](double alpha) {   // candidate = Project(x - alpha * scale * y); |delta|^2, |x|^2
      OSFM_CUDA(cudaMemsetAsync(&d_sc.p->step_norm2, 0, sizeof(double) * 2, stream));
      Params a = params_of(cur), b = params_of(cand);
      if (K) { ba_update<<<grid_for(K, 128), 128, 0, stream>>>(0, K, d_cam_poff.p, d_cam_off.p, d_cam_np.p, 0, 0, a.cam, b.cam, d_scale.p, d_y.p, d_sc.p, rank == 0, nullptr, alpha); OSFM_LAUNCH_CHECK(); }
      if (NI) { ba_update<<<grid_for(NI, 128), 128, 0, stream>>>(1, NI, d_inst_poff.p, nullptr, nullptr, 6, 0, a.inst, b.inst, d_scale.p, d_y.p, d_sc.p, rank == 0, nullptr, alpha); OSFM_LAUNCH_CHECK(); }
      if (NR) { ba_update<<<grid_for(NR, 128), 128, 0, stream>>>(2, NR, d_rc_poff.p, nullptr, nullptr, 6, 0, a.rc, b.rc, d_scale.p, d_y.p, d_sc.p, rank == 0, nullptr, alpha); OSFM_LAUNCH_CHECK(); }
      if (P) { ba_update<<<grid_for(P, 128), 128, 0, stream>>>(3, P, d_pt_poff.p, nullptr, nullptr, 3, nc, a.pts, b.pts, d_scale.p, d_y.p, d_sc.p, 1, nullptr, alpha); OSFM_LAUNCH_CHECK(); }
      if (NE) { ba_update<<<grid_for(NE, 128), 128, 0, stream>>>(0, NE, d_ext_poff.p, d_ext_off.p, d_ext_np.p, 0, 0, a.ext, b.ext, d_scale.p, d_y.p, d_sc.p, rank == 0, d_ext_lower.p, alpha); OSFM_LAUNCH_CHECK(); }
      if (world > 1) allreduce_dev(&d_sc.p->step_norm2, 2);
      return std::sqrt(read_scalars().step_norm2);
    };
    {
      Params a = params_of(cur), b = params_of(cand);
      if (K) { ba_update<<<grid_for(K, 128), 128, 0, stream>>>(0, K, d_cam_poff.p, d_cam_off.p, d_cam_np.p, 0, 0, a.cam, b.cam, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
      if (NI) { ba_update<<<grid_for(NI, 128), 128, 0, stream>>>(1, NI, d_inst_poff.p, nullptr, nullptr, 6, 0, a.inst, b.inst, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
      if (NR) { ba_update<<<grid_for(NR, 128), 128, 0, stream>>>(2, NR, d_rc_poff.p, nullptr, nullptr, 6, 0, a.rc, b.rc, d_scale.p, d_y.p, d_sc.p, rank == 0); OSFM_LAUNCH_CHECK(); }
      if (P) { ba_update<<<grid_for(P, 128), 128, 0, stream>>>(3, P, d_pt_poff.p, nullptr, nullptr, 3, nc, a.pts, b.pts, d_scale.p, d_y.p, d_sc.p, 1); OSFM_LAUNCH_CHECK(); }
      if (NE) { ba_update<<<grid_for(NE, 128), 128, 0, stream>>>(0, NE, d_ext_poff.p, d_ext_off.p, d_ext_np.p, 0, 0, a.ext, b.ext, d_scale.p, d_y.p, d_sc.p, rank == 0, d_ext_lower.p); OSFM_LAUNCH_CHECK(); }
    }
    if (world > 1) allreduce_dev(&d_sc.p->model_change, 3);
    const Scalars sm = read_scalars();
    if (nc > 0) {
      pcg_it = h_pcg.p->iterations;
      pcg_total += pcg_it;
      const double rr = h_pcg.p->rr_final;
      if (trace_on)
        fprintf(stderr, "[osfm_ba] pcg %d its, CTA0 clocks/it: stage %lld matvec %lld reduce1 %lld phaseB %lld reduce2 %lld (resident %d)\n",
                pcg_it, h_pcg.p->prof[0] / std::max(pcg_it, 1), h_pcg.p->prof[1] / std::max(pcg_it, 1),
                h_pcg.p->prof[2] / std::max(pcg_it, 1), h_pcg.p->prof[3] / std::max(pcg_it, 1),
                h_pcg.p->prof[4] / std::max(pcg_it, 1), (int)pcg_resident);
      if (!(rr == rr)) ok = false;
    }
    const double model_change = sm.model_change;
    double step_norm = std::sqrt(sm.step_norm2);
    if (!ok || !(model_change > 0.0) || !std::isfinite(step_norm)) {
      if (++n_invalid >= 5) { termination = 2; message = "Too many consecutive invalid steps."; break; }
      radius *= 0.5;
      reuse_diagonal = true;
      continue;
    }
    n_invalid = 0;
    double cand_cost = eval_cost(cand);
    if (constrained) {
      // Ceres: a problem with parameter bounds is "constrained": TrustRegionMinimizer::DoLineSearch runs a projected
      // Armijo search along the step (sufficient decrease 1e-4, at most 20 contractions; bisection here, Ceres'
      // default interpolates a cubic) and the candidate is the point it accepts.  The model cost change stays
      // that of the full step, as in Ceres.
      OSFM_CUDA(cudaMemsetAsync(&d_sc.p->gdot, 0, sizeof(double), stream));
      ba_grad_dot<<<grid_for(n, 256), 256, 0, stream>>>(n, nc, rank == 0, d_grad.p, d_scale.p, d_y.p, d_sc.p);
      OSFM_LAUNCH_CHECK();
      if (world > 1) allreduce_dev(&d_sc.p->gdot, 1);
      const double g0 = read_scalars().gdot;
      double alpha = 1.0;
      bool ok_ls = false;
      for (int ls = 0; ls < 20; ++ls) {
        if (ls > 0) { step_norm = update_candidate(alpha); cand_cost = eval_cost(cand); }
        if (std::isfinite(cand_cost) && cand_cost <= cost + 1e-4 * g0 * alpha) { ok_ls = true; break; }
        alpha *= 0.5;
      }
      if (!ok_ls) { step_norm = update_candidate(1.0); cand_cost = eval_cost(cand); }
    }
    if (step_norm <= ptol * (x_norm + ptol)) { termination = 0; message = "Parameter tolerance reached."; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= ftol * cost) { termination = 0; message = "Function tolerance reached."; break; }
    const double rel = cost_change / model_change;
    if (rel > min_rel_decrease) {
      cur = cand;
      cost = linearize(cur, &grad_max);
      x_norm = x_norm_of(cur);
      radius = std::min(max_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3)));
      decrease_factor = 2.0;
      reuse_diagonal = false;
      ++n_success;
      if (grad_max <= gtol) { termination = 0; message = "Gradient tolerance reached."; break; }
    } else {
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  OSFM_CUDA(cudaEventRecord(ev1, stream));
  const double final_cost = eval_cost(cur);

  trace("lm");
  // ---- results back to the host ----
  OSFM_CUDA(cudaMemcpyAsync(cam_params.data(), d_cam[cur].p, sizeof(double) * cam_params.size(), cudaMemcpyDeviceToHost, stream));
  OSFM_CUDA(cudaMemcpyAsync(inst.data(), d_inst[cur].p, sizeof(double) * inst.size(), cudaMemcpyDeviceToHost, stream));
  if (!rc.empty())
    OSFM_CUDA(cudaMemcpyAsync(rc.data(), d_rc[cur].p, sizeof(double) * rc.size(), cudaMemcpyDeviceToHost, stream));
  if (!ext_values.empty())
    OSFM_CUDA(cudaMemcpyAsync(ext_values.data(), d_ext[cur].p, sizeof(double) * ext_values.size(), cudaMemcpyDeviceToHost, stream));
  d_full_pts.reserve(3 * (size_t)std::max(Pfull, 1));
  if (world > 1) OSFM_CUDA(cudaMemsetAsync(d_full_pts.p, 0, sizeof(double) * 3 * (size_t)Pfull, stream));
  if (P > 0) {
    ord_scatter_points<<<grid_for(P, 256), 256, 0, stream>>>(P, d_pts[cur].p, d_global_of.p, d_full_pts.p);
    OSFM_LAUNCH_CHECK();
  }
  allreduce_dev(d_full_pts.p, 3 * (long long)Pfull);
  if (Pfull > 0)
    OSFM_CUDA(cudaMemcpyAsync(pts.data(), d_full_pts.p, sizeof(double) * 3 * (size_t)Pfull, cudaMemcpyDeviceToHost, stream));
  // the reprojection errors stay on the device until osfm_ba_get_reprojection_errors fetches them
  reproj_valid = false;
  if (compute_reproj && Nfull > 0) {
    d_reproj.reserve(3 * (size_t)Nfull);
    OSFM_CUDA(cudaMemsetAsync(d_reproj.p, 0, sizeof(double) * 3 * (size_t)Nfull, stream));
    if (N > 0) {
      ba_linearize<2><<<grid_for(N, 128), 128, 0, stream>>>(v, params_of(cur), d_sc.p, d_reproj.p);
      OSFM_LAUNCH_CHECK();
    }
    allreduce_dev(d_reproj.p, 3 * Nfull);
    reproj_valid = true;
  }
  OSFM_CUDA(cudaStreamSynchronize(stream));
  trace("results");
  if (trace_on && world > 1)
    fprintf(stderr, "[osfm_ba] rank %d: %d all-reduces, host %.3f ms, device (traced, serialised) %.3f ms\n", rank, ar_calls,
            ar_host_ms, ar_dev_ms);
  float dev_ms = 0.f;
  OSFM_CUDA(cudaEventElapsedTime(&dev_ms, ev0, ev1));
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  tm_lin.collect(); tm_schur.collect(); tm_pcg.collect(); tm_back.collect();

  summary = osfm_ba_summary{};
  summary.iterations = it;
  summary.successful_steps = n_success;
  summary.linear_solves = n_solves;
  summary.pcg_iterations = pcg_total;
  summary.termination = termination;
  summary.initial_cost = initial_cost;
  summary.final_cost = final_cost;
  summary.time_device_ms = dev_ms;
  summary.time_linearize_ms = tm_lin.total_ms;
  summary.linearize_launches = tm_lin.count;
  summary.time_schur_ms = tm_schur.total_ms;
  summary.schur_launches = tm_schur.count;
  summary.time_pcg_ms = tm_pcg.total_ms;
  summary.time_backsub_ms = tm_back.total_ms;
  summary.num_observations_local = N;
  summary.reduced_dim = nc;
  summary.reduced_blocks = n_blocks_all;
  summary.reduced_nnz = s_total;
  summary.jac_planes = nres * (wc + 3 + 1);
  tm_lin.destroy(); tm_schur.destroy(); tm_pcg.destroy(); tm_back.destroy();
  summary.kernel_launches = g_kernel_launches.load() - launches0;
  snprintf(summary.message, sizeof(summary.message), "%s", message.c_str());
  summary.time_run_s =
      std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t_start).count();
  has_run = true;
}

}  // namespace osfm

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct osfm_ba {
  osfm::BA impl;
  explicit osfm_ba(int dev) : impl(dev) {}
};
using osfm::ArgError;

extern "C" {

int osfm_camera_num_params(int projection_type) {
  if (projection_type < 0 || projection_type > 9) return -1;
  return osfm::model_num_params(projection_type);
}

int osfm_ba_create(int device, osfm_ba** out) {
  OSFM_API_BEGIN
  if (!out) throw ArgError("null out");
  int count = 0;
  OSFM_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) throw ArgError("no such CUDA device");
  *out = new osfm_ba(device);
  OSFM_API_END
}
int osfm_ba_destroy(osfm_ba* ba) {
  OSFM_API_BEGIN
  delete ba;
  OSFM_API_END
}
#define OSFM_BA_CHECK if (!ba) throw ArgError("null ba handle");

int osfm_ba_set_cameras(osfm_ba* ba, int n, const int32_t* type, const double* params, const int32_t* constant,
                        const double* prior, const double* prior_sigma, const int32_t* prior_log) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  auto& b = ba->impl;
  if (n < 0 || (n > 0 && (!type || !params || !constant || !prior || !prior_sigma || !prior_log)))
    throw ArgError("bad camera arrays");
  int total = 0;
  for (int k = 0; k < n; ++k) {
    if (type[k] < 0 || type[k] > 9) throw ArgError("Invalid ProjectionType");  // camera_instances.h:232
    total += osfm::model_num_params(type[k]);
  }
  b.cam_type.assign(type, type + n);
  b.cam_const.assign(constant, constant + n);
  b.cam_params.assign(params, params + total);
  b.cam_prior.assign(prior, prior + total);
  b.cam_prior_sigma.assign(prior_sigma, prior_sigma + total);
  b.cam_prior_log.assign(prior_log, prior_log + total);
  OSFM_API_END
}
int osfm_ba_set_rig_instances(osfm_ba* ba, int n, const double* pose6, const int32_t* constant,
                              const int32_t* has_position_prior, const double* prior_position3,
                              const double* prior_std3) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  auto& b = ba->impl;
  if (n < 0 || (n > 0 && (!pose6 || !constant))) throw ArgError("bad rig instance arrays");
  b.inst.assign(pose6, pose6 + 6 * (size_t)n);
  b.inst_const.assign(constant, constant + n);
  b.inst_has_prior.assign(n, 0);
  b.inst_prior_pos.assign(3 * (size_t)n, 0.0);
  b.inst_prior_std.assign(3 * (size_t)n, 1.0);
  if (has_position_prior) {
    if (!prior_position3 || !prior_std3) throw ArgError("position prior arrays missing");
    b.inst_has_prior.assign(has_position_prior, has_position_prior + n);
    b.inst_prior_pos.assign(prior_position3, prior_position3 + 3 * (size_t)n);
    b.inst_prior_std.assign(prior_std3, prior_std3 + 3 * (size_t)n);
  }
  OSFM_API_END
}
int osfm_ba_set_rig_cameras(osfm_ba* ba, int n, const double* pose6, const int32_t* constant) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!pose6 || !constant))) throw ArgError("bad rig camera arrays");
  ba->impl.rc.assign(pose6, pose6 + 6 * (size_t)n);
  ba->impl.rc_const.assign(constant, constant + n);
  ba->impl.rc_prior.clear();
  ba->impl.rc_prior_sigma.clear();
  OSFM_API_END
}
int osfm_ba_set_rig_camera_priors(osfm_ba* ba, const double* prior6, const double* sigma6) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  auto& b = ba->impl;
  const size_t n = b.rc_const.size();
  if (!prior6 || !sigma6) { b.rc_prior.clear(); b.rc_prior_sigma.clear(); }
  else { b.rc_prior.assign(prior6, prior6 + 6 * n); b.rc_prior_sigma.assign(sigma6, sigma6 + 6 * n); }
  OSFM_API_END
}
int osfm_ba_set_point_priors(osfm_ba* ba, int n, const int32_t* point, const double* prior3, const double* sigma3,
                             const int32_t* has_altitude) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!point || !prior3 || !sigma3 || !has_altitude))) throw ArgError("bad point prior arrays");
  auto& b = ba->impl;
  b.pp_point.assign(point, point + n);
  b.pp_prior.assign(prior3, prior3 + 3 * (size_t)n);
  b.pp_sigma.assign(sigma3, sigma3 + 3 * (size_t)n);
  b.pp_alt.assign(has_altitude, has_altitude + n);
  OSFM_API_END
}
int osfm_ba_set_ext_blocks(osfm_ba* ba, int n, const int32_t* size, const double* values, const int32_t* constant,
                           const double* lower_bound) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!size || !values || !constant || !lower_bound))) throw ArgError("bad ext block arrays");
  auto& b = ba->impl;
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (size[i] < 1 || size[i] > 16) throw ArgError("ext block size must be in [1, 16]");
    total += (size_t)size[i];
  }
  b.ext_size.assign(size, size + n);
  b.ext_const.assign(constant, constant + n);
  b.ext_values.assign(values, values + total);
  b.ext_lower.assign(lower_bound, lower_bound + total);
  OSFM_API_END
}
int osfm_ba_get_ext_blocks(osfm_ba* ba, double* values) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  std::copy(ba->impl.ext_values.begin(), ba->impl.ext_values.end(), values);
  OSFM_API_END
}
int osfm_ba_set_side_terms(osfm_ba* ba, int n, const osfm_side_term* terms, int64_t nconsts, const double* consts) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || nconsts < 0 || (n > 0 && !terms) || (nconsts > 0 && !consts)) throw ArgError("bad side term arrays");
  auto& b = ba->impl;
  b.side_terms.assign(terms, terms + n);
  b.side_consts.assign(consts, consts + nconsts);
  for (const auto& t : b.side_terms)
    if (t.loss < -1 || t.loss > OSFM_LOSS_TUKEY) throw ArgError("ceres::LossFunction with that name not found.");
  OSFM_API_END
}
int osfm_ba_set_shots(osfm_ba* ba, int n, const int32_t* rig_instance, const int32_t* camera,
                      const int32_t* rig_camera, const int32_t* use_rig_camera) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!rig_instance || !camera || !rig_camera || !use_rig_camera))) throw ArgError("bad shot arrays");
  ba->impl.shot_inst.assign(rig_instance, rig_instance + n);
  ba->impl.shot_cam.assign(camera, camera + n);
  ba->impl.shot_rc.assign(rig_camera, rig_camera + n);
  ba->impl.shot_use_rc.assign(use_rig_camera, use_rig_camera + n);
  OSFM_API_END
}
int osfm_ba_set_points(osfm_ba* ba, int n, const double* xyz, const int32_t* constant) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!xyz || !constant))) throw ArgError("bad point arrays");
  ba->impl.pts.assign(xyz, xyz + 3 * (size_t)n);
  ba->impl.pt_const.assign(constant, constant + n);
  OSFM_API_END
}
int osfm_ba_set_observations(osfm_ba* ba, int64_t n, const int32_t* shot, const int32_t* point, const double* xy,
                             const double* std_deviation) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!shot || !point || !xy || !std_deviation))) throw ArgError("bad observation arrays");
  // straight to the device: the ordering, the index checks and the 1/sigma happen there (ba_order.cuh)
  auto& b = ba->impl;
  OSFM_CUDA(cudaSetDevice(b.device));
  if (b.obs_pending) { OSFM_CUDA(cudaStreamSynchronize(b.copy_stream)); b.obs_pending = false; }
  const size_t nz = (size_t)std::max<int64_t>(n, 1);
  b.d_raw_shot.reserve(nz); b.d_raw_point.reserve(nz); b.d_raw_xy.reserve(2 * nz); b.d_raw_sigma.reserve(nz);
  if (n > 0) {
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_shot.p, shot, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_point.p, point, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_xy.p, xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_sigma.p, std_deviation, sizeof(double) * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaStreamSynchronize(b.own_stream));
  }
  b.n_obs_full = n;
  b.has_run = false;
  OSFM_API_END
}
int osfm_ba_set_observations_async(osfm_ba* ba, int64_t n, const int32_t* shot, const int32_t* point, const double* xy,
                                   const double* std_deviation) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (n < 0 || (n > 0 && (!shot || !point || !xy || !std_deviation))) throw ArgError("bad observation arrays");
  auto& b = ba->impl;
  OSFM_CUDA(cudaSetDevice(b.device));
  if (b.obs_pending) { OSFM_CUDA(cudaStreamSynchronize(b.copy_stream)); b.obs_pending = false; }
  const size_t nz = (size_t)std::max<int64_t>(n, 1);
  b.d_raw_shot.reserve(nz); b.d_raw_point.reserve(nz); b.d_raw_xy.reserve(2 * nz); b.d_raw_sigma.reserve(nz);
  if (n > 0) {
    // the indices are what the ordering needs first: they are on the device when the call returns; the measurements
    // follow on the copy stream and run() waits for them (an event) right before the kernel that gathers them
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_shot.p, shot, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_point.p, point, sizeof(int32_t) * n, cudaMemcpyHostToDevice, b.own_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_xy.p, xy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, b.copy_stream));
    OSFM_CUDA(cudaMemcpyAsync(b.d_raw_sigma.p, std_deviation, sizeof(double) * n, cudaMemcpyHostToDevice, b.copy_stream));
    OSFM_CUDA(cudaEventRecord(b.ev_obs, b.copy_stream));
    b.obs_pending = true;
    OSFM_CUDA(cudaStreamSynchronize(b.own_stream));
  }
  b.n_obs_full = n;
  b.has_run = false;
  OSFM_API_END
}
int osfm_ba_set_options(osfm_ba* ba, int loss, double loss_threshold, int max_iterations, const char* linear_solver,
                        int compute_reprojection_errors) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (loss < 0 || loss > 4) throw ArgError("ceres::LossFunction with that name not found.");  // bundle_adjuster.cc:427
  if (linear_solver) {
    const std::string s(linear_solver);
    // ceres::StringToLinearSolverType names (bundle_adjuster.cc:1105-1109)
    static const char* known[] = {"DENSE_NORMAL_CHOLESKY", "DENSE_QR", "SPARSE_NORMAL_CHOLESKY", "DENSE_SCHUR",
                                  "SPARSE_SCHUR", "ITERATIVE_SCHUR", "CGNR"};
    bool found = false;
    for (const char* k : known) found |= (s == k);
    if (!found) throw std::runtime_error("Linear solver type " + s + " doesn't exist.");
  }
  ba->impl.loss = loss;
  ba->impl.loss_a = loss_threshold;
  ba->impl.max_iterations = max_iterations;
  ba->impl.compute_reproj = compute_reprojection_errors != 0;
  OSFM_API_END
}
int osfm_ba_set_distributed(osfm_ba* ba, int rank, int world, osfm_allreduce_fn fn, void* user) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (world < 1 || rank < 0 || rank >= world) throw ArgError("bad rank/world");
  if (world > 1 && !fn && !(ba->impl.nccl_comm && ba->impl.nccl_world == world && ba->impl.nccl_rank == rank))
    throw ArgError("world > 1 needs an all-reduce callback or osfm_ba_set_nccl");
  ba->impl.rank = rank; ba->impl.world = world; ba->impl.allreduce = fn; ba->impl.allreduce_user = user;
  OSFM_API_END
}
int osfm_nccl_unique_id(char* out128) {
  OSFM_API_BEGIN
  if (!out128) throw ArgError("null id buffer");
  OsfmNcclId id;
  osfm::nccl_check(osfm::nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, id.internal, 128);
  OSFM_API_END
}
int osfm_ba_set_nccl(osfm_ba* ba, int rank, int world, const char* id128) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (world < 2 || rank < 0 || rank >= world || !id128) throw ArgError("bad rank/world/id");
  auto& b = ba->impl;
  OSFM_CUDA(cudaSetDevice(b.device));
  if (b.nccl_comm) { osfm::nccl_api().CommDestroy(b.nccl_comm); b.nccl_comm = nullptr; }
  OsfmNcclId id;
  std::memcpy(id.internal, id128, 128);
  osfm::nccl_check(osfm::nccl_api().CommInitRank(&b.nccl_comm, world, id, rank), "ncclCommInitRank");
  b.nccl_rank = rank; b.nccl_world = world;
  OSFM_API_END
}
int osfm_ba_set_stream(osfm_ba* ba, void* cuda_stream) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  ba->impl.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ba->impl.own_stream;
  OSFM_API_END
}
int osfm_ba_run(osfm_ba* ba) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  try {
    ba->impl.run();
  } catch (...) {
    // an asynchronous observation upload must not outlive the call: the caller may free its arrays now
    if (ba->impl.obs_pending) { cudaStreamSynchronize(ba->impl.copy_stream); ba->impl.obs_pending = false; }
    throw;
  }
  OSFM_API_END
}
int osfm_ba_get_summary(osfm_ba* ba, osfm_ba_summary* out) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (!out) throw ArgError("null out");
  *out = ba->impl.summary;
  OSFM_API_END
}
int osfm_ba_get_cameras(osfm_ba* ba, double* params_flat) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  std::copy(ba->impl.cam_params.begin(), ba->impl.cam_params.end(), params_flat);
  OSFM_API_END
}
int osfm_ba_get_rig_instances(osfm_ba* ba, double* pose6) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  std::copy(ba->impl.inst.begin(), ba->impl.inst.end(), pose6);
  OSFM_API_END
}
int osfm_ba_get_rig_cameras(osfm_ba* ba, double* pose6) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  std::copy(ba->impl.rc.begin(), ba->impl.rc.end(), pose6);
  OSFM_API_END
}
int osfm_ba_get_points(osfm_ba* ba, double* xyz) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  std::copy(ba->impl.pts.begin(), ba->impl.pts.end(), xyz);
  OSFM_API_END
}
int osfm_ba_get_reprojection_errors(osfm_ba* ba, double* out_n_by_3) {
  OSFM_API_BEGIN
  OSFM_BA_CHECK
  if (!ba->impl.has_run) throw std::runtime_error("run() has not been called");
  auto& b = ba->impl;
  OSFM_CUDA(cudaSetDevice(b.device));
  if (b.n_obs_full > 0) {
    if (b.reproj_valid)
      OSFM_CUDA(cudaMemcpy(out_n_by_3, b.d_reproj.p, sizeof(double) * 3 * (size_t)b.n_obs_full, cudaMemcpyDeviceToHost));
    else
      std::fill(out_n_by_3, out_n_by_3 + 3 * (size_t)b.n_obs_full, 0.0);
  }
  OSFM_API_END
}

int osfm_ba_eval_observation(int device, int projection_type, const double* camera, const double* rig_instance,
                             const double* rig_camera, int use_rig_camera, const double* point,
                             const double* observed, double std_deviation, double* r, double* jac_camera,
                             double* jac_instance, double* jac_rig_camera, double* jac_point, int* num_residuals) {
  OSFM_API_BEGIN
  if (projection_type < 0 || projection_type > 9) throw ArgError("Invalid ProjectionType");
  OSFM_CUDA(cudaSetDevice(device));
  const int C = osfm::model_num_params(projection_type);
  double in[34] = {0};
  for (int i = 0; i < C; ++i) in[i] = camera[i];
  for (int i = 0; i < 6; ++i) in[16 + i] = rig_instance[i];
  for (int i = 0; i < 6; ++i) in[22 + i] = rig_camera ? rig_camera[i] : 0.0;
  for (int i = 0; i < 3; ++i) in[28 + i] = point[i];
  in[31] = observed[0]; in[32] = observed[1]; in[33] = 1.0 / std_deviation;
  double *d_in = nullptr, *d_out = nullptr;
  int* d_n = nullptr;
  OSFM_CUDA(cudaMalloc(&d_in, sizeof(in)));
  OSFM_CUDA(cudaMalloc(&d_out, sizeof(double) * 96));
  OSFM_CUDA(cudaMalloc(&d_n, sizeof(int)));
  OSFM_CUDA(cudaMemcpy(d_in, in, sizeof(in), cudaMemcpyHostToDevice));
  osfm::ba_eval_one<<<1, 1>>>(projection_type, d_in, use_rig_camera, d_out, d_n);
  OSFM_LAUNCH_CHECK();
  double out[96];
  int nres = 0;
  OSFM_CUDA(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
  OSFM_CUDA(cudaMemcpy(&nres, d_n, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_n);
  for (int i = 0; i < nres; ++i) r[i] = out[i];
  for (int i = 0; i < nres * C; ++i) jac_camera[i] = out[3 + i];
  for (int i = 0; i < nres * 6; ++i) jac_instance[i] = out[51 + i];
  for (int i = 0; i < nres * 6; ++i) jac_rig_camera[i] = out[69 + i];
  for (int i = 0; i < nres * 3; ++i) jac_point[i] = out[87 + i];
  if (num_residuals) *num_residuals = nres;
  OSFM_API_END
}

}  // extern "C"
