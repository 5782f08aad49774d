/* opensfm_b200 — C ABI of the B200-native hot paths of OpenSfM.
 *
 * Plain C, plain pointers and sizes; no torch / pybind types.  Every entry
 * point returns 0 on success and a non-zero code on failure; the message of
 * the last failure on the calling thread is osfm_last_error().
 *
 * Two paths:
 *   MATCH  brute-force descriptor matching + Lowe ratio test, replacing
 *          opensfm/matching.py:723-777 (cv2.BFMatcher.knnMatch k=2).
 *   BA     bundle adjustment, replacing bundle::BundleAdjuster
 *          (opensfm/src/bundle/bundle_adjuster.h:178-299,
 *           opensfm/src/bundle/src/bundle_adjuster.cc) behind pybundle /
 *          sfm::BAHelpers (opensfm/src/sfm/src/ba_helpers.cc:117,408,581).
 *
 * There is no CPU fallback: every call needs a CUDA device (sm_100a).
 */
#ifndef OPENSFM_B200_H_
#define OPENSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSFM_OK 0
#define OSFM_ERR_CUDA 1
#define OSFM_ERR_ARG 2
#define OSFM_ERR_RUNTIME 3

const char* osfm_last_error(void);
int osfm_version(void);
/* number of this library's kernels launched by the calling process so far */
int64_t osfm_kernel_launch_count(void);

/* ------------------------------------------------------------------------
 * MATCH
 * ---------------------------------------------------------------------- */
typedef struct osfm_matcher osfm_matcher;

/* A matcher owns one CUDA stream and its workspaces on `device`.  One matcher
 * per host thread: matching.match_brute_force is called concurrently from
 * joblib threads (opensfm/context.py:59-64). */
int osfm_matcher_create(int device, osfm_matcher** out);
int osfm_matcher_destroy(osfm_matcher* m);

/* One-shot, host buffers: the drop-in for
 *   matching.match_brute_force(f1, f2, config, maskij)            (matching.py:723-756)
 *   matching.match_brute_force_symmetric(fi, fj, config, maskij)  (matching.py:759-777)
 * f1: n1 x dim, f2: n2 x dim, row-major, float32 (L2, "BruteForce") or
 * uint8 (Hamming, "BruteForce-Hamming", dim = bytes per descriptor).
 * mask: NULL or n1 x n2 bytes, non-zero = allowed (matching.py:745).
 * out_match[i] = matched index in f2 for row i of f1, or -1.  For the one-way
 * call the reference's list is [(i, out_match[i]) for i ascending if >= 0]. */
int osfm_bf_match_f32(osfm_matcher* m, const float* f1, int n1, const float* f2, int n2, int dim,
                      double lowes_ratio, const uint8_t* mask, int symmetric, int32_t* out_match);
int osfm_bf_match_u8(osfm_matcher* m, const uint8_t* f1, int n1, const uint8_t* f2, int n2, int nbytes,
                     double lowes_ratio, const uint8_t* mask, int symmetric, int32_t* out_match);

/* Resident descriptor sets + batched pair list: the drop-in for the pair loop of
 * matching.match_images_with_pairs (matching.py:63-98).  Descriptors are
 * uploaded once; a pair list is matched in one submission. */
int osfm_matcher_add_f32(osfm_matcher* m, const float* desc, int n, int dim, int* out_id);
int osfm_matcher_add_u8(osfm_matcher* m, const uint8_t* desc, int n, int nbytes, int* out_id);
/* Upload `count` descriptor matrices (desc[i] is n[i] x dim) with one host synchronisation at the end:
 * the per-image loop of matching.match_images_with_pairs loading features (matching.py:70-82). */
int osfm_matcher_add_batch_f32(osfm_matcher* m, int count, const float* const* desc, const int* n, int dim, int* out_ids);
int osfm_matcher_add_batch_u8(osfm_matcher* m, int count, const uint8_t* const* desc, const int* n, int nbytes,
                              int* out_ids);
/* uint8-STORED descriptors compared with L2 ("BruteForce"): HAHOG / SIFT are integers 0..255 and live as uint8 on
 * disk (opensfm/features.py:169-170, 526-534); the reference widens them to float32 on load.  Uploading the bytes
 * and widening on the device moves a quarter of the data; results are identical to adding the float32 matrix. */
int osfm_matcher_add_u8_l2(osfm_matcher* m, const uint8_t* desc, int n, int dim, int* out_id);
int osfm_matcher_add_batch_u8_l2(osfm_matcher* m, int count, const uint8_t* const* desc, const int* n, int dim,
                                 int* out_ids);
int osfm_matcher_remove(osfm_matcher* m, int id);
int osfm_matcher_clear(osfm_matcher* m);
/* Enqueue matching of npairs pairs (ids_a[p], ids_b[p]).  Results stay on the
 * device until fetched.  Total result length = sum_p n(ids_a[p]). */
int osfm_matcher_match_pairs_async(osfm_matcher* m, int npairs, const int* ids_a, const int* ids_b,
                                   double lowes_ratio, int symmetric);
/* Guided matching (matching._match_descriptors_guided_impl, matching.py:260-338): unit bearing vectors of a
 * resident set (n x 3 float32, feature_loader.load_bearings), then a pair list with the relative pose of image b
 * w.r.t. image a -- pose12[p] = [R (cam b -> cam a) row-major, 9 | origin of b in a, 3] as doubles
 * (pose.get_R_cam_to_world(), pose.get_origin()).  The epipolar mask
 * compute_inliers_bearing_epipolar(b1, b2, pose, threshold) (matching.py:847-868,
 * geometry/src/triangulation.cc:195-219) is evaluated on the device into a bitmask that never leaves HBM, and
 * the pairs are matched like osfm_matcher_match_pairs_async with that mask (transposed for the b -> a pass). */
int osfm_matcher_set_bearings(osfm_matcher* m, int id, const float* bearings_n_by_3);
int osfm_matcher_match_pairs_guided_async(osfm_matcher* m, int npairs, const int* ids_a, const int* ids_b,
                                          const double* pose12, double threshold, double lowes_ratio, int symmetric);
int osfm_matcher_sync(osfm_matcher* m);
/* Copy the last batch's results to the host: concatenated per pair, n(ids_a[p]) entries each. */
int osfm_matcher_fetch(osfm_matcher* m, int32_t* out_match, int64_t capacity);
/* The last batch as the reference's lists: (query, train) int32 rows packed pair after pair in query order
 * (matching.py:749-756, 775-777), compacted on the device.  offsets_out[npairs + 1] = first row of every pair,
 * *total_rows = offsets_out[npairs] rows written to pairs_out (capacity_rows >= sum of n(ids_a[p]) always fits). */
int osfm_matcher_fetch_pairs(osfm_matcher* m, int64_t* offsets_out, int32_t* pairs_out, int64_t capacity_rows,
                             int64_t* total_rows);
/* Milliseconds spent on the device by the last batch (CUDA events on the matcher's stream). */
int osfm_matcher_last_device_ms(osfm_matcher* m, float* ms_total, float* ms_distance_kernel);
/* 0 = pick automatically, 1 = force the exact SIMT kernel, 2 = force the tcgen05 kernel
 * (fails at match time if the descriptors are not exactly representable). */
int osfm_matcher_set_kernel(osfm_matcher* m, int which);
/* Which distance kernel the last batch used: 1 = SIMT, 2 = tcgen05 (L2), 3 = tcgen05 fp8 (Hamming). */
int osfm_matcher_last_kernel(osfm_matcher* m);
/* Device memory held by the matcher's descriptor slabs (bytes reserved / bytes in use by live sets). */
int osfm_matcher_device_bytes(osfm_matcher* m, int64_t* reserved, int64_t* in_use);

/* WORDS matcher: features::match_using_words (opensfm/src/features/src/matching.cc:24-88; pyfeatures, called by
 * matching.match_words, matching.py:636-656).  words1: n1 x words_per_feature nearest visual words of every feature
 * of image 1; words2: the nearest word of every feature of image 2 (words2[:, 0] in the reference's call).
 * out_match[i] = matched feature of image 2 or -1; the reference returns the rows (i, out_match[i]) with a match. */
int osfm_match_words(osfm_matcher* m, const float* f1, int n1, const int32_t* words1, int words_per_feature,
                     const float* f2, int n2, const int32_t* words2, int dim, float lowes_ratio, int max_checks,
                     int32_t* out_match);
/* features::compute_vlad_distances (matching.cc:122-145): Euclidean distance of VLAD descriptor `query` to each of
 * the n descriptors (n x dim float32, row-major); out_n[query] = 0. */
int osfm_vlad_distances(osfm_matcher* m, const float* vlad, int n, int dim, int query, double* out_n);

/* ------------------------------------------------------------------------
 * BA
 * ---------------------------------------------------------------------- */
typedef struct osfm_ba osfm_ba;

/* geometry::ProjectionType (opensfm/src/geometry/camera_instances.h:8-20) */
enum {
  OSFM_PERSPECTIVE = 0, OSFM_BROWN = 1, OSFM_FISHEYE = 2, OSFM_FISHEYE_OPENCV = 3, OSFM_FISHEYE62 = 4,
  OSFM_FISHEYE624 = 5, OSFM_SPHERICAL = 6, OSFM_DUAL = 7, OSFM_RADIAL = 8, OSFM_SIMPLE_RADIAL = 9
};
/* ceres loss names accepted by CreateLossFunction (bundle_adjuster.cc:414-429) */
enum { OSFM_LOSS_TRIVIAL = 0, OSFM_LOSS_HUBER = 1, OSFM_LOSS_SOFTLONE = 2, OSFM_LOSS_CAUCHY = 3, OSFM_LOSS_ARCTAN = 4,
       OSFM_LOSS_TUKEY = 5 /* side terms only (common position, bundle_adjuster.cc:905) */ };

int osfm_ba_create(int device, osfm_ba** out);
int osfm_ba_destroy(osfm_ba* ba);
int osfm_camera_num_params(int projection_type);

/* Bulk SoA setters — replace the string-keyed AddCamera / AddRigInstance /
 * AddRigCamera / AddPoint / AddPointProjectionObservation calls
 * (bundle_adjuster.cc:94-260).  All arrays are host memory and are copied. */
int osfm_ba_set_cameras(osfm_ba* ba, int n, const int32_t* type, const double* params /*flat*/,
                        const int32_t* constant, const double* prior /*flat*/, const double* prior_sigma /*flat*/,
                        const int32_t* prior_log /*flat*/);
int osfm_ba_set_rig_instances(osfm_ba* ba, int n, const double* pose6, const int32_t* constant,
                              const int32_t* has_position_prior, const double* prior_position3,
                              const double* prior_std3);
int osfm_ba_set_rig_cameras(osfm_ba* ba, int n, const double* pose6, const int32_t* constant);
int osfm_ba_set_shots(osfm_ba* ba, int n, const int32_t* rig_instance, const int32_t* camera,
                      const int32_t* rig_camera, const int32_t* use_rig_camera);
int osfm_ba_set_points(osfm_ba* ba, int n, const double* xyz, const int32_t* constant);
int osfm_ba_set_observations(osfm_ba* ba, int64_t n, const int32_t* shot, const int32_t* point,
                             const double* xy, const double* std_deviation);
/* The same for PAGE-LOCKED arrays: the indices are on the device when the call returns, xy and std_deviation may
 * still be in flight on a copy stream (osfm_ba_run waits for them right before the kernel that needs them, i.e. the
 * upload overlaps the device-side ordering): both arrays must stay valid and unchanged until osfm_ba_run returns.
 * With pageable memory the copies are staged by the driver and the call behaves like osfm_ba_set_observations. */
int osfm_ba_set_observations_async(osfm_ba* ba, int64_t n, const int32_t* shot, const int32_t* point,
                                   const double* xy, const double* std_deviation);

/* Rig-camera pose priors: DataPriorError<Pose> with sigma GetDefaultRigPoseSigma, one per rig camera
 * (bundle_adjuster.cc:779-790; residual dropped when the rig camera is constant).  prior6 / sigma6: n x 6
 * in the order [rx, ry, rz, tx, ty, tz]; NULL removes the priors.  Call after osfm_ba_set_rig_cameras. */
int osfm_ba_set_rig_camera_priors(osfm_ba* ba, const double* prior6, const double* sigma6);
/* Point priors (GCP): AddPointPrior (bundle_adjuster.cc:224-236, residual :688-708): residuals on x, y
 * (and z when has_altitude) with scale 1 / max(sigma, eps).  n priors on points point[i]. */
int osfm_ba_set_point_priors(osfm_ba* ba, int n, const int32_t* point, const double* prior3, const double* sigma3,
                             const int32_t* has_altitude);
/* Extra parameter blocks of the camera side: camera biases (7: [R | t | scale], data/bias.h:10-31), per-instance
 * reconstruction scales (1, lower bound 0, bundle_adjuster.cc:672-685), std-deviation scales of the position-prior
 * groups (1, lower bound 1e-10, :727-736).  values / lower_bound are flat over the blocks (-inf = unbounded). */
int osfm_ba_set_ext_blocks(osfm_ba* ba, int n, const int32_t* size, const double* values, const int32_t* constant,
                           const double* lower_bound);
int osfm_ba_get_ext_blocks(osfm_ba* ba, double* values);

/* Side terms: the O(#shots) residual blocks next to the point projections.  Each names up to 6 parameter blocks
 * (kind: 0 camera, 1 rig instance, 2 rig camera, 3 ext block; idx within the kind), a ceres loss (-1 = none) and
 * `nconst` constants starting at consts[cofs].  Types, blocks and constants (reference functor):
 *   UP_VECTOR           [inst, rigcam]  c = unit acceleration[3], 1/std          absolute_motion_errors.h:12-39
 *   PAN / TILT / ROLL   [inst, rigcam]  c = angle, 1/std                         :41-136
 *   RELATIVE_MOTION     [inst_i, inst_j, scale_i(, scale_j)]  c = Rts[7], scale_matrix[49] row-major,
 *                       observed_scale; aux[0] = block of scale_j (2 or 3)       relative_motion_errors.h:14-72
 *   RELATIVE_ROTATION   [inst_i, inst_j(, rigcam_i)(, rigcam_j)]  c = Rij[3], scale_matrix[9];
 *                       aux[0], aux[1] = rig-camera block of i, j or -1          :74-103
 *   COMMON_POSITION     same blocks / aux; c = margin, 1/std                     :105-138
 *   LINEAR_MOTION       [inst0, inst1, inst2(, rigcams)]  c = alpha, 1/pos_std, 1/ori_std; aux[0..2] = rig-camera
 *                       blocks or -1                                             motion_prior_errors.h:13-76
 *   TRANSLATION_PRIOR   [inst1, inst2]  c = max(prior norm, 1e-20)               absolute_motion_errors.h:180-202
 *   PARAMETER_BARRIER   [camera]  c = lower, upper; aux[0] = parameter index     parameters_errors.h:20-36
 *   STD_DEVIATION       [ext]                                                    parameters_errors.h:7-18
 *   POSITION_PRIOR      [inst, bias ext(7), std-scale ext(1)]  c = prior[3], 1/sigma[3], adjust flag
 *                       (the general form of the position prior: bias transform + scale group,
 *                        bundle_adjuster.cc:745-778, data/bias.h:33-53) */
enum {
  OSFM_SIDE_UP_VECTOR = 0, OSFM_SIDE_PAN = 1, OSFM_SIDE_TILT = 2, OSFM_SIDE_ROLL = 3, OSFM_SIDE_RELATIVE_MOTION = 4,
  OSFM_SIDE_RELATIVE_ROTATION = 5, OSFM_SIDE_COMMON_POSITION = 6, OSFM_SIDE_LINEAR_MOTION = 7,
  OSFM_SIDE_TRANSLATION_PRIOR = 8, OSFM_SIDE_PARAMETER_BARRIER = 9, OSFM_SIDE_STD_DEVIATION = 10,
  OSFM_SIDE_POSITION_PRIOR = 11,
  /* ReconstructionAlignment (opensfm/src/bundle/reconstruction_alignment.h:140-365): "shots" are rig-instance blocks
   * holding [R | t] world-to-camera, "reconstructions" are 7-parameter ext blocks [R | t | scale] (scale >= 0.1):
   *   RA_RELATIVE_MOTION             [reconstruction, shot]  c = Rtai[6], scale_matrix[36]
   *   RA_ABSOLUTE_POSITION           [shot]                  c = position[3], 1/std
   *   RA_RELATIVE_ABSOLUTE_POSITION  [reconstruction]        c = position[3], shot[6], 1/std
   *   RA_COMMON_POINT                [reconstruction a, b]   c = point_a[3], point_b[3], 1/std
   *   RA_COMMON_CAMERA               [reconstruction a, b]   c = shot_a[6], shot_b[6], 1/std_centre, 1/std_rotation */
  OSFM_SIDE_RA_RELATIVE_MOTION = 12, OSFM_SIDE_RA_ABSOLUTE_POSITION = 13, OSFM_SIDE_RA_RELATIVE_ABSOLUTE_POSITION = 14,
  OSFM_SIDE_RA_COMMON_POINT = 15, OSFM_SIDE_RA_COMMON_CAMERA = 16, OSFM_SIDE_NUM_TYPES = 17
};
typedef struct {
  int32_t type, nres, nblocks;
  int32_t kind[6], idx[6];
  int32_t loss;          /* OSFM_LOSS_*, or -1 for no loss function */
  double loss_a;
  int32_t cofs;          /* index of the term's first constant in `consts` */
  int32_t aux[4];
} osfm_side_term;
int osfm_ba_set_side_terms(osfm_ba* ba, int n, const osfm_side_term* terms, int64_t nconsts, const double* consts);
/* SetPointProjectionLossFunction / SetMaxNumIterations / SetLinearSolverType
 * (bundle_adjuster.cc:262-372).  linear_solver: "SPARSE_SCHUR", "DENSE_SCHUR",
 * "ITERATIVE_SCHUR" are all served by Schur elimination + PCG; unknown names fail
 * like ceres::StringToLinearSolverType (bundle_adjuster.cc:1105-1109). */
int osfm_ba_set_options(osfm_ba* ba, int loss, double loss_threshold, int max_iterations,
                        const char* linear_solver, int compute_reprojection_errors);
/* Multi-GPU: this process holds shard `rank` of `world` (observations are
 * sharded by point).  allreduce_sum(buf, count, user) must sum `count` doubles
 * at device pointer `buf` across ranks on `stream` (NCCL); NULL when world == 1. */
typedef int (*osfm_allreduce_fn)(void* device_buf, int64_t count, void* stream, void* user);
int osfm_ba_set_distributed(osfm_ba* ba, int rank, int world, osfm_allreduce_fn fn, void* user);
/* Alternative to the callback: the library opens libnccl.so.2 itself and owns a communicator.
 * Rank 0 calls osfm_nccl_unique_id, the caller broadcasts the 128 bytes by its own means, every rank calls
 * osfm_ba_set_nccl (collective), then osfm_ba_set_distributed(ba, rank, world, NULL, NULL).  The all-reduces
 * of a run are then plain ncclAllReduce calls on the library's stream (no host code in between). */
int osfm_nccl_unique_id(char* out128);
int osfm_ba_set_nccl(osfm_ba* ba, int rank, int world, const char* id128);
/* Use an externally owned stream (e.g. torch's current stream); NULL = own stream. */
int osfm_ba_set_stream(osfm_ba* ba, void* cuda_stream);

/* BundleAdjuster::Run (bundle_adjuster.cc:595-1121). */
int osfm_ba_run(osfm_ba* ba);

typedef struct {
  int iterations;             /* LM iterations executed (successful + unsuccessful) */
  int successful_steps;
  int linear_solves;
  int pcg_iterations;         /* total CG iterations */
  int termination;            /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
  double initial_cost, final_cost;
  double time_run_s;          /* wall time of run() (ba_helpers.cc:749-753) */
  double time_device_ms;      /* CUDA-event time of the LM loop */
  double time_linearize_ms;   /* summed CUDA-event time of the linearise+accumulate kernel */
  int64_t linearize_launches;
  double time_schur_ms;       /* summed CUDA-event time of the Schur-complement kernel */
  int64_t schur_launches;
  double time_pcg_ms;         /* summed CUDA-event time of the PCG solves */
  double time_backsub_ms;
  int64_t num_observations_local; /* observations held by this rank */
  int reduced_dim;            /* dimension of the reduced camera system */
  int reduced_blocks;         /* stored blocks of the block-sparse reduced system (both triangles) */
  int64_t reduced_nnz;        /* stored doubles of the reduced system */
  int jac_planes;             /* doubles stored per observation: nres * (wc + 3 + 1) */
  int64_t kernel_launches;
  char message[128];
} osfm_ba_summary;
int osfm_ba_get_summary(osfm_ba* ba, osfm_ba_summary* out);

int osfm_ba_get_cameras(osfm_ba* ba, double* params_flat);
int osfm_ba_get_rig_instances(osfm_ba* ba, double* pose6);
int osfm_ba_get_rig_cameras(osfm_ba* ba, double* pose6);
int osfm_ba_get_points(osfm_ba* ba, double* xyz);
/* ComputeReprojectionErrors (bundle_adjuster.cc:1196-1208): unscaled residuals,
 * n x 3 (third column 0 for 2-D errors), observation order. */
int osfm_ba_get_reprojection_errors(osfm_ba* ba, double* out_n_by_3);

/* Single-observation evaluation on the device (test hook for the per-observation
 * kernel; ReprojectionError2DAnalytic::Evaluate, projection_errors.h:67-207).
 * Outputs: r[3], jac_camera[3*C], jac_instance[18], jac_rig_camera[18], jac_point[9]. */
int osfm_ba_eval_observation(int device, int projection_type, const double* camera, const double* rig_instance,
                             const double* rig_camera, int use_rig_camera, const double* point,
                             const double* observed, double std_deviation, double* r, double* jac_camera,
                             double* jac_instance, double* jac_rig_camera, double* jac_point, int* num_residuals);

#ifdef __cplusplus
}
#endif
#endif /* OPENSFM_B200_H_ */
