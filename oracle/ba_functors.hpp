// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement (fp64, scalar C++) of the per-observation arithmetic of the
// reference's bundle-adjustment hot loop.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg may use anything in
// oracle/.  The CUDA engine in opensfm_b200/csrc never includes this file.
//
// Every function cites the reference file:line (relative to the OpenSfM tree)
// whose algorithm it restates.  The forward maps are templated on the scalar so
// that the same formula can be evaluated with `Dual` numbers: that reproduces
// the reference's own test strategy (analytic Jacobian == autodiff of the
// templated functor, opensfm/src/bundle/test/reprojection_errors_test.cc:52-80).
//
// Parity pin: see oracle/README.md.  The reference cannot be compiled in this
// image (Eigen/Ceres absent), so the pin is (a) the reference's golden pixels
// (geometry/test/camera_test.cc:119-172), (b) analytic==autodiff on the
// reference's own test inputs at the reference's own tolerance (1e-14), and
// (c) project(bearing(x)) == x.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstring>
#include <stdexcept>

namespace oracle {

// ---------------------------------------------------------------------------
// Forward-mode dual number with a runtime number of directions (<= MAXD).
// Stands in for Eigen::AutoDiffScalar<VectorXd> used by the reference tests.
// ---------------------------------------------------------------------------
constexpr int MAXD = 40;
struct Dual {
  double v;
  double d[MAXD];
  Dual() : v(0.0) { std::memset(d, 0, sizeof(d)); }
  Dual(double x) : v(x) { std::memset(d, 0, sizeof(d)); }  // NOLINT implicit
  static Dual variable(double x, int idx) {
    Dual r(x);
    r.d[idx] = 1.0;
    return r;
  }
};
#define ORACLE_DUAL_BIN(op, val, da, db)                           \
  inline Dual operator op(const Dual& a, const Dual& b) {          \
    Dual r;                                                        \
    r.v = val;                                                     \
    for (int i = 0; i < MAXD; ++i) r.d[i] = (da)*a.d[i] + (db)*b.d[i]; \
    return r;                                                      \
  }
ORACLE_DUAL_BIN(+, a.v + b.v, 1.0, 1.0)
ORACLE_DUAL_BIN(-, a.v - b.v, 1.0, -1.0)
ORACLE_DUAL_BIN(*, a.v* b.v, b.v, a.v)
ORACLE_DUAL_BIN(/, a.v / b.v, 1.0 / b.v, -a.v / (b.v * b.v))
#undef ORACLE_DUAL_BIN
inline Dual operator-(const Dual& a) {
  Dual r;
  r.v = -a.v;
  for (int i = 0; i < MAXD; ++i) r.d[i] = -a.d[i];
  return r;
}
inline bool operator>(const Dual& a, const Dual& b) { return a.v > b.v; }
inline bool operator<(const Dual& a, const Dual& b) { return a.v < b.v; }
inline Dual unary(const Dual& a, double val, double dval) {
  Dual r;
  r.v = val;
  for (int i = 0; i < MAXD; ++i) r.d[i] = dval * a.d[i];
  return r;
}
inline Dual sqrt(const Dual& a) {
  const double s = std::sqrt(a.v);
  return unary(a, s, 0.5 / s);
}
inline Dual sin(const Dual& a) { return unary(a, std::sin(a.v), std::cos(a.v)); }
inline Dual cos(const Dual& a) { return unary(a, std::cos(a.v), -std::sin(a.v)); }
inline Dual log(const Dual& a) { return unary(a, std::log(a.v), 1.0 / a.v); }
inline Dual atan2(const Dual& y, const Dual& x) {
  Dual r;
  r.v = std::atan2(y.v, x.v);
  const double den = x.v * x.v + y.v * y.v;
  for (int i = 0; i < MAXD; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / den;
  return r;
}
using std::atan2;
using std::cos;
using std::log;
using std::sin;
using std::sqrt;

// ---------------------------------------------------------------------------
// Camera model ids == geometry::ProjectionType (camera_instances.h:8-20).
// ---------------------------------------------------------------------------
enum ProjectionType {
  PERSPECTIVE = 0,
  BROWN = 1,
  FISHEYE = 2,
  FISHEYE_OPENCV = 3,
  FISHEYE62 = 4,
  FISHEYE624 = 5,
  SPHERICAL = 6,
  DUAL = 7,
  RADIAL = 8,
  SIMPLE_RADIAL = 9,
};

// Stage ids.  A camera is PROJ -> DISTO -> AFFINE (camera_instances.h:132-193);
// parameters are stored in that order [PROJ | DISTO | AFF] (functions.h:58-67,
// geometry/src/camera.cc:9-178).
enum StageId {
  S_POSE,
  S_NORMALIZE,
  S_PROJ_PERSPECTIVE,
  S_PROJ_FISHEYE,
  S_PROJ_DUAL,
  S_PROJ_SPHERICAL,
  S_DISTO_2,
  S_DISTO_24,
  S_DISTO_2468,
  S_DISTO_62,
  S_DISTO_624,
  S_DISTO_BROWN,
  S_AFF_AFFINE,
  S_AFF_UNIFORM,
  S_AFF_IDENTITY,
};

struct StageShape {
  int in, np, out;
};
inline StageShape shape_of(StageId s) {
  switch (s) {
    case S_POSE: return {3, 6, 3};
    case S_NORMALIZE: return {3, 0, 3};
    case S_PROJ_PERSPECTIVE: return {3, 0, 2};
    case S_PROJ_FISHEYE: return {3, 0, 2};
    case S_PROJ_DUAL: return {3, 1, 2};
    case S_PROJ_SPHERICAL: return {3, 0, 2};
    case S_DISTO_2: return {2, 1, 2};
    case S_DISTO_24: return {2, 2, 2};
    case S_DISTO_2468: return {2, 4, 2};
    case S_DISTO_62: return {2, 8, 2};
    case S_DISTO_624: return {2, 12, 2};
    case S_DISTO_BROWN: return {2, 5, 2};
    case S_AFF_AFFINE: return {2, 4, 2};
    case S_AFF_UNIFORM: return {2, 1, 2};
    case S_AFF_IDENTITY: return {2, 0, 2};
  }
  throw std::runtime_error("bad stage");
}

// camera_instances.h:181-193 (the ten aliases) and :198-235 (Dispatch).
inline void camera_stages(int type, StageId out[3]) {
  switch (type) {
    case PERSPECTIVE: out[0] = S_PROJ_PERSPECTIVE; out[1] = S_DISTO_24; out[2] = S_AFF_UNIFORM; return;
    case BROWN: out[0] = S_PROJ_PERSPECTIVE; out[1] = S_DISTO_BROWN; out[2] = S_AFF_AFFINE; return;
    case FISHEYE: out[0] = S_PROJ_FISHEYE; out[1] = S_DISTO_24; out[2] = S_AFF_UNIFORM; return;
    case FISHEYE_OPENCV: out[0] = S_PROJ_FISHEYE; out[1] = S_DISTO_2468; out[2] = S_AFF_AFFINE; return;
    case FISHEYE62: out[0] = S_PROJ_FISHEYE; out[1] = S_DISTO_62; out[2] = S_AFF_AFFINE; return;
    case FISHEYE624: out[0] = S_PROJ_FISHEYE; out[1] = S_DISTO_624; out[2] = S_AFF_AFFINE; return;
    case SPHERICAL: out[0] = S_PROJ_SPHERICAL; out[1] = S_AFF_IDENTITY; out[2] = S_AFF_IDENTITY; return;
    case DUAL: out[0] = S_PROJ_DUAL; out[1] = S_DISTO_24; out[2] = S_AFF_UNIFORM; return;
    case RADIAL: out[0] = S_PROJ_PERSPECTIVE; out[1] = S_DISTO_24; out[2] = S_AFF_AFFINE; return;
    case SIMPLE_RADIAL: out[0] = S_PROJ_PERSPECTIVE; out[1] = S_DISTO_2; out[2] = S_AFF_AFFINE; return;
  }
  throw std::runtime_error("Invalid ProjectionType");  // camera_instances.h:232
}

// Number of stored camera parameters (SizeTraits, camera_instances.h:121-128:
// max(1, sum of the three stages' sizes); spherical stores one dummy value).
inline int camera_num_params(int type) {
  StageId st[3];
  camera_stages(type, st);
  int n = shape_of(st[0]).np + shape_of(st[1]).np + shape_of(st[2]).np;
  return n < 1 ? 1 : n;
}

// ---------------------------------------------------------------------------
// Forward maps (templated).
// ---------------------------------------------------------------------------

// geometry::PoseFunctor::Forward, transformations_functions.h:113-144.
// rt = [angle-axis of R_cam->world | camera origin]; x_cam = R(-r) (X - t).
template <class T>
void pose_forward(const T* point, const T* rt, T* out) {
  const T x = point[0] - rt[3];
  const T y = point[1] - rt[4];
  const T z = point[2] - rt[5];
  const T a = -rt[0];
  const T b = -rt[1];
  const T c = -rt[2];
  const T cp_x = b * z - c * y;
  const T cp_y = c * x - a * z;
  const T cp_z = a * y - b * x;
  const T theta2 = a * a + b * b + c * c;
  if (theta2 > T(DBL_EPSILON)) {
    const T theta = sqrt(theta2);
    const T cos_theta = cos(theta);
    const T sin_theta = sin(theta) / theta;
    const T dot_pt_p = (a * x + b * y + c * z) * (T(1.0) - cos_theta) / theta2;
    const T ox = x * cos_theta + sin_theta * cp_x + a * dot_pt_p;
    const T oy = y * cos_theta + sin_theta * cp_y + b * dot_pt_p;
    const T oz = z * cos_theta + sin_theta * cp_z + c * dot_pt_p;
    out[0] = ox; out[1] = oy; out[2] = oz;
  } else {
    const T ox = x + cp_x, oy = y + cp_y, oz = z + cp_z;
    out[0] = ox; out[1] = oy; out[2] = oz;
  }
}

// PoseFunctor::AngleAxisToRotation, transformations_functions.h:217-262
// (row-major R from an angle-axis vector).
template <class T>
void angle_axis_to_rotation(const T* aa, T* R) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 < T(DBL_EPSILON)) {
    // R = I + [aa]x
    R[0] = T(1.0); R[1] = -aa[2]; R[2] = aa[1];
    R[3] = aa[2]; R[4] = T(1.0); R[5] = -aa[0];
    R[6] = -aa[1]; R[7] = aa[0]; R[8] = T(1.0);
    return;
  }
  const T theta = sqrt(theta2);
  const T c = cos(theta);
  const T s = sin(theta);
  const T t = T(1.0) - c;
  const T inv_theta2 = T(1.0) / theta2;
  const T inv_theta = T(1.0) / theta;
  const T xx = aa[0] * aa[0] * inv_theta2;
  const T xy = aa[0] * aa[1] * inv_theta2;
  const T xz = aa[0] * aa[2] * inv_theta2;
  const T yy = aa[1] * aa[1] * inv_theta2;
  const T yz = aa[1] * aa[2] * inv_theta2;
  const T zz = aa[2] * aa[2] * inv_theta2;
  const T xs = aa[0] * inv_theta * s;
  const T ys = aa[1] * inv_theta * s;
  const T zs = aa[2] * inv_theta * s;
  R[0] = t * xx + c; R[1] = t * xy - zs; R[2] = t * xz + ys;
  R[3] = t * xy + zs; R[4] = t * yy + c; R[5] = t * yz - xs;
  R[6] = t * xz - ys; R[7] = t * yz + xs; R[8] = t * zz + c;
}

// geometry::Normalize::Forward, transformations_functions.h:266-272.
template <class T>
void normalize_forward(const T* p, T* out) {
  const T inv_norm = T(1.0) / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  for (int i = 0; i < 3; ++i) out[i] = p[i] * inv_norm;
}

// PerspectiveProjection::Forward, camera_projections_functions.h:89-93.
template <class T>
void perspective_forward(const T* p, T* out) {
  out[0] = p[0] / p[2];
  out[1] = p[1] / p[2];
}

// FisheyeProjection::Forward, camera_projections_functions.h:9-20.
template <class T>
void fisheye_forward(const T* p, T* out) {
  const T r = sqrt(p[0] * p[0] + p[1] * p[1]);
  if (r < T(1e-8)) {
    out[0] = p[0] / p[2];
    out[1] = p[1] / p[2];
    return;
  }
  const T theta = atan2(r, p[2]);
  out[0] = theta / r * p[0];
  out[1] = theta / r * p[1];
}

// DualProjection::Forward, camera_projections_functions.h:124-134.
template <class T>
void dual_forward(const T* p, const T* k, T* out) {
  T pp[2], pf[2];
  perspective_forward(p, pp);
  fisheye_forward(p, pf);
  out[0] = k[0] * pp[0] + (T(1.0) - k[0]) * pf[0];
  out[1] = k[0] * pp[1] + (T(1.0) - k[0]) * pf[1];
}

// SphericalProjection::Forward, camera_projections_functions.h:215-223.
template <class T>
void spherical_forward(const T* p, T* out) {
  const T lon = atan2(p[0], p[2]);
  const T lat = atan2(-p[1], sqrt(p[0] * p[0] + p[2] * p[2]));
  const T inv_norm = T(1.0 / (2.0 * M_PI));
  out[0] = lon * inv_norm;
  out[1] = -lat * inv_norm;
}

// Radial polynomials.  Disto2::Distortion (camera_distortions_functions.h:91-94),
// Disto24::Distortion (:193-196), Disto2468::Distortion (:312-316),
// Disto62::RadialDistortion (:480-485), DistoBrown::RadialDistortion (:830-834).
template <class T>
T radial_poly(const T& r2, const T* k, int n) {
  T acc = k[n - 1];
  for (int i = n - 2; i >= 0; --i) acc = k[i] + r2 * acc;
  return T(1.0) + r2 * acc;
}

// Disto62::TangentialDistortion (:487-492) == DistoBrown::TangentialDistortion (:835-841).
template <class T>
void tangential(const T& r2, const T& x, const T& y, const T& p1, const T& p2, T* out) {
  out[0] = T(2.0) * p1 * x * y + p2 * (r2 + T(2.0) * x * x);
  out[1] = T(2.0) * p2 * x * y + p1 * (r2 + T(2.0) * y * y);
}

// Forward of every distortion stage (camera_distortions_functions.h:16-21,
// :108-114, :211-218, :334-348, :512-537, :711-721).
template <class T>
void disto_forward(StageId s, const T* p, const T* k, T* out) {
  const T x = p[0], y = p[1];
  const T r2 = x * x + y * y;
  switch (s) {
    case S_DISTO_2: { const T d = radial_poly(r2, k, 1); out[0] = x * d; out[1] = y * d; return; }
    case S_DISTO_24: { const T d = radial_poly(r2, k, 2); out[0] = x * d; out[1] = y * d; return; }
    case S_DISTO_2468: { const T d = radial_poly(r2, k, 4); out[0] = x * d; out[1] = y * d; return; }
    case S_DISTO_62: {
      const T d = radial_poly(r2, k, 6);
      T t[2];
      tangential(r2, x, y, k[6], k[7], t);
      out[0] = x * d + t[0]; out[1] = y * d + t[1];
      return;
    }
    case S_DISTO_624: {
      const T d = radial_poly(r2, k, 6);
      T t[2];
      tangential(r2, x, y, k[6], k[7], t);
      // ThinPrismDistortion, :700-703
      const T tp0 = k[8] * r2 + k[9] * r2 * r2;
      const T tp1 = k[10] * r2 + k[11] * r2 * r2;
      out[0] = x * d + t[0] + tp0; out[1] = y * d + t[1] + tp1;
      return;
    }
    case S_DISTO_BROWN: {
      const T d = radial_poly(r2, k, 3);
      T t[2];
      tangential(r2, x, y, k[3], k[4], t);
      out[0] = x * d + t[0]; out[1] = y * d + t[1];
      return;
    }
    default: throw std::runtime_error("not a distortion stage");
  }
}

// Affine (transformations_functions.h:15-20), UniformScale (:55-59), Identity (:84-88).
template <class T>
void affine_forward(StageId s, const T* p, const T* a, T* out) {
  switch (s) {
    case S_AFF_AFFINE:
      out[0] = a[0] * p[0] + a[2];
      out[1] = a[0] * a[1] * p[1] + a[3];
      return;
    case S_AFF_UNIFORM:
      out[0] = a[0] * p[0];
      out[1] = a[0] * p[1];
      return;
    case S_AFF_IDENTITY:
      out[0] = p[0];
      out[1] = p[1];
      return;
    default: throw std::runtime_error("not an affine stage");
  }
}

template <class T>
void stage_forward(StageId s, const T* in, const T* prm, T* out) {
  switch (s) {
    case S_POSE: pose_forward(in, prm, out); return;
    case S_NORMALIZE: normalize_forward(in, out); return;
    case S_PROJ_PERSPECTIVE: perspective_forward(in, out); return;
    case S_PROJ_FISHEYE: fisheye_forward(in, out); return;
    case S_PROJ_DUAL: dual_forward(in, prm, out); return;
    case S_PROJ_SPHERICAL: spherical_forward(in, out); return;
    case S_DISTO_2: case S_DISTO_24: case S_DISTO_2468: case S_DISTO_62:
    case S_DISTO_624: case S_DISTO_BROWN:
      disto_forward(s, in, prm, out); return;
    default: affine_forward(s, in, prm, out); return;
  }
}

// ProjectGeneric::Forward (camera_instances.h:141-145): AFF(DISTO(PROJ(point))).
template <class T>
void camera_project(int type, const T* point, const T* params, T* out) {
  StageId st[3];
  camera_stages(type, st);
  T a[3], b[3];
  int off = 0;
  stage_forward(st[0], point, params + off, a);
  off += shape_of(st[0]).np;
  stage_forward(st[1], a, params + off, b);
  off += shape_of(st[1]).np;
  stage_forward(st[2], b, params + off, out);
}

// ---------------------------------------------------------------------------
// Analytic Jacobians of each stage (double only).  J is row-major,
// out x (in + np): [d/d in | d/d params].
// ---------------------------------------------------------------------------

// PoseFunctor::ForwardDerivatives<T,true>, transformations_functions.h:147-212.
// The reference obtains dR/d(angle-axis) with AutoDiffScalar<Vec3d> (:151-174)
// and chains it with d(R(x-t))/d(R, x, t) (:178-199); restated the same way
// with the Dual type above.  Output order [X | r | t] (:201-207).
inline void pose_jacobian(const double* point, const double* rt, double* out, double* J /*3x9*/) {
  Dual aa[3];
  for (int i = 0; i < 3; ++i) {
    aa[i] = Dual(-rt[i]);
    aa[i].d[i] = -1.0;  // :156-161: derivative of -r wrt r
  }
  Dual R[9];
  angle_axis_to_rotation(aa, R);
  const double xyz[3] = {point[0] - rt[3], point[1] - rt[4], point[2] - rt[5]};
  for (int i = 0; i < 3; ++i) {
    for (int k = 0; k < 3; ++k) {
      // d out_i / d r_k = sum_j dR_ij/dr_k * xyz_j
      double s = 0.0;
      for (int j = 0; j < 3; ++j) s += R[i * 3 + j].d[k] * xyz[j];
      J[i * 9 + 3 + k] = s;
      J[i * 9 + k] = R[i * 3 + k].v;       // d/dX = R
      J[i * 9 + 6 + k] = -R[i * 3 + k].v;  // d/dt = -R
    }
  }
  pose_forward(point, rt, out);
}

// Normalize::ForwardDerivatives, transformations_functions.h:274-304.
inline void normalize_jacobian(const double* p, double* out, double* J /*3x3*/) {
  const double x = p[0], y = p[1], z = p[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double norm2 = x2 + y2 + z2;
  const double norm = std::sqrt(norm2);
  const double inv_norm32 = 1.0 / (norm * norm2);
  J[0] = (y2 + z2) * inv_norm32; J[1] = (-x * y) * inv_norm32; J[2] = (-x * z) * inv_norm32;
  J[3] = (-y * x) * inv_norm32; J[4] = (x2 + z2) * inv_norm32; J[5] = (-y * z) * inv_norm32;
  J[6] = (-z * x) * inv_norm32; J[7] = (-z * y) * inv_norm32; J[8] = (x2 + y2) * inv_norm32;
  normalize_forward(p, out);
}

// PerspectiveProjection::ForwardDerivatives, camera_projections_functions.h:95-108.
inline void perspective_jacobian(const double* p, double* out, double* J, int stride) {
  J[0] = 1.0 / p[2]; J[1] = 0.0; J[2] = -p[0] / (p[2] * p[2]);
  J[stride] = 0.0; J[stride + 1] = J[0]; J[stride + 2] = -p[1] / (p[2] * p[2]);
  perspective_forward(p, out);
}

// FisheyeProjection::ForwardDerivatives, camera_projections_functions.h:22-68.
inline void fisheye_jacobian(const double* p, double* out, double* J, int stride) {
  const double r2 = p[0] * p[0] + p[1] * p[1];
  const double r = std::sqrt(r2);
  if (r < 1e-8) {
    const double inv_z = 1.0 / p[2], inv_z2 = inv_z * inv_z;
    J[0] = inv_z; J[1] = 0.0; J[2] = -p[0] * inv_z2;
    J[stride] = 0.0; J[stride + 1] = inv_z; J[stride + 2] = -p[1] * inv_z2;
    fisheye_forward(p, out);
    return;
  }
  const double R2 = r2 + p[2] * p[2];
  const double theta = std::atan2(r, p[2]);
  const double x2 = p[0] * p[0], y2 = p[1] * p[1], z2 = p[2] * p[2];
  const double inv_denom = 1.0 / (r2 * R2 * r);
  J[0] = (x2 * y2 * theta + y2 * y2 * theta + y2 * z2 * theta + x2 * p[2] * r) * inv_denom;
  J[1] = p[0] * (p[1] * p[2] * r - p[1] * theta * R2) * inv_denom;
  J[2] = -p[0] / R2;
  J[stride] = p[1] * (p[0] * p[2] * r - p[0] * theta * R2) * inv_denom;
  J[stride + 1] = (x2 * y2 * theta + x2 * x2 * theta + x2 * z2 * theta + y2 * p[2] * r) * inv_denom;
  J[stride + 2] = -p[1] / R2;
  fisheye_forward(p, out);
}

// DualProjection::ForwardDerivatives<T,true>, camera_projections_functions.h:136-174.
inline void dual_jacobian(const double* p, const double* k, double* out, double* J /*2x4*/) {
  double jp[6], jf[6], tmp[2];
  perspective_jacobian(p, tmp, jp, 3);
  fisheye_jacobian(p, tmp, jf, 3);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) J[i * 4 + j] = k[0] * jp[i * 3 + j] + (1.0 - k[0]) * jf[i * 3 + j];
  double pp[2], pf[2];
  perspective_forward(p, pp);
  fisheye_forward(p, pf);
  J[3] = pp[0] - pf[0];
  J[4 + 3] = pp[1] - pf[1];
  out[0] = k[0] * pp[0] + (1.0 - k[0]) * pf[0];
  out[1] = k[0] * pp[1] + (1.0 - k[0]) * pf[1];
}

// SphericalProjection::ForwardDerivatives, camera_projections_functions.h:225-240.
inline void spherical_jacobian(const double* p, double* out, double* J /*2x3*/) {
  const double rt2 = p[0] * p[0] + p[2] * p[2];
  const double rt = std::sqrt(rt2);
  const double R2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
  J[0] = p[2] / (2.0 * M_PI * rt2); J[1] = 0.0; J[2] = -p[0] / (2.0 * M_PI * rt2);
  J[3] = -(p[0] * p[1]) / (2.0 * M_PI * R2 * rt);
  J[4] = rt / (2.0 * M_PI * R2);
  J[5] = -(p[1] * p[2]) / (2.0 * M_PI * R2 * rt);
  spherical_forward(p, out);
}

// Distortion Jacobians.  Each restates the reference formula of that struct's
// ForwardDerivatives<T,true>:
//   Disto2 :23-47, Disto24 :116-147, Disto2468 :220-266, Disto62 :350-411,
//   Disto624 :539-630, DistoBrown :723-773 (camera_distortions_functions.h).
inline void disto_jacobian(StageId s, const double* pt, const double* k, double* out, double* J) {
  const int np = shape_of(s).np;
  const int stride = 2 + np;
  const double x = pt[0], y = pt[1];
  const double x2 = x * x, y2 = y * y, r2 = x2 + y2;
  switch (s) {
    case S_DISTO_2: {
      const double k1 = k[0];
      J[0] = 3.0 * k1 * x2 + k1 * y2 + 1.0;
      J[1] = x * 2.0 * k1 * y;
      J[stride] = y * 2.0 * k1 * x;
      J[stride + 1] = k1 * (3.0 * y2 + x2) + 1.0;
      J[2] = x * r2;
      J[stride + 2] = y * r2;
      break;
    }
    case S_DISTO_24: {
      const double k1 = k[0], k2 = k[1];
      const double x4 = x2 * x2, y4 = y2 * y2;
      J[0] = 5.0 * k2 * x4 + 3.0 * k1 * x2 + 6.0 * k2 * x2 * y2 + k2 * y4 + k1 * y2 + 1.0;
      J[1] = x * (2.0 * k1 * y + 4.0 * k2 * y * r2);
      J[stride] = y * (2.0 * k1 * x + 4.0 * k2 * x * r2);
      J[stride + 1] = 5.0 * k2 * y4 + 3.0 * k1 * y2 + 6.0 * k2 * y2 * x2 + k2 * x4 + k1 * x2 + 1.0;
      J[2] = x * r2; J[3] = x * r2 * r2;
      J[stride + 2] = y * r2; J[stride + 3] = y * r2 * r2;
      break;
    }
    case S_DISTO_2468: {
      const double k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3];
      const double poly = k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2 + k4 * r2 * r2 * r2 * r2 + 1.0;
      const double gx = 2.0 * k1 * x + 4.0 * k2 * x * r2 + 6.0 * k3 * x * r2 * r2 + 8.0 * k4 * x * r2 * r2 * r2;
      const double gy = 2.0 * k1 * y + 4.0 * k2 * y * r2 + 6.0 * k3 * y * r2 * r2 + 8.0 * k4 * y * r2 * r2 * r2;
      J[0] = x * gx + poly;
      J[1] = x * gy;
      J[stride] = y * gx;
      J[stride + 1] = y * gy + poly;
      J[2] = x * r2; J[3] = x * r2 * r2; J[4] = x * r2 * r2 * r2; J[5] = x * r2 * r2 * r2 * r2;
      J[stride + 2] = y * r2; J[stride + 3] = y * r2 * r2; J[stride + 4] = y * r2 * r2 * r2;
      J[stride + 5] = y * r2 * r2 * r2 * r2;
      break;
    }
    case S_DISTO_62:
    case S_DISTO_624: {
      const double k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3], k5 = k[4], k6 = k[5];
      const double p1 = k[6], p2 = k[7];
      const double r2_2 = r2 * r2, r2_3 = r2_2 * r2, r2_4 = r2_3 * r2, r2_5 = r2_4 * r2, r2_6 = r2_5 * r2;
      const double dx_dxt = 2.0 * y * p1 + 6.0 * p2 * x;
      const double dx_dyt = 2.0 * x * p1 + 2.0 * p2 * y;
      const double dy_dxt = dx_dyt;
      const double dy_dyt = 2.0 * x * p2 + 6.0 * p1 * y;
      const double p = radial_poly(r2, k, 6);
      const double dr_dx = 2.0 * x, dr_dy = 2.0 * y;
      const double dp_dr = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r2_2 + 4.0 * k4 * r2_3 + 5.0 * k5 * r2_4 + 6.0 * k6 * r2_5;
      J[0] = p + x * dp_dr * dr_dx + dx_dxt;
      J[1] = x * dp_dr * dr_dy + dx_dyt;
      J[stride] = y * dp_dr * dr_dx + dy_dxt;
      J[stride + 1] = p + y * dp_dr * dr_dy + dy_dyt;
      if (s == S_DISTO_624) {
        const double s0 = k[8], s1 = k[9], s2 = k[10], s3 = k[11];
        J[0] += s0 * 2.0 * x + s1 * 4.0 * x * r2;
        J[1] += s0 * 2.0 * y + s1 * 4.0 * y * r2;
        J[stride] += s2 * 2.0 * x + s3 * 4.0 * x * r2;
        J[stride + 1] += s2 * 2.0 * y + s3 * 4.0 * y * r2;
      }
      const double pw[6] = {r2, r2_2, r2_3, r2_4, r2_5, r2_6};
      for (int i = 0; i < 6; ++i) {
        J[2 + i] = x * pw[i];
        J[stride + 2 + i] = y * pw[i];
      }
      J[8] = 2.0 * x * y;
      J[9] = 3.0 * x2 + y2;
      J[stride + 8] = 3.0 * y2 + x2;
      J[stride + 9] = J[8];
      if (s == S_DISTO_624) {
        // thin-prism parameter derivatives, :616-626
        J[10] = r2; J[11] = r2_2; J[12] = 0.0; J[13] = 0.0;
        J[stride + 10] = 0.0; J[stride + 11] = 0.0; J[stride + 12] = r2; J[stride + 13] = r2_2;
      }
      break;
    }
    case S_DISTO_BROWN: {
      const double k1 = k[0], k2 = k[1], k3 = k[2], p1 = k[3], p2 = k[4];
      const double x4 = x2 * x2, y4 = y2 * y2, r4 = r2 * r2, r6 = r4 * r2;
      J[0] = 5.0 * k2 * x4 + 3.0 * k1 * x2 + 6.0 * k3 * x2 * r4 + 6.0 * k2 * x2 * y2 + k3 * r6 + k2 * y4 +
             k1 * y2 + 1.0 + 2.0 * p1 * y + 6.0 * p2 * x;
      J[1] = x * (2.0 * k1 * y + 4.0 * k2 * y * r2 + 6.0 * k3 * y * r4) + 2.0 * p1 * x + 2.0 * p2 * y;
      J[stride + 1] = 5.0 * k2 * y4 + 3.0 * k1 * y2 + 6.0 * k3 * y2 * r4 + 6.0 * k2 * x2 * y2 + k3 * r6 +
                      k2 * x4 + k1 * x2 + 1.0 + 2.0 * p2 * x + 6.0 * p1 * y;
      J[stride] = y * (2.0 * k1 * x + 4.0 * k2 * x * r2 + 6.0 * k3 * x * r4) + 2.0 * p2 * y + 2.0 * p1 * x;
      J[2] = x * r2; J[3] = x * r2 * r2; J[4] = x * r2 * r2 * r2; J[5] = 2.0 * x * y; J[6] = 3.0 * x2 + y2;
      J[stride + 2] = y * r2; J[stride + 3] = y * r2 * r2; J[stride + 4] = y * r2 * r2 * r2;
      J[stride + 5] = 3.0 * y2 + x2; J[stride + 6] = 2.0 * x * y;
      break;
    }
    default: throw std::runtime_error("not a distortion stage");
  }
  disto_forward(s, pt, k, out);
}

// Affine (transformations_functions.h:22-40), UniformScale (:61-72), Identity (:90-99).
inline void affine_jacobian(StageId s, const double* p, const double* a, double* out, double* J) {
  const int stride = 2 + shape_of(s).np;
  switch (s) {
    case S_AFF_AFFINE:
      J[0] = a[0]; J[1] = 0.0; J[2] = p[0]; J[3] = 0.0; J[4] = 1.0; J[5] = 0.0;
      J[stride] = 0.0; J[stride + 1] = a[0] * a[1];
      J[stride + 2] = p[1] * a[1]; J[stride + 3] = p[1] * a[0]; J[stride + 4] = 0.0; J[stride + 5] = 1.0;
      break;
    case S_AFF_UNIFORM:
      J[0] = a[0]; J[1] = 0.0; J[2] = p[0];
      J[stride] = 0.0; J[stride + 1] = a[0]; J[stride + 2] = p[1];
      break;
    case S_AFF_IDENTITY:
      J[0] = 1.0; J[1] = 0.0; J[stride] = 0.0; J[stride + 1] = 1.0;
      break;
    default: throw std::runtime_error("not an affine stage");
  }
  affine_forward(s, p, a, out);
}

inline void stage_jacobian(StageId s, const double* in, const double* prm, double* out, double* J) {
  switch (s) {
    case S_POSE: pose_jacobian(in, prm, out, J); return;
    case S_NORMALIZE: normalize_jacobian(in, out, J); return;
    case S_PROJ_PERSPECTIVE: perspective_jacobian(in, out, J, 3); return;
    case S_PROJ_FISHEYE: fisheye_jacobian(in, out, J, 3); return;
    case S_PROJ_DUAL: dual_jacobian(in, prm, out, J); return;
    case S_PROJ_SPHERICAL: spherical_jacobian(in, out, J); return;
    case S_DISTO_2: case S_DISTO_24: case S_DISTO_2468: case S_DISTO_62:
    case S_DISTO_624: case S_DISTO_BROWN:
      disto_jacobian(s, in, prm, out, J); return;
    default: affine_jacobian(s, in, prm, out, J); return;
  }
}

// ---------------------------------------------------------------------------
// Chain rule.  geometry::ComposeForwardDerivatives (functions.h:77-97) +
// ComposeDerivatives (functions.h:28-43): stages are applied first to last,
// each reading its parameters at the running offset; the composed Jacobian is
// row-major [input | params(stage 0) | params(stage 1) | ...].
// ---------------------------------------------------------------------------
constexpr int MAXW = 3 + 6 + 6 + 16;  // widest composed Jacobian
inline int chain_jacobian(const StageId* st, int nst, const double* in, const double* params, double* out,
                          double* J /* out_last x width */) {
  double cur[3], nxt[3];
  double Jc[3 * MAXW], Jn[3 * MAXW], Js[3 * (3 + 16)];
  int width = 0, rows = 0, off = 0;
  for (int s = 0; s < nst; ++s) {
    const StageShape sh = shape_of(st[s]);
    const int sstride = sh.in + sh.np;
    stage_jacobian(st[s], s == 0 ? in : cur, params + off, nxt, Js);
    if (s == 0) {
      width = sstride;
      rows = sh.out;
      for (int i = 0; i < rows * width; ++i) Jc[i] = Js[i];
    } else {
      const int nwidth = width + sh.np;
      for (int i = 0; i < sh.out; ++i) {
        for (int j = 0; j < width; ++j) {
          double acc = 0.0;
          for (int k = 0; k < sh.in; ++k) acc += Js[i * sstride + k] * Jc[k * width + j];
          Jn[i * nwidth + j] = acc;
        }
        for (int j = 0; j < sh.np; ++j) Jn[i * nwidth + width + j] = Js[i * sstride + sh.in + j];
      }
      width = nwidth;
      rows = sh.out;
      for (int i = 0; i < rows * width; ++i) Jc[i] = Jn[i];
    }
    for (int i = 0; i < sh.out; ++i) cur[i] = nxt[i];
    off += sh.np;
  }
  for (int i = 0; i < rows; ++i) out[i] = cur[i];
  for (int i = 0; i < rows * width; ++i) J[i] = Jc[i];
  return width;
}

// ---------------------------------------------------------------------------
// The reprojection residual.
// ---------------------------------------------------------------------------

// bundle::ReprojectionError2D::operator() (projection_errors.h:36-56) with
// WorldToCameraCoordinatesRig (error_utils.h:53-85) expressed through the
// PoseFunctor chain of ReprojectionError2DAnalytic (projection_errors.h:80-84):
// residual = (1/sigma) (project(pose_rc(pose_ri(X))) - observed);
// spherical cameras: ReprojectionError3D (projection_errors.h:214-243),
// residual = (1/sigma) (normalize(x_cam) - bearing(observed)), 3 rows.
template <class T>
int reprojection_residual(int type, const T* camera, const T* rig_instance, const T* rig_camera,
                          bool use_rig_camera, const T* point, const double* observed, double sigma, T* r) {
  const double scale = 1.0 / sigma;  // projection_errors.h:21
  T xc[3];
  pose_forward(point, rig_instance, xc);
  if (use_rig_camera) {
    T tmp[3];
    pose_forward(xc, rig_camera, tmp);
    for (int i = 0; i < 3; ++i) xc[i] = tmp[i];
  }
  if (type == SPHERICAL) {
    // bearing of the observation, projection_errors.h:218-222
    const double lon = observed[0] * 2 * M_PI;
    const double lat = -observed[1] * 2 * M_PI;
    const double b[3] = {std::cos(lat) * std::sin(lon), -std::sin(lat), std::cos(lat) * std::cos(lon)};
    T n[3];
    normalize_forward(xc, n);
    for (int i = 0; i < 3; ++i) r[i] = T(scale) * (n[i] - T(b[i]));
    return 3;
  }
  T pr[2];
  camera_project(type, xc, camera, pr);
  r[0] = T(scale) * (pr[0] - T(observed[0]));
  r[1] = T(scale) * (pr[1] - T(observed[1]));
  return 2;
}

// ReprojectionError2DAnalytic<C>::Evaluate (projection_errors.h:67-207) and
// ReprojectionError3DAnalytic::Evaluate (:256-374).  Jacobian blocks are
// row-major: jac_camera nres x C, jac_instance nres x 6, jac_rig_camera
// nres x 6 (zero when !use_rig_camera, :185-189), jac_point nres x 3, all
// multiplied by 1/sigma (:117-147).  Returns the number of residual rows.
inline int reprojection_analytic(int type, const double* camera, const double* rig_instance,
                                 const double* rig_camera, bool use_rig_camera, const double* point,
                                 const double* observed, double sigma, double* r, double* jac_camera,
                                 double* jac_instance, double* jac_rig_camera, double* jac_point) {
  const double scale = 1.0 / sigma;
  const int C = camera_num_params(type);
  StageId st[5];
  int nst = 0;
  double all_params[6 + 6 + 16];
  int np = 0;
  st[nst++] = S_POSE;
  for (int i = 0; i < 6; ++i) all_params[np++] = rig_instance[i];
  if (use_rig_camera) {
    st[nst++] = S_POSE;
    for (int i = 0; i < 6; ++i) all_params[np++] = rig_camera[i];
  }
  const int pose_w = use_rig_camera ? 12 : 6;
  double J[3 * MAXW];
  double pred[3];
  int nres, width;
  if (type == SPHERICAL) {
    st[nst++] = S_NORMALIZE;
    width = chain_jacobian(st, nst, point, all_params, pred, J);
    nres = 3;
    const double lon = observed[0] * 2 * M_PI;
    const double lat = -observed[1] * 2 * M_PI;
    const double b[3] = {std::cos(lat) * std::sin(lon), -std::sin(lat), std::cos(lat) * std::cos(lon)};
    for (int i = 0; i < 3; ++i) r[i] = scale * (pred[i] - b[i]);
    if (jac_camera)
      for (int i = 0; i < 3; ++i) jac_camera[i] = 0.0;  // :303-307
  } else {
    StageId cs[3];
    camera_stages(type, cs);
    for (int i = 0; i < 3; ++i) st[nst++] = cs[i];
    for (int i = 0; i < C; ++i) all_params[np++] = camera[i];
    width = chain_jacobian(st, nst, point, all_params, pred, J);
    nres = 2;
    for (int i = 0; i < 2; ++i) r[i] = scale * (pred[i] - observed[i]);
    if (jac_camera)
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < C; ++j) jac_camera[i * C + j] = scale * J[i * width + 3 + pose_w + j];
  }
  for (int i = 0; i < nres; ++i) {
    if (jac_point)
      for (int j = 0; j < 3; ++j) jac_point[i * 3 + j] = scale * J[i * width + j];
    if (jac_instance)
      for (int j = 0; j < 6; ++j) jac_instance[i * 6 + j] = scale * J[i * width + 3 + j];
    if (jac_rig_camera)
      for (int j = 0; j < 6; ++j) jac_rig_camera[i * 6 + j] = use_rig_camera ? scale * J[i * width + 9 + j] : 0.0;
  }
  return nres;
}

// Same blocks through dual numbers (the reference tests' expected values,
// reprojection_errors_test.cc:30-50,84-96).  Direction order:
// [point(3) | rig_instance(6) | rig_camera(6) | camera(C)].
inline int reprojection_autodiff(int type, const double* camera, const double* rig_instance,
                                 const double* rig_camera, bool use_rig_camera, const double* point,
                                 const double* observed, double sigma, double* r, double* jac_camera,
                                 double* jac_instance, double* jac_rig_camera, double* jac_point) {
  const int C = camera_num_params(type);
  Dual P[3], RI[6], RC[6], CAM[16];
  for (int i = 0; i < 3; ++i) P[i] = Dual::variable(point[i], i);
  for (int i = 0; i < 6; ++i) RI[i] = Dual::variable(rig_instance[i], 3 + i);
  for (int i = 0; i < 6; ++i) RC[i] = Dual::variable(rig_camera ? rig_camera[i] : 0.0, 9 + i);
  for (int i = 0; i < C; ++i) CAM[i] = Dual::variable(camera[i], 15 + i);
  Dual res[3];
  const int nres = reprojection_residual<Dual>(type, CAM, RI, RC, use_rig_camera, P, observed, sigma, res);
  for (int i = 0; i < nres; ++i) {
    r[i] = res[i].v;
    for (int j = 0; j < 3; ++j) jac_point[i * 3 + j] = res[i].d[j];
    for (int j = 0; j < 6; ++j) jac_instance[i * 6 + j] = res[i].d[3 + j];
    for (int j = 0; j < 6; ++j) jac_rig_camera[i * 6 + j] = res[i].d[9 + j];
    for (int j = 0; j < C; ++j) jac_camera[i * C + j] = res[i].d[15 + j];
  }
  return nres;
}

}  // namespace oracle
