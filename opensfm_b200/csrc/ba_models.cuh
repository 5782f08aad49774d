// Per-observation reprojection residual and analytic Jacobian for every OpenSfM
// camera model — the inner loop of bundle adjustment, hand-derived for the GPU.
//
// What this replaces (behaviour, not code): the Ceres cost functor
//   bundle::ReprojectionError2DAnalytic<C>::Evaluate   opensfm/src/bundle/error/projection_errors.h:59-208
//   bundle::ReprojectionError3DAnalytic::Evaluate      projection_errors.h:248-375
// which the reference evaluates through a generic chain-rule composer
// (geometry/functions.h:70-97) over PoseFunctor / PROJ / DISTO / AFFINE functors
// (geometry/transformations_functions.h, camera_projections_functions.h,
// camera_distortions_functions.h) and a dual-number rotation derivative.
//
// Here the chain is fused and specialised:
//  * the rotation derivative is closed form,
//      d(R(v)u)/dv = -R [u]x (v v^T + (R^T - I)[v]x) / |v|^2,
//    evaluated with R - I = A [v]x + B [v]x^2, A = sin(t)/t, B = 2 sin^2(t/2)/t^2
//    (no 1 - cos cancellation), small-angle branch as the reference
//    (transformations_functions.h:127-143);
//  * all six distortion families are one polynomial form
//    (n radial coefficients, optional tangential pair, optional thin prism);
//  * 2x2 / 2x3 products are written out, nothing is staged through memory.
//
// Parameter order is the reference's [PROJ | DISTO | AFFINE]
// (geometry/src/camera.cc:9-178).  Functions are __host__ __device__ so the
// same source is checked on the CPU against the oracle's dual numbers
// (tests/test_ba_models_cpu.py); the product only ever calls them from kernels.
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define OSFM_HD __host__ __device__ __forceinline__
#else
#define OSFM_HD inline
#endif

namespace osfm {

enum { PT_PERSPECTIVE = 0, PT_BROWN = 1, PT_FISHEYE = 2, PT_FISHEYE_OPENCV = 3, PT_FISHEYE62 = 4,
       PT_FISHEYE624 = 5, PT_SPHERICAL = 6, PT_DUAL = 7, PT_RADIAL = 8, PT_SIMPLE_RADIAL = 9 };

constexpr int MAX_CAM_PARAMS = 16;

struct ModelSpec {
  int proj;      // 0 perspective, 1 fisheye, 2 dual, 3 spherical
  int n_radial;  // radial coefficients
  int tangential;
  int prism;
  int affine;    // 0 uniform scale [f], 1 affine [f, ar, cx, cy], 2 identity
};

// camera_instances.h:181-193
OSFM_HD ModelSpec model_spec(int type) {
  switch (type) {
    case PT_PERSPECTIVE: return {0, 2, 0, 0, 0};
    case PT_BROWN: return {0, 3, 1, 0, 1};
    case PT_FISHEYE: return {1, 2, 0, 0, 0};
    case PT_FISHEYE_OPENCV: return {1, 4, 0, 0, 1};
    case PT_FISHEYE62: return {1, 6, 1, 0, 1};
    case PT_FISHEYE624: return {1, 6, 1, 1, 1};
    case PT_SPHERICAL: return {3, 0, 0, 0, 2};
    case PT_DUAL: return {2, 2, 0, 0, 0};
    case PT_RADIAL: return {0, 2, 0, 0, 1};
    case PT_SIMPLE_RADIAL: return {0, 1, 0, 0, 1};
  }
  return {0, 0, 0, 0, 2};
}

OSFM_HD int model_num_params(int type) {
  const ModelSpec m = model_spec(type);
  if (m.proj == 3) return 1;  // spherical stores one unused value (camera_instances.h:121-128)
  return (m.proj == 2 ? 1 : 0) + m.n_radial + 2 * m.tangential + 4 * m.prism + (m.affine == 0 ? 1 : 4);
}

// x_cam = R(-r) (X - t) and its derivatives.  rt = [r | t], camera->world
// angle-axis and camera origin (bundle/data/pose.h:34-43).
// Outputs: xc[3]; R row-major 3x3 (= d xc / dX; d xc / dt = -R); Jr row-major 3x3 = d xc / d r.
OSFM_HD void pose_apply(const double* X, const double* rt, double* xc, double* R, double* Jr) {
  const double d0 = X[0] - rt[3], d1 = X[1] - rt[4], d2 = X[2] - rt[5];
  const double v0 = -rt[0], v1 = -rt[1], v2 = -rt[2];
  const double th2 = v0 * v0 + v1 * v1 + v2 * v2;
  if (th2 > DBL_EPSILON) {
    const double th = sqrt(th2);
    double sn, cs;
#ifdef __CUDA_ARCH__
    sincos(th, &sn, &cs);
#else
    sn = sin(th); cs = cos(th);
#endif
    const double sh = sin(0.5 * th);
    const double A = sn / th;
    const double B = 2.0 * sh * sh / th2;
    // K = [v]x ; K2 = [v]x^2 = v v^T - th2 I
    const double k2_00 = v0 * v0 - th2, k2_11 = v1 * v1 - th2, k2_22 = v2 * v2 - th2;
    const double k2_01 = v0 * v1, k2_02 = v0 * v2, k2_12 = v1 * v2;
    R[0] = 1.0 + B * k2_00;       R[1] = -A * v2 + B * k2_01;  R[2] = A * v1 + B * k2_02;
    R[3] = A * v2 + B * k2_01;    R[4] = 1.0 + B * k2_11;      R[5] = -A * v0 + B * k2_12;
    R[6] = -A * v1 + B * k2_02;   R[7] = A * v0 + B * k2_12;   R[8] = 1.0 + B * k2_22;
    xc[0] = R[0] * d0 + R[1] * d1 + R[2] * d2;
    xc[1] = R[3] * d0 + R[4] * d1 + R[5] * d2;
    xc[2] = R[6] * d0 + R[7] * d1 + R[8] * d2;
    if (Jr) {
      // M = (v v^T + (R^T - I) K) / th2,  R^T - I = -A K + B K2
      // (R^T - I) K = -A K2 + B K2 K = -A K2 - B th2 K      (K^3 = -th2 K)
      double M[9];
      const double inv = 1.0 / th2;
      M[0] = (v0 * v0 - A * k2_00) * inv;
      M[4] = (v1 * v1 - A * k2_11) * inv;
      M[8] = (v2 * v2 - A * k2_22) * inv;
      M[1] = (v0 * v1 - A * k2_01 + B * th2 * v2) * inv;   // K01 = -v2
      M[3] = (v0 * v1 - A * k2_01 - B * th2 * v2) * inv;   // K10 =  v2
      M[2] = (v0 * v2 - A * k2_02 - B * th2 * v1) * inv;   // K02 =  v1
      M[6] = (v0 * v2 - A * k2_02 + B * th2 * v1) * inv;   // K20 = -v1
      M[5] = (v1 * v2 - A * k2_12 + B * th2 * v0) * inv;   // K12 = -v0
      M[7] = (v1 * v2 - A * k2_12 - B * th2 * v0) * inv;   // K21 =  v0
      // d xc / d v = -R [d]x M ;  d xc / d r = +R [d]x M
      // T = [d]x M
      double T[9];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T[0 + j] = -d2 * M[3 + j] + d1 * M[6 + j];
        T[3 + j] = d2 * M[0 + j] - d0 * M[6 + j];
        T[6 + j] = -d1 * M[0 + j] + d0 * M[3 + j];
      }
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          Jr[i * 3 + j] = R[i * 3] * T[j] + R[i * 3 + 1] * T[3 + j] + R[i * 3 + 2] * T[6 + j];
    }
  } else {
    // R = I + [v]x (first order), as the reference's Taylor branch
    R[0] = 1.0; R[1] = -v2; R[2] = v1;
    R[3] = v2;  R[4] = 1.0; R[5] = -v0;
    R[6] = -v1; R[7] = v0;  R[8] = 1.0;
    xc[0] = d0 + (v1 * d2 - v2 * d1);
    xc[1] = d1 + (v2 * d0 - v0 * d2);
    xc[2] = d2 + (v0 * d1 - v1 * d0);
    if (Jr) {
      // d([v]x d)/dv = -[d]x ; d/dr = +[d]x
      Jr[0] = 0.0; Jr[1] = -d2; Jr[2] = d1;
      Jr[3] = d2;  Jr[4] = 0.0; Jr[5] = -d0;
      Jr[6] = -d1; Jr[7] = d0;  Jr[8] = 0.0;
    }
  }
}

// 3-D -> 2-D projection stage.  Ju = d(u)/d(xc) 2x3 row-major.
OSFM_HD void proj_perspective(const double* p, double* u, double* Ju) {
  const double iz = 1.0 / p[2];
  u[0] = p[0] * iz;
  u[1] = p[1] * iz;
  if (Ju) {
    Ju[0] = iz; Ju[1] = 0.0; Ju[2] = -u[0] * iz;
    Ju[3] = 0.0; Ju[4] = iz; Ju[5] = -u[1] * iz;
  }
}

OSFM_HD void proj_fisheye(const double* p, double* u, double* Ju) {
  const double x = p[0], y = p[1], z = p[2];
  const double r2 = x * x + y * y;
  const double r = sqrt(r2);
  if (r < 1e-8) {  // camera_projections_functions.h:11-15,31-44
    proj_perspective(p, u, Ju);
    return;
  }
  const double theta = atan2(r, z);
  const double s = theta / r;
  u[0] = s * x;
  u[1] = s * y;
  if (Ju) {
    const double R2 = r2 + z * z;
    // ds/dr = (dtheta/dr * r - theta) / r^2, dtheta/dr = z / R2, dtheta/dz = -r / R2
    const double ds_r = (z * r / R2 - theta) / r2;
    const double sx = ds_r * x / r, sy = ds_r * y / r;
    Ju[0] = s + x * sx; Ju[1] = x * sy;     Ju[2] = -x / R2;
    Ju[3] = y * sx;     Ju[4] = s + y * sy; Ju[5] = -y / R2;
  }
}

// Radial + tangential + thin-prism distortion.  k = [k_1..k_n | p1 p2 | s0 s1 s2 s3].
// Jv = d(out)/d(in) 2x2; Jk = d(out)/d(k) 2 x nk (row-major, nk = n + 2 tan + 4 prism).
OSFM_HD void distort(const ModelSpec& m, const double* in, const double* k, double* out, double* Jv, double* Jk,
                     int nk) {
  const double x = in[0], y = in[1];
  const double x2 = x * x, y2 = y * y, r2 = x2 + y2;
  // rad = 1 + sum k_i r2^i ; drad = d rad / d r2
  double rad = 0.0, drad = 0.0;
  for (int i = m.n_radial - 1; i >= 0; --i) {
    drad = drad * r2 + (i + 1) * k[i];
    rad = (rad + k[i]) * r2;
  }
  rad += 1.0;
  double ox = x * rad, oy = y * rad;
  double jxx = rad + 2.0 * x2 * drad, jxy = 2.0 * x * y * drad, jyx = jxy, jyy = rad + 2.0 * y2 * drad;
  int o = m.n_radial;
  if (Jk) {
    double pw = r2;
    for (int i = 0; i < m.n_radial; ++i) {
      Jk[i] = x * pw;
      Jk[nk + i] = y * pw;
      pw *= r2;
    }
  }
  if (m.tangential) {
    const double p1 = k[o], p2 = k[o + 1];
    ox += 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x2);
    oy += 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y2);
    jxx += 2.0 * p1 * y + 6.0 * p2 * x;
    jxy += 2.0 * p1 * x + 2.0 * p2 * y;
    jyx += 2.0 * p2 * y + 2.0 * p1 * x;
    jyy += 2.0 * p2 * x + 6.0 * p1 * y;
    if (Jk) {
      Jk[o] = 2.0 * x * y;          Jk[o + 1] = r2 + 2.0 * x2;
      Jk[nk + o] = r2 + 2.0 * y2;   Jk[nk + o + 1] = 2.0 * x * y;
    }
    o += 2;
  }
  if (m.prism) {
    const double s0 = k[o], s1 = k[o + 1], s2 = k[o + 2], s3 = k[o + 3];
    const double r4 = r2 * r2;
    ox += s0 * r2 + s1 * r4;
    oy += s2 * r2 + s3 * r4;
    const double gx = 2.0 * (s0 + 2.0 * s1 * r2), gy = 2.0 * (s2 + 2.0 * s3 * r2);
    jxx += gx * x; jxy += gx * y;
    jyx += gy * x; jyy += gy * y;
    if (Jk) {
      Jk[o] = r2; Jk[o + 1] = r4; Jk[o + 2] = 0.0; Jk[o + 3] = 0.0;
      Jk[nk + o] = 0.0; Jk[nk + o + 1] = 0.0; Jk[nk + o + 2] = r2; Jk[nk + o + 3] = r4;
    }
  }
  out[0] = ox;
  out[1] = oy;
  if (Jv) {
    Jv[0] = jxx; Jv[1] = jxy; Jv[2] = jyx; Jv[3] = jyy;
  }
}

// Full camera: pixel = AFFINE(DISTO(PROJ(xc))).  params in the reference order.
// Jx: d(pixel)/d(xc) 2x3; Jc: d(pixel)/d(params) 2xC (either may be null together).
OSFM_HD void camera_project(int type, const double* params, const double* xc, double* px, double* Jx, double* Jc) {
  const ModelSpec m = model_spec(type);
  const int C = model_num_params(type);
  const int np_proj = m.proj == 2 ? 1 : 0;
  const int nk = m.n_radial + 2 * m.tangential + 4 * m.prism;
  const bool want = Jx != nullptr;
  double u[2], Ju[6], Jt[2] = {0.0, 0.0};
  if (m.proj == 0) {
    proj_perspective(xc, u, want ? Ju : nullptr);
  } else if (m.proj == 1) {
    proj_fisheye(xc, u, want ? Ju : nullptr);
  } else {  // dual: blend of both (camera_projections_functions.h:124-174)
    double up[2], uf[2], Jp[6], Jf[6];
    proj_perspective(xc, up, want ? Jp : nullptr);
    proj_fisheye(xc, uf, want ? Jf : nullptr);
    const double t = params[0];
    u[0] = t * up[0] + (1.0 - t) * uf[0];
    u[1] = t * up[1] + (1.0 - t) * uf[1];
    if (want) {
#pragma unroll
      for (int i = 0; i < 6; ++i) Ju[i] = t * Jp[i] + (1.0 - t) * Jf[i];
      Jt[0] = up[0] - uf[0];
      Jt[1] = up[1] - uf[1];
    }
  }
  double v[2], Jv[4], Jk[2 * 12];
  distort(m, u, params + np_proj, v, want ? Jv : nullptr, want ? Jk : nullptr, nk);
  const double* a = params + np_proj + nk;
  double fx, fy;
  if (m.affine == 0) {
    fx = a[0]; fy = a[0];
    px[0] = fx * v[0];
    px[1] = fy * v[1];
  } else {
    fx = a[0]; fy = a[0] * a[1];
    px[0] = fx * v[0] + a[2];
    px[1] = fy * v[1] + a[3];
  }
  if (!want) return;
  // G = diag(fx, fy) Jv  (2x2)
  const double g00 = fx * Jv[0], g01 = fx * Jv[1], g10 = fy * Jv[2], g11 = fy * Jv[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    Jx[j] = g00 * Ju[j] + g01 * Ju[3 + j];
    Jx[3 + j] = g10 * Ju[j] + g11 * Ju[3 + j];
  }
  if (np_proj) {
    Jc[0] = g00 * Jt[0] + g01 * Jt[1];
    Jc[C] = g10 * Jt[0] + g11 * Jt[1];
  }
  for (int j = 0; j < nk; ++j) {
    Jc[np_proj + j] = fx * Jk[j];
    Jc[C + np_proj + j] = fy * Jk[nk + j];
  }
  const int oa = np_proj + nk;
  if (m.affine == 0) {
    Jc[oa] = v[0];
    Jc[C + oa] = v[1];
  } else {
    Jc[oa] = v[0];           Jc[oa + 1] = 0.0;          Jc[oa + 2] = 1.0; Jc[oa + 3] = 0.0;
    Jc[C + oa] = a[1] * v[1]; Jc[C + oa + 1] = a[0] * v[1]; Jc[C + oa + 2] = 0.0; Jc[C + oa + 3] = 1.0;
  }
}

// One observation.  Residual rows nres = 2 (pixel error) or 3 (spherical bearing error,
// projection_errors.h:214-243).  Outputs are scaled by 1/sigma (projection_errors.h:203-205):
//   r[nres]; Jcam[nres x C]; Jri[nres x 6] (rig instance); Jrc[nres x 6] (rig camera, only if
//   use_rc, else untouched); Jpt[nres x 3].  Any J pointer set may be null (residual only).
OSFM_HD int observation_eval(int type, const double* cam, const double* ri, const double* rc, bool use_rc,
                             const double* X, double ox, double oy, double inv_sigma, double* r, double* Jcam,
                             double* Jri, double* Jrc, double* Jpt) {
  const bool want = Jpt != nullptr;
  double x1[3], R1[9], Jr1[9];
  pose_apply(X, ri, x1, R1, want ? Jr1 : nullptr);
  double xc[3], R2[9], Jr2[9];
  if (use_rc) {
    pose_apply(x1, rc, xc, R2, want ? Jr2 : nullptr);
  } else {
    xc[0] = x1[0]; xc[1] = x1[1]; xc[2] = x1[2];
  }
  // D: nres x 3 = d(residual)/d(xc) (unscaled)
  double D[9];
  int nres;
  const int C = model_num_params(type);
  if (type == PT_SPHERICAL) {
    nres = 3;
    const double n2 = xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2];
    const double inv_n = 1.0 / sqrt(n2);
    const double b0 = xc[0] * inv_n, b1 = xc[1] * inv_n, b2 = xc[2] * inv_n;
    const double lon = ox * 2.0 * M_PI, lat = -oy * 2.0 * M_PI;
    const double cl = cos(lat);
    r[0] = inv_sigma * (b0 - cl * sin(lon));
    r[1] = inv_sigma * (b1 + sin(lat));
    r[2] = inv_sigma * (b2 - cl * cos(lon));
    if (want) {
      D[0] = (1.0 - b0 * b0) * inv_n; D[1] = -b0 * b1 * inv_n;        D[2] = -b0 * b2 * inv_n;
      D[3] = D[1];                     D[4] = (1.0 - b1 * b1) * inv_n; D[5] = -b1 * b2 * inv_n;
      D[6] = D[2];                     D[7] = D[5];                    D[8] = (1.0 - b2 * b2) * inv_n;
      if (Jcam) { Jcam[0] = 0.0; Jcam[1] = 0.0; Jcam[2] = 0.0; }
    }
  } else {
    nres = 2;
    double px[2], Jc[2 * MAX_CAM_PARAMS];
    camera_project(type, cam, xc, px, want ? D : nullptr, want ? Jc : nullptr);
    r[0] = inv_sigma * (px[0] - ox);
    r[1] = inv_sigma * (px[1] - oy);
    if (want && Jcam)
      for (int i = 0; i < 2 * C; ++i) Jcam[i] = inv_sigma * Jc[i];
  }
  if (!want) return nres;
  // E = D * R2 (or D): d(residual)/d(x1)
  double E[9];
  if (use_rc) {
    for (int i = 0; i < nres; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
        E[i * 3 + j] = D[i * 3] * R2[j] + D[i * 3 + 1] * R2[3 + j] + D[i * 3 + 2] * R2[6 + j];
      if (Jrc) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          Jrc[i * 6 + j] = inv_sigma * (D[i * 3] * Jr2[j] + D[i * 3 + 1] * Jr2[3 + j] + D[i * 3 + 2] * Jr2[6 + j]);
          Jrc[i * 6 + 3 + j] = -inv_sigma * E[i * 3 + j];
        }
      }
    }
  } else {
    for (int i = 0; i < nres * 3; ++i) E[i] = D[i];
  }
  for (int i = 0; i < nres; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double dX = E[i * 3] * R1[j] + E[i * 3 + 1] * R1[3 + j] + E[i * 3 + 2] * R1[6 + j];
      Jpt[i * 3 + j] = inv_sigma * dX;
      Jri[i * 6 + 3 + j] = -inv_sigma * dX;
      Jri[i * 6 + j] = inv_sigma * (E[i * 3] * Jr1[j] + E[i * 3 + 1] * Jr1[3 + j] + E[i * 3 + 2] * Jr1[6 + j]);
    }
  }
  return nres;
}

// ceres::LossFunction values used by CreateLossFunction (bundle_adjuster.cc:414-429):
// returns rho(s) and sets *w = sqrt(rho'(s)) (the Corrector scale; rho'' <= 0 for all five).
OSFM_HD double robust_loss(int loss, double a, double s, double* w) {
  const double kMin = DBL_MIN;
  double rho0, rho1;
  switch (loss) {
    default:
    case 0: rho0 = s; rho1 = 1.0; break;
    case 1: {  // Huber
      const double b = a * a;
      if (s > b) {
        const double rr = sqrt(s);
        rho0 = 2.0 * a * rr - b;
        rho1 = fmax(kMin, a / rr);
      } else {
        rho0 = s; rho1 = 1.0;
      }
      break;
    }
    case 2: {  // SoftLOne
      const double b = a * a;
      const double tmp = sqrt(1.0 + s / b);
      rho0 = 2.0 * b * (tmp - 1.0);
      rho1 = fmax(kMin, 1.0 / tmp);
      break;
    }
    case 3: {  // Cauchy
      const double b = a * a;
      const double sum = 1.0 + s / b;
      rho0 = b * log(sum);
      rho1 = fmax(kMin, 1.0 / sum);
      break;
    }
    case 4: {  // Arctan
      const double sum = 1.0 + s * s / (a * a);
      rho0 = a * atan2(s, a);
      rho1 = fmax(kMin, 1.0 / sum);
      break;
    }
  }
  *w = sqrt(rho1);
  return rho0;
}

}  // namespace osfm
