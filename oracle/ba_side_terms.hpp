// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the reference's secondary residual functors, templated on the scalar like the
// originals so that `Dual` (ba_functors.hpp) differentiates them the way ceres::Jet does:
//   bundle/error/error_utils.h:14-97        MultRotations, RotatePoint, DiffBetweenAngles
//   bundle/error/position_functors.h:14-66  ShotPositionFunctor, ShotRotationFunctor
//   bundle/error/absolute_motion_errors.h   UpVectorError, Pan/Tilt/RollAngleError, TranslationPriorError
//   bundle/error/relative_motion_errors.h   RelativeMotionError, RelativeRotationError, CommonPositionError
//   bundle/error/motion_prior_errors.h      LinearMotionError
//   bundle/error/parameters_errors.h        StdDeviationConstraint, ParameterBarrier
//   bundle/error/prior_error.h + data/bias.h  DataPriorError<Pose, SimilarityPriorTransform>
// The quaternion / angle-axis helpers restate ceres/rotation.h (third-party, Ceres 2.1 / 2.2: conda.yml:10,
// Dockerfile.ubuntu24:12), which the reference calls from error_utils.h:15-61.
// Parity pin: the reference's own known-answer tests (opensfm/test/test_bundle.py:46-106, 181-316, ...),
// ported to tests/test_bundle_reference.py and run against this oracle and against the CUDA engine.
#pragma once
#include "ba_functors.hpp"

namespace oracle {

inline Dual asin(const Dual& a) { return unary(a, std::asin(a.v), 1.0 / std::sqrt(1.0 - a.v * a.v)); }
inline Dual abs(const Dual& a) { return a.v < 0.0 ? -a : a; }
using std::abs;
using std::asin;
inline double value_of(double x) { return x; }
inline double value_of(const Dual& x) { return x.v; }

// ---- ceres/rotation.h --------------------------------------------------------------------------
template <class T>
void AngleAxisToQuaternion(const T* aa, T* q) {
  const T& a0 = aa[0];
  const T& a1 = aa[1];
  const T& a2 = aa[2];
  const T theta_squared = a0 * a0 + a1 * a1 + a2 * a2;
  if (value_of(theta_squared) > 0.0) {
    const T theta = sqrt(theta_squared);
    const T half_theta = theta * T(0.5);
    const T k = sin(half_theta) / theta;
    q[0] = cos(half_theta);
    q[1] = a0 * k; q[2] = a1 * k; q[3] = a2 * k;
  } else {
    const T k(0.5);
    q[0] = T(1.0);
    q[1] = a0 * k; q[2] = a1 * k; q[3] = a2 * k;
  }
}
template <class T>
void QuaternionToAngleAxis(const T* q, T* aa) {
  const T& q1 = q[1];
  const T& q2 = q[2];
  const T& q3 = q[3];
  const T sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3;
  if (value_of(sin_squared_theta) > 0.0) {
    const T sin_theta = sqrt(sin_squared_theta);
    const T& cos_theta = q[0];
    const T two_theta = T(2.0) * ((value_of(cos_theta) < 0.0) ? atan2(-sin_theta, -cos_theta) : atan2(sin_theta, cos_theta));
    const T k = two_theta / sin_theta;
    aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
  } else {
    const T k(2.0);
    aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
  }
}
template <class T>
void QuaternionProduct(const T z[4], const T w[4], T zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}
template <class T>
void AngleAxisRotatePoint(const T aa[3], const T pt[3], T result[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (value_of(theta2) > DBL_EPSILON) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    for (int i = 0; i < 3; ++i) result[i] = pt[i] * costheta + w_cross_pt[i] * sintheta + w[i] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    for (int i = 0; i < 3; ++i) result[i] = pt[i] + w_cross_pt[i];
  }
}

// ---- error_utils.h / position_functors.h ---------------------------------------------------------
template <class T>
struct V3 {
  T x[3];
  T& operator[](int i) { return x[i]; }
  const T& operator[](int i) const { return x[i]; }
};
template <class T>
V3<T> neg(const V3<T>& a) { return V3<T>{{-a[0], -a[1], -a[2]}}; }
template <class T>
V3<T> MultRotations(const V3<T>& R1, const V3<T>& R2) {
  T q1[4], q2[4], q[4];
  AngleAxisToQuaternion(R1.x, q1);
  AngleAxisToQuaternion(R2.x, q2);
  QuaternionProduct(q1, q2, q);
  V3<T> out;
  QuaternionToAngleAxis(q, out.x);
  return out;
}
template <class T>
V3<T> MultRotations(const V3<T>& R1, const V3<T>& R2, const V3<T>& R3) {
  T q1[4], q2[4], q3[4], q12[4], q[4];
  AngleAxisToQuaternion(R1.x, q1);
  AngleAxisToQuaternion(R2.x, q2);
  AngleAxisToQuaternion(R3.x, q3);
  QuaternionProduct(q1, q2, q12);
  QuaternionProduct(q12, q3, q);
  V3<T> out;
  QuaternionToAngleAxis(q, out.x);
  return out;
}
template <class T>
V3<T> RotatePoint(const V3<T>& R, const V3<T>& x) {
  V3<T> out;
  AngleAxisRotatePoint(R.x, x.x, out.x);
  return out;
}
template <class T>
T DiffBetweenAngles(const T& a, const T& b) {
  const T d = a - b;
  if (value_of(d) > M_PI) return d - T(2 * M_PI);
  if (value_of(d) < -M_PI) return d + T(2 * M_PI);
  return d;
}
// p[i] = parameter block i (pose blocks: [RX RY RZ TX TY TZ]); index -1 = FUNCTOR_NOT_SET
template <class T>
V3<T> ShotPosition(T const* const* p, int inst, int rc) {
  const T* ri = p[inst];
  V3<T> pos{{ri[3], ri[4], ri[5]}};
  if (rc >= 0 && p[rc] != nullptr) {
    const V3<T> R{{ri[0], ri[1], ri[2]}}, t{{p[rc][3], p[rc][4], p[rc][5]}};
    const V3<T> c = RotatePoint(R, t);
    for (int i = 0; i < 3; ++i) pos[i] = pos[i] + c[i];
  }
  return pos;
}
template <class T>
V3<T> ShotRotation(T const* const* p, int inst, int rc) {
  const V3<T> Ri{{p[inst][0], p[inst][1], p[inst][2]}};
  if (rc >= 0 && p[rc] != nullptr) {
    const V3<T> Rc{{p[rc][0], p[rc][1], p[rc][2]}};
    return MultRotations(Ri, Rc);
  }
  return Ri;
}

// ---- the functors: same block order / constants as include/opensfm_b200.h documents ------------
enum SideType {
  SIDE_UP_VECTOR = 0, SIDE_PAN, SIDE_TILT, SIDE_ROLL, SIDE_RELATIVE_MOTION, SIDE_RELATIVE_ROTATION,
  SIDE_COMMON_POSITION, SIDE_LINEAR_MOTION, SIDE_TRANSLATION_PRIOR, SIDE_PARAMETER_BARRIER, SIDE_STD_DEVIATION,
  SIDE_POSITION_PRIOR, SIDE_RA_RELATIVE_MOTION, SIDE_RA_ABSOLUTE_POSITION, SIDE_RA_RELATIVE_ABSOLUTE_POSITION,
  SIDE_RA_COMMON_POINT, SIDE_RA_COMMON_CAMERA
};

// reconstruction_alignment.h:224-234
template <class T>
void transform_point(const T* const reconstruction, const double* point, T* transformed) {
  const T* const R = reconstruction + 0;
  const T* const t = reconstruction + 3;
  const T& scale = reconstruction[6];
  T p_t_s[3] = {(T(point[0]) - t[0]) / scale, (T(point[1]) - t[1]) / scale, (T(point[2]) - t[2]) / scale};
  T Rt[3] = {-R[0], -R[1], -R[2]};
  AngleAxisRotatePoint(Rt, p_t_s, transformed);
}

// Returns false when the functor cannot be evaluated (RelativeMotionError with a zero scale).
template <class T>
bool side_residual(int type, const double* c, const int* aux, T const* const* p, T* r) {
  switch (type) {
    case SIDE_UP_VECTOR: {  // absolute_motion_errors.h:12-39
      const V3<T> R = ShotRotation(p, 0, 1);
      const V3<T> acceleration{{T(c[0]), T(c[1]), T(c[2])}};
      const V3<T> z_world = RotatePoint(R, acceleration);
      r[0] = T(c[3]) * (z_world[0] - T(0.0));
      r[1] = T(c[3]) * (z_world[1] - T(0.0));
      r[2] = T(c[3]) * (z_world[2] - T(1.0));
      return true;
    }
    case SIDE_PAN: {  // :41-65
      const V3<T> R = ShotRotation(p, 0, 1);
      const V3<T> z_axis{{T(0.0), T(0.0), T(1.0)}};
      const V3<T> z_world = RotatePoint(R, z_axis);
      if (value_of(abs(z_world[0])) < 1e-8 && value_of(abs(z_world[1])) < 1e-8) {
        r[0] = T(0.0);
      } else {
        const T predicted_angle = atan2(z_world[0], z_world[1]);
        r[0] = T(c[1]) * DiffBetweenAngles(predicted_angle, T(c[0]));
      }
      return true;
    }
    case SIDE_TILT: {  // :67-90
      const V3<T> R = ShotRotation(p, 0, 1);
      const V3<T> ez{{T(0.0), T(0.0), T(1.0)}};
      const V3<T> Rt_ez = RotatePoint(R, ez);
      const T l = sqrt(Rt_ez[0] * Rt_ez[0] + Rt_ez[1] * Rt_ez[1]);
      const T predicted_angle = -atan2(Rt_ez[2], l);
      r[0] = T(c[1]) * DiffBetweenAngles(predicted_angle, T(c[0]));
      return true;
    }
    case SIDE_ROLL: {  // :92-136
      const V3<T> R = ShotRotation(p, 0, 1);
      const V3<T> ex{{T(1.0), T(0.0), T(0.0)}}, ez{{T(0.0), T(0.0), T(1.0)}};
      const V3<T> Rt_ex = RotatePoint(R, ex), Rt_ez = RotatePoint(R, ez);
      T a[3] = {Rt_ez[1], -Rt_ez[0], T(0.0)};
      const T la = sqrt(a[0] * a[0] + a[1] * a[1]);
      const double eps = 1e-5;
      if (value_of(la) < eps) { r[0] = T(0.0); return true; }
      a[0] = a[0] / la;
      a[1] = a[1] / la;
      const T b[3] = {Rt_ex[1] * a[2] - Rt_ex[2] * a[1], Rt_ex[2] * a[0] - Rt_ex[0] * a[2], Rt_ex[0] * a[1] - Rt_ex[1] * a[0]};
      const T sin_roll = Rt_ez[0] * b[0] + Rt_ez[1] * b[1] + Rt_ez[2] * b[2];
      if (value_of(sin_roll) <= -(1.0 - eps)) { r[0] = T(0.0); return true; }
      const T predicted_angle = asin(sin_roll);
      r[0] = T(c[1]) * DiffBetweenAngles(predicted_angle, T(c[0]));
      return true;
    }
    case SIDE_RELATIVE_MOTION: {  // relative_motion_errors.h:14-72
      const V3<T> Ri = ShotRotation(p, 0, -1), ti = ShotPosition(p, 0, -1);
      const V3<T> Rj = ShotRotation(p, 1, -1), tj = ShotPosition(p, 1, -1);
      T residual[7];
      const V3<T> Rij{{T(c[0]), T(c[1]), T(c[2])}};
      const V3<T> rot = MultRotations(Rij, neg(Ri), Rj);
      for (int k = 0; k < 3; ++k) residual[k] = rot[k];
      const T* scale_i = p[2];
      const T* scale_j = p[aux[0]];
      const V3<T> dt{{ti[0] - tj[0], ti[1] - tj[1], ti[2] - tj[2]}};
      const V3<T> rt = RotatePoint(neg(Rj), dt);
      for (int k = 0; k < 3; ++k) residual[3 + k] = T(c[3 + k]) - scale_j[0] * rt[k];
      if (value_of(scale_i[0]) == 0.0 || value_of(scale_j[0]) == 0.0) return false;
      if (c[56] != 0.0) residual[6] = T(c[6]) - scale_j[0] / scale_i[0];
      else residual[6] = T(0.0);
      for (int a = 0; a < 7; ++a) {
        T s(0.0);
        for (int b = 0; b < 7; ++b) s = s + T(c[7 + 7 * a + b]) * residual[b];
        r[a] = s;
      }
      return true;
    }
    case SIDE_RELATIVE_ROTATION: {  // :74-103
      const V3<T> Ri = ShotRotation(p, 0, aux[0]), Rj = ShotRotation(p, 1, aux[1]);
      const V3<T> Rij{{T(c[0]), T(c[1]), T(c[2])}};
      const V3<T> e = MultRotations(Rij, neg(Ri), Rj);
      for (int a = 0; a < 3; ++a) r[a] = T(c[3 + 3 * a]) * e[0] + T(c[4 + 3 * a]) * e[1] + T(c[5 + 3 * a]) * e[2];
      return true;
    }
    case SIDE_COMMON_POSITION: {  // :105-138
      const V3<T> t1 = ShotPosition(p, 0, aux[0]), t2 = ShotPosition(p, 1, aux[1]);
      T error[3] = {t1[0] - t2[0], t1[1] - t2[1], t1[2] - t2[2]};
      for (int i = 0; i < 2; ++i) {
        const T m = abs(error[i]) - T(c[0]);
        error[i] = value_of(m) > 0.0 ? m : T(0.0);  // std::max(T(0.0), m)
      }
      for (int i = 0; i < 3; ++i) r[i] = T(c[1]) * error[i];
      return true;
    }
    case SIDE_LINEAR_MOTION: {  // motion_prior_errors.h:13-76
      const V3<T> R0 = ShotRotation(p, 0, aux[0]), t0 = ShotPosition(p, 0, aux[0]);
      const V3<T> R1 = ShotRotation(p, 1, aux[1]), t1 = ShotPosition(p, 1, aux[1]);
      const V3<T> R2 = ShotRotation(p, 2, aux[2]), t2 = ShotPosition(p, 2, aux[2]);
      const double eps = 1e-15;
      const T t2_t0[3] = {t2[0] - t0[0], t2[1] - t0[1], t2[2] - t0[2]};
      const T t1_t0[3] = {t1[0] - t0[0], t1[1] - t0[1], t1[2] - t0[2]};
      const T t2_t0_norm = sqrt(t2_t0[0] * t2_t0[0] + t2_t0[1] * t2_t0[1] + t2_t0[2] * t2_t0[2]);
      const T t1_t0_norm = sqrt(t1_t0[0] * t1_t0[0] + t1_t0[1] * t1_t0[1] + t1_t0[2] * t1_t0[2]);
      for (int i = 0; i < 3; ++i) {
        if (value_of(t2_t0_norm) > eps) r[i] = T(c[1]) * (T(c[0]) - t1_t0_norm / t2_t0_norm);
        else r[i] = T(c[1]) * (T(c[0]) * t2_t0[i] - t1_t0[i]);
      }
      V3<T> R2_R0t = MultRotations(R2, neg(R0));
      for (int i = 0; i < 3; ++i) R2_R0t[i] = T(c[0]) * R2_R0t[i];
      const V3<T> R0_R1t = MultRotations(R0, neg(R1));
      const V3<T> e = MultRotations(R2_R0t, R0_R1t);
      for (int i = 0; i < 3; ++i) r[3 + i] = T(c[2]) * e[i];
      return true;
    }
    case SIDE_TRANSLATION_PRIOR: {  // absolute_motion_errors.h:180-202
      const T* a = p[0];
      const T* b = p[1];
      const T d[3] = {a[3] - b[3], a[4] - b[4], a[5] - b[5]};
      const T safe_norm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + T(1e-20));
      r[0] = log(safe_norm / T(c[0]));
      return true;
    }
    case SIDE_PARAMETER_BARRIER: {  // parameters_errors.h:20-36
      const T eps = T(1e-10);
      const T value = p[0][aux[0]];
      const T zero = T(2.0 * std::log((c[1] - c[0]) * 0.5));
      const T penalty = log(value - T(c[0]) + eps) + log(T(c[1]) - value + eps);
      r[0] = penalty + zero;
      return true;
    }
    case SIDE_STD_DEVIATION: {  // parameters_errors.h:7-18
      const T std_sq = p[0][0] * p[0][0] + T(1e-20);
      r[0] = log(T(1.0) / sqrt(T(2.0 * M_PI) * std_sq));
      return true;
    }
    case SIDE_POSITION_PRIOR: {  // prior_error.h:55-96, data/bias.h:33-53
      const T* inst = p[0];
      const T* bias = p[1];
      const V3<T> R{{bias[0], bias[1], bias[2]}};
      const V3<T> prior{{T(c[0]), T(c[1]), T(c[2])}};
      const V3<T> rp = RotatePoint(R, prior);
      for (int k = 0; k < 3; ++k) {
        const T point = bias[6] * rp[k] + bias[3 + k];
        T scale = T(c[3 + k]);
        if (c[6] != 0.0) scale = scale / p[2][0];
        r[k] = scale * (inst[3 + k] - point);
      }
      return true;
    }
    // ---- ReconstructionAlignment functors (reconstruction_alignment.h) ----
    case SIDE_RA_RELATIVE_MOTION: {  // :140-197
      const T* const reconstruction_a = p[0];
      const T* const shot_i = p[1];
      T rr[6];
      const T* const Ri = shot_i + 0;
      const T* const Ra = reconstruction_a + 0;
      T Rit[3] = {-Ri[0], -Ri[1], -Ri[2]};
      const T* const ti = shot_i + 3;
      const T* const ta = reconstruction_a + 3;
      const T* const scale_a = reconstruction_a + 6;
      T Rai[3] = {T(c[0]), T(c[1]), T(c[2])};
      T tai[3] = {T(c[3]), T(c[4]), T(c[5])};
      T Rait[3] = {-Rai[0], -Rai[1], -Rai[2]};
      T qRai[4], qRa[4], qRit[4], qRa_Rit[4], qRai_Ra_Rit[4];
      AngleAxisToQuaternion(Rai, qRai);
      AngleAxisToQuaternion(Ra, qRa);
      AngleAxisToQuaternion(Rit, qRit);
      QuaternionProduct(qRa, qRit, qRa_Rit);
      QuaternionProduct(qRai, qRa_Rit, qRai_Ra_Rit);
      T Rai_Ra_Rit[3];
      QuaternionToAngleAxis(qRai_Ra_Rit, Rai_Ra_Rit);
      rr[0] = Rai_Ra_Rit[0]; rr[1] = Rai_Ra_Rit[1]; rr[2] = Rai_Ra_Rit[2];
      T Rait_tai[3], Rit_ti[3], Ra_Rit_ti[3];
      AngleAxisRotatePoint(Rait, tai, Rait_tai);
      AngleAxisRotatePoint(Rit, ti, Rit_ti);
      AngleAxisRotatePoint(Ra, Rit_ti, Ra_Rit_ti);
      for (int k = 0; k < 3; ++k) rr[3 + k] = Rait_tai[k] - scale_a[0] * Ra_Rit_ti[k] + ta[k];
      for (int i = 0; i < 6; ++i) {
        r[i] = T(0.0);
        for (int j = 0; j < 6; ++j) r[i] = r[i] + T(c[6 + 6 * i + j]) * rr[j];
      }
      return true;
    }
    case SIDE_RA_ABSOLUTE_POSITION: {  // :199-222
      const T* const shot_i = p[0];
      T Rit[3] = {-shot_i[0], -shot_i[1], -shot_i[2]};
      T Rit_ti[3];
      AngleAxisRotatePoint(Rit, shot_i + 3, Rit_ti);
      for (int k = 0; k < 3; ++k) r[k] = T(c[3]) * (T(c[k]) + Rit_ti[k]);
      return true;
    }
    case SIDE_RA_RELATIVE_ABSOLUTE_POSITION: {  // :236-265
      const double* const Ri = c + 3;
      const double* const ti = c + 6;
      double Rit[3] = {-Ri[0], -Ri[1], -Ri[2]};
      double Rit_ti[3];
      AngleAxisRotatePoint(Rit, ti, Rit_ti);
      double minus_Rit_ti[3] = {-Rit_ti[0], -Rit_ti[1], -Rit_ti[2]};
      T transformed[3];
      transform_point(p[0], minus_Rit_ti, transformed);
      for (int k = 0; k < 3; ++k) r[k] = T(c[9]) * (T(c[k]) - transformed[k]);
      return true;
    }
    case SIDE_RA_COMMON_POINT: {  // :267-296
      T transformed_pai[3], transformed_pbi[3];
      transform_point(p[0], c, transformed_pai);
      transform_point(p[1], c + 3, transformed_pbi);
      const T scale_factor = p[0][6] + p[1][6];
      for (int k = 0; k < 3; ++k) r[k] = T(c[6]) * scale_factor * (transformed_pai[k] - transformed_pbi[k]);
      return true;
    }
    case SIDE_RA_COMMON_CAMERA: {  // :298-365
      const double* const rotation_ai = c;
      const double* const rotation_bi = c + 6;
      const double* const translation_ai = c + 3;
      const double* const translation_bi = c + 9;
      double pose_ai[3], pose_bi[3];
      const double rotation_ait[3] = {-rotation_ai[0], -rotation_ai[1], -rotation_ai[2]};
      AngleAxisRotatePoint(rotation_ait, translation_ai, pose_ai);
      const double rotation_bit[3] = {-rotation_bi[0], -rotation_bi[1], -rotation_bi[2]};
      AngleAxisRotatePoint(rotation_bit, translation_bi, pose_bi);
      for (int i = 0; i < 3; ++i) { pose_ai[i] = -pose_ai[i]; pose_bi[i] = -pose_bi[i]; }
      T world_pose_ai[3], world_pose_bi[3];
      transform_point(p[0], pose_ai, world_pose_ai);
      transform_point(p[1], pose_bi, world_pose_bi);
      const T* const Ra = p[0];
      const T* const Rb = p[1];
      const T Rbt[3] = {-Rb[0], -Rb[1], -Rb[2]};
      const T Rbit[3] = {T(-rotation_bi[0]), T(-rotation_bi[1]), T(-rotation_bi[2])};
      const T Rai[3] = {T(rotation_ai[0]), T(rotation_ai[1]), T(rotation_ai[2])};
      T qRai[4], qRa[4], qRbt[4], qRbit[4], qRai_qRa[4], qRai_qRa_qRbt[4], qRai_qRa_qRbt_qRbit[4];
      AngleAxisToQuaternion(Rai, qRai);
      AngleAxisToQuaternion(Ra, qRa);
      AngleAxisToQuaternion(Rbt, qRbt);
      AngleAxisToQuaternion(Rbit, qRbit);
      QuaternionProduct(qRai, qRa, qRai_qRa);
      QuaternionProduct(qRai_qRa, qRbt, qRai_qRa_qRbt);
      QuaternionProduct(qRai_qRa_qRbt, qRbit, qRai_qRa_qRbt_qRbit);
      QuaternionToAngleAxis(qRai_qRa_qRbt_qRbit, r);
      for (int i = 0; i < 3; ++i) {
        r[i] = r[i] * T(c[13]);
        r[3 + i] = T(c[12]) * (world_pose_ai[i] - world_pose_bi[i]);
      }
      return true;
    }
  }
  throw std::runtime_error("unknown side term");
}

// ceres::TukeyLoss (used with a = 1 for CommonPositionError, bundle_adjuster.cc:905)
inline void tukey_loss(double a, double s, double* rho) {
  const double a2 = a * a;
  if (s <= a2) {
    const double value = 1.0 - s / a2, value_sq = value * value;
    rho[0] = a2 / 3.0 * (1.0 - value_sq * value);
    rho[1] = value_sq;
  } else {
    rho[0] = a2 / 3.0;
    rho[1] = 0.0;
  }
}

}  // namespace oracle
