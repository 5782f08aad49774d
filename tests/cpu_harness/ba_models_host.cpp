// Test-only: compiles the product's __host__ __device__ per-observation math with g++
// so tests can compare it against the oracle's dual numbers without a GPU.
#include "../../opensfm_b200/csrc/ba_models.cuh"
extern "C" {
int hd_num_params(int type) { return osfm::model_num_params(type); }
int hd_observation_eval(int type, const double* cam, const double* ri, const double* rc, int use_rc,
                        const double* X, const double* obs, double sigma, double* r, double* jc, double* jri,
                        double* jrc, double* jp, int want_jac) {
  return osfm::observation_eval(type, cam, ri, rc, use_rc != 0, X, obs[0], obs[1], 1.0 / sigma, r,
                                want_jac ? jc : nullptr, want_jac ? jri : nullptr, want_jac ? jrc : nullptr,
                                want_jac ? jp : nullptr);
}
double hd_loss(int loss, double a, double s, double* w) { return osfm::robust_loss(loss, a, s, w); }
}
