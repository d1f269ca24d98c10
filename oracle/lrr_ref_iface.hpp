// lrr_ref_iface.hpp — glue between the C ABI `lrrr_*` (lrr_ref_capi.cpp) and the translation units that compile the
// REFERENCE's own estimator sources from /root/reference (lrr_ref_cdcl.cpp, lrr_ref_rrl.cpp).  TEST INFRASTRUCTURE ONLY.
//
// oracle/_ref/liblrr_ref.so = the unmodified reference hot path (cdcl_estimation_ICP_matching_all,
// cdcl_estimation_symmetric_ICP_matching_all, rrl_estimation_ICP_matching_all, rrl_estimation_symmetric_ICP_matching_all,
// cdcl_trans_affine / _rigid3d, cdcl_obj_fun_cov) compiled against oracle/vxl_shim (a from-scratch stand-in for the VXL /
// ITK API subset those files use — the libraries themselves are absent from this image).  It exists to PIN the oracle:
// tests/test_reference_pin.py runs oracle and reference on the same inputs, and tests/golden/ref_*.npz are its outputs.
#ifndef LRR_REF_IFACE_HPP_
#define LRR_REF_IFACE_HPP_
#include <cstdint>
#include <vector>
#include "../include/lrr_gpu.h"

struct RefFeat { double loc[3], P[9], cov[9], strength; int shape; double dirs[3][3]; };
struct RefTrans { int kind; double A[9], t[3], c[3]; std::vector<double> cov; };

struct RefEst {
  virtual ~RefEst() {}
  virtual void set_subsets(const int32_t* mi, int64_t nm, const int32_t* fi, int64_t nf) = 0;
  virtual void set_transform(const RefTrans& t) = 0;
  virtual void set_flags(int normalize, int weight_by_strength) = 0;
  virtual bool set_number_points(int dir, int64_t n, int* at_max) = 0;
  virtual void find_closest() = 0;
  virtual bool one_iteration(unsigned it) = 0;
  virtual bool get_transform(int dir, RefTrans& t) = 0;
  virtual int64_t get_matches(int dir, int64_t cap, int32_t* from, int32_t* to, double* w) = 0;
  virtual bool metric(int which, int dir, double* out) = 0;
  virtual bool estimate_ls_backward(int estimate_parameters) = 0;
  virtual bool transfer_cov(int dir, double* S9) = 0;
  virtual void get_state(lrr_est_state* s) = 0;
  virtual int oscillated() = 0;
  virtual bool objcov_eval(int dir, const double* x, double* f, double* g) = 0;
  virtual bool objpar_eval(int dir, const double* theta, double* f, double* g) = 0;
};

// the full CDC estimator's E-step (cdcl_estimation<3,12>::compute_weights, lrr_ref_cdc.cpp)
struct RefCdc {
  virtual ~RefCdc() {}
  virtual void set_transform(const RefTrans& t) = 0;
  virtual bool compute_weights(lrr_cdc_summary* out) = 0;
  virtual int64_t rows() const = 0;
  virtual int64_t entries() const = 0;
  virtual void get_matches(int32_t* row_from, int64_t* row_start, int32_t* to_idx, double* w) = 0;
  virtual bool objcov(const double* x, double* f, double* g) = 0;
  virtual bool objpar(const double* theta, double* f, double* g) = 0;
  virtual void normalisation(double* norm8) = 0;
};
RefCdc* lrr_ref_make_cdc(const std::vector<RefFeat>& moving, const int32_t* mi, int64_t nm, const std::vector<RefFeat>& fixed);

RefEst* lrr_ref_make_cdcl(int kind, int dof, const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed);
RefEst* lrr_ref_make_rrl(int kind, int dof, const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed);
RefEst* lrr_ref_make_icp2d(const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed);   // cdcl_estimation_ICP<2,4> (lrr_ref_icp2d.cpp)
void lrr_ref_projector2d(const double* C3, double* P3);
void lrr_ref_sim2d_roundtrip(double* s, double* ang, double* t2, double* cov16, double rm, double rf, const double* cm2, const double* cf2, int do_unnormalize);
void lrr_ref_set_shuffle(int mode, unsigned long long seed);
// small direct probes of reference functions (known-answer pins of the oracle's restatements)
double lrr_ref_weight_BT(double u);
double lrr_ref_rho(double u);
double lrr_ref_rho_p(double u);
double lrr_ref_mad_scale(const double* v, int64_t n);
void lrr_ref_trans_roundtrip(int kind, double* A9, double* t3, double* c3, double* cov, double rm, double rf, const double* cm, const double* cf, int do_unnormalize);
void lrr_ref_trans_inverse(int kind, double* A9, double* t3, double* c3);
void lrr_ref_recompose(double* R9, double* t3, const double* dR9, const double* dt3);
#endif
