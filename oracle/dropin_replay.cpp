// dropin_replay.cpp — TEST INFRASTRUCTURE.  Proves that include/lrr/lrr_dropin.h is a drop-in for the reference's
// rrl_estimation_symmetric_ICP_matching_all<3,12>: the call sites of rrl/itkLocationRegistration.cxx are replayed VERBATIM, once
// with the driver's own typedef (the reference's class, compiled unmodified from /root/reference against oracle/vxl_shim) and
// once with the typedef pointed at lrr_b200:: (the CUDA backend), on the same reference feature / transform objects, and the
// results are compared.  Built only where /root/reference exists (oracle/Makefile -> oracle/_ref/dropin_replay); run on the GPU
// by tests/test_gpu_dropin.py.  Nothing of this file ships in the product.
#include "vxl_shim/vcl_shim_all.h"
#include "vxl_shim/vnl/algo/vnl_shim_algo.h"
#include "vxl_shim/rsdl/rsdl_kd_tree.h"
#include "vxl_shim/rrel/rrel_muset_obj.h"
#include "vxl_shim/itk/itkMultiThreader.h"
#include <vnl/vnl_math.h>
#include <cdcl/cdcl_trans.txx>
#include <cdcl/cdcl_trans_affine.txx>
#include <rrl/rrl_estimation_ICP_matching_all.txx>
#include <rrl/rrl_estimation_symmetric_ICP_matching_all.txx>
#include "../include/lrr/lrr_dropin.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

static const unsigned int dim = 3, dof = 12;
typedef cdcl_feature_ICP<dim> feature_type;
typedef feature_type::sptr feature_sptr_type;
typedef cdcl_trans<dim, dof> trans_type;
typedef itk::Image<unsigned int, 3> VoronoiMapImageType;

// ---- deterministic synthetic volumes: features on spheres (sheets), lines (tubes) and corner sites ----------------------
static unsigned long long g_rng = 0x9E3779B97F4A7C15ULL;
static double urand() { g_rng = g_rng * 6364136223846793005ULL + 1442695040888963407ULL; return (double)((g_rng >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53); }
static double nrand() { return std::sqrt(-2.0 * std::log(urand() + 1e-300)) * std::cos(2.0 * vnl_math::pi * urand()); }
static vnl_vector_fixed<double, 3> unit(vnl_vector_fixed<double, 3> v) { return v / v.two_norm(); }

static feature_sptr_type make_feature(vnl_vector_fixed<double, 3> const& loc, int shape, vnl_vector_fixed<double, 3> const& e0, vnl_vector_fixed<double, 3> const& e2) {
  vnl_vector_fixed<double, 3> e1 = vnl_cross_3d(e2, e0);
  vnl_matrix_fixed<double, 3, 3> P(0.0);
  vcl_vector< vnl_vector_fixed<double, 3> > dirs;
  for (unsigned r = 0; r < 3; ++r) for (unsigned c = 0; c < 3; ++c) {
    if (shape == 2) P(r, c) = e2[r] * e2[c];                               // sheet: n n^T
    else if (shape == 1) P(r, c) = (r == c ? 1.0 : 0.0) - e0[r] * e0[c];   // tube: I - t t^T
    else P(r, c) = r == c ? 1.0 : 0.0;
  }
  if (shape == 2) dirs.push_back(e2);
  else if (shape == 1) { dirs.push_back(e1); dirs.push_back(e2); }
  else { dirs.push_back(e0); dirs.push_back(e1); dirs.push_back(e2); }
  const double strength = std::exp(4.6 + 0.7 * nrand());
  return new feature_type(loc, strength, P, (feature_type::shape_type)shape, dirs);
}

struct Scene { vcl_vector< vnl_vector_fixed<double, 3> > sc, la, lb, cs; vcl_vector<double> sr; };
static void sample(Scene const& S, unsigned n, vnl_matrix_fixed<double, 3, 3> const& Ainv, vnl_vector_fixed<double, 3> const& shift, double noise, double outliers,
                   double size, vcl_vector<feature_sptr_type>& out) {
  for (unsigned i = 0; i < n; ++i) {
    const double u = urand();
    vnl_vector_fixed<double, 3> p, e0, e2; int shape;
    if (u < 0.55) {
      const unsigned s = (unsigned)(urand() * S.sc.size());
      vnl_vector_fixed<double, 3> d(nrand(), nrand(), nrand()); d = unit(d);
      p = S.sc[s] + S.sr[s] * d; e2 = d; shape = 2;
      vnl_vector_fixed<double, 3> a(nrand(), nrand(), nrand()); a -= dot_product(a, d) * d; e0 = unit(a);
    } else if (u < 0.90) {
      const unsigned s = (unsigned)(urand() * S.la.size());
      const double t = urand();
      p = S.la[s] + t * (S.lb[s] - S.la[s]); e0 = unit(S.lb[s] - S.la[s]); shape = 1;
      vnl_vector_fixed<double, 3> a(nrand(), nrand(), nrand()); a -= dot_product(a, e0) * e0; e2 = unit(a);
    } else {
      const unsigned s = (unsigned)(urand() * S.cs.size());
      p = S.cs[s] + vnl_vector_fixed<double, 3>(0.05 * nrand(), 0.05 * nrand(), 0.05 * nrand()); shape = 0;
      e0 = vnl_vector_fixed<double, 3>(1, 0, 0); e2 = vnl_vector_fixed<double, 3>(0, 0, 1);
    }
    if (p[0] < 0 || p[0] > size || p[1] < 0 || p[1] > size || p[2] < 0 || p[2] > size) { --i; continue; }
    p += vnl_vector_fixed<double, 3>(noise * nrand(), noise * nrand(), noise * nrand());
    if (urand() < outliers) p = vnl_vector_fixed<double, 3>(size * urand(), size * urand(), size * urand());
    p = Ainv * (p - shift) + vnl_vector_fixed<double, 3>(size / 2, size / 2, size / 2);
    // directions: tangent-like vectors map with A^-1, normals with A^-T; re-orthonormalised
    vnl_vector_fixed<double, 3> t0 = unit(Ainv * e0), n2 = e2;   // (small transform: the normal is kept and re-orthogonalised against the mapped tangent)
    n2 -= dot_product(n2, t0) * t0; n2 = unit(n2);
    out.push_back(make_feature(p, shape, t0, n2));
  }
}

static VoronoiMapImageType::Pointer ideal_map(vcl_vector<feature_sptr_type> const& set, vbl_smart_ptr<rsdl_kd_tree>& kd) {
  std::vector<rsdl_point> pts; rsdl_point pt(3, 0);
  for (size_t i = 0; i < set.size(); ++i) { pt.set_cartesian(set[i]->location_); pts.push_back(pt); }
  kd = new rsdl_kd_tree(pts);
  rsdl_kd_tree* tree = kd.as_pointer();
  VoronoiMapImageType::Pointer m = VoronoiMapImageType::New();
  m->SetLookup([tree](const double* p) -> unsigned int {
    rsdl_point q(3, 0); q.set_cartesian(p);
    std::vector<rsdl_point> cp; std::vector<int> ci;
    tree->n_nearest(q, 1, cp, ci);
    return (unsigned int)ci[0];
  });
  return m;
}

struct Outcome {
  double rank_angles[4]; unsigned int iterations; bool converged, oscillated;
  vnl_matrix_fixed<double, 3, 3> A, Ab, covJ; vnl_vector_fixed<double, 3> t; vnl_matrix_fixed<double, dof, dof> cov, cov_b;
  double m[6], rms;
};

// FeaturesInROIs (rrl/itkLocationRegistration.cxx:1324-1360): deep clones of the features inside the cube
static void FeaturesInROI(vcl_vector<feature_sptr_type> const& all, vnl_vector_fixed<double, 3> const& c, double half, vcl_vector<feature_sptr_type>& inside) {
  inside.clear();
  for (size_t i = 0; i < all.size(); ++i) {
    bool in = true;
    for (unsigned d = 0; d < 3; ++d) if (all[i]->location_[d] < c[d] - half || all[i]->location_[d] > c[d] + half) in = false;
    if (in) inside.push_back(new feature_type(*all[i]));
  }
}

template <class estimation_symmetric_type>
static Outcome driver_replay(vcl_vector< vcl_vector<feature_sptr_type> > const& m_FeaturesMovingAll, vcl_vector< vcl_vector<feature_sptr_type> > const& m_FeaturesFixedAll,
                             VoronoiMapImageType::Pointer const& m_MovingVoronoiMap, VoronoiMapImageType::Pointer const& m_FixedVoronoiMap,
                             trans_type::sptr const& transform0, vnl_vector_fixed<double, 3> const& cm, vnl_vector_fixed<double, 3> const& cf) {
  Outcome R;
  typename estimation_symmetric_type::sptr estimation;   // rrl/itkLocationRegistration.h:337
  for (int pass = 0; pass < 2; ++pass) {                 // second pass: the "already ran" branch with a larger ROI (:1635-1637)
    vcl_vector<feature_sptr_type> moving_inside, fixed_inside;
    FeaturesInROI(m_FeaturesMovingAll[0], cm, pass == 0 ? 25.0 : 40.0, moving_inside);
    FeaturesInROI(m_FeaturesFixedAll[0], cf, pass == 0 ? 25.0 : 40.0, fixed_inside);
    trans_type::sptr transform = pass == 0 ? transform0->clone() : estimation->get_transform();
    // ---- rrl/itkLocationRegistration.cxx:1632-1644, verbatim ----
    if( moving_inside.size() < 1000 || fixed_inside.size() < 1000 ) { std::printf("ROI too small: %zu %zu\n", moving_inside.size(), fixed_inside.size()); std::exit(2); }
    if( estimation ) {
      estimation->set_features( moving_inside, fixed_inside );
      estimation->set_transform( transform );
    }
    else {
      estimation = new estimation_symmetric_type( m_FeaturesMovingAll[0], m_FeaturesFixedAll[0], transform, transform->inverse(), moving_inside, fixed_inside );
      estimation->weight_by_strength( true );
      estimation->normalize_matches();
      estimation->set_voronoi_maps( m_MovingVoronoiMap, m_FixedVoronoiMap );
    }
    // ---- :2303-2313, verbatim ----
    estimation->find_closest_euclidean();
    double sheet_angles = estimation->sheet_angles() * 180.0 / vnl_math::pi;
    double tube_angles = estimation->tube_angles() * 180.0 / vnl_math::pi;
    double sheet_angles_backward = estimation->sheet_angles_backward() * 180.0 / vnl_math::pi;
    double tube_angles_backward = estimation->tube_angles_backward() * 180.0 / vnl_math::pi;
    if (pass == 0) { R.rank_angles[0] = sheet_angles; R.rank_angles[1] = tube_angles; R.rank_angles[2] = sheet_angles_backward; R.rank_angles[3] = tube_angles_backward; }
    // ---- :2418-2446: the hot loop ----
    bool converged = false;
    unsigned int iteration = 0;
    while( !converged && iteration < 70 ) {
      converged = estimation->one_iteration( 0, 0, iteration );
      ++iteration;
    }
    R.iterations = iteration; R.converged = converged;
    R.oscillated = estimation->oscillated();   // :2504
  }
  // ---- ComputeDecisionMeasurements, :1828-1845, verbatim ----
  double sheet_angles = estimation->sheet_angles() * 180.0 / vnl_math::pi;
  double tube_angles = estimation->tube_angles() * 180.0 / vnl_math::pi;
  double weighted_error = estimation->weighted_error();
  estimation->estimate_LS_backward( false );
  double sheet_angles_backward = estimation->sheet_angles_backward() * 180.0 / vnl_math::pi;
  double tube_angles_backward = estimation->tube_angles_backward() * 180.0 / vnl_math::pi;
  double weighted_error_backward = estimation->weighted_error_backward();
  R.m[0] = sheet_angles; R.m[1] = tube_angles; R.m[2] = weighted_error; R.m[3] = sheet_angles_backward; R.m[4] = tube_angles_backward; R.m[5] = weighted_error_backward;
  // ---- WriteResults, :1952-1966 ----
  estimation->compute_transfer_error_covariance();
  trans_type::sptr T = estimation->get_transform(), Tb = estimation->get_transform_backward();
  R.A = T->get_A(); R.t = T->get_translation(); R.cov = T->get_covariance(); R.covJ = T->get_covarianceJ();
  R.Ab = Tb->get_A(); R.cov_b = Tb->get_covariance();
  R.rms = estimation->RMS_error();
  return R;
}

template <class M> static double rel(M const& a, M const& b) { return (a - b).frobenius_norm() / b.frobenius_norm(); }

int main() {
  const double size = 110.0;
  Scene S;
  for (int i = 0; i < 12; ++i) { S.sc.push_back(vnl_vector_fixed<double, 3>(size * (0.1 + 0.8 * urand()), size * (0.1 + 0.8 * urand()), size * (0.1 + 0.8 * urand()))); S.sr.push_back(size * (0.12 + 0.25 * urand())); }
  for (int i = 0; i < 25; ++i) { S.la.push_back(vnl_vector_fixed<double, 3>(size * urand(), size * urand(), size * urand())); S.lb.push_back(vnl_vector_fixed<double, 3>(size * urand(), size * urand(), size * urand())); }
  for (int i = 0; i < 600; ++i) S.cs.push_back(vnl_vector_fixed<double, 3>(size * (0.05 + 0.9 * urand()), size * (0.05 + 0.9 * urand()), size * (0.05 + 0.9 * urand())));
  // ground truth: small affine; moving = T_gt^-1 (fixed scene)
  vnl_matrix_fixed<double, 3, 3> Agt; Agt.set_identity();
  Agt(0, 1) = 0.03; Agt(1, 0) = -0.025; Agt(0, 2) = -0.02; Agt(2, 0) = 0.018; Agt(1, 1) = 1.02; Agt(2, 2) = 0.985;
  vnl_vector_fixed<double, 3> c0(size / 2, size / 2, size / 2), tgt(3.0, -2.0, 1.5);
  vnl_matrix_fixed<double, 3, 3> I3; I3.set_identity();
  vcl_vector< vcl_vector<feature_sptr_type> > m_FeaturesMovingAll(1), m_FeaturesFixedAll(1);
  sample(S, 40000, I3, c0, 0.0, 0.0, size, m_FeaturesFixedAll[0]);   // fixed volume = the scene itself (p = I (p - c0) + c0)
  sample(S, 40000, vnl_inverse(Agt), c0 + tgt, 0.3, 0.15, size, m_FeaturesMovingAll[0]);
  // initial transform: the ground truth perturbed (y = A (x - c0) + t)
  vnl_matrix_fixed<double, 3, 3> A0 = Agt; A0(0, 1) += 0.02; A0(2, 1) -= 0.015;
  vnl_vector_fixed<double, 3> t0 = c0 + tgt + vnl_vector_fixed<double, 3>(1.5, -1.0, 2.0);
  trans_type::sptr transform0 = new cdcl_trans_affine<dim, dof>(A0, t0, c0);
  vnl_vector_fixed<double, 3> cf(55.0, 50.0, 60.0);
  vnl_vector_fixed<double, 3> cm = vnl_inverse(Agt) * (cf - c0 - tgt) + c0;
  vbl_smart_ptr<rsdl_kd_tree> kdm, kdf;
  VoronoiMapImageType::Pointer m_MovingVoronoiMap = ideal_map(m_FeaturesMovingAll[0], kdm), m_FixedVoronoiMap = ideal_map(m_FeaturesFixedAll[0], kdf);

  Outcome ref = driver_replay< ::rrl_estimation_symmetric_ICP_matching_all<dim, dof> >(m_FeaturesMovingAll, m_FeaturesFixedAll, m_MovingVoronoiMap, m_FixedVoronoiMap, transform0, cm, cf);
  Outcome gpu = driver_replay< lrr_b200::rrl_estimation_symmetric_ICP_matching_all<dim, dof> >(m_FeaturesMovingAll, m_FeaturesFixedAll, m_MovingVoronoiMap, m_FixedVoronoiMap, transform0, cm, cf);

  bool ok = ref.iterations == gpu.iterations && ref.converged == gpu.converged && ref.oscillated == gpu.oscillated;
  std::printf("iterations %u / %u converged %d / %d oscillated %d / %d\n", ref.iterations, gpu.iterations, (int)ref.converged, (int)gpu.converged, (int)ref.oscillated, (int)gpu.oscillated);
  double worst = 0;
  for (int k = 0; k < 4; ++k) worst = std::max(worst, std::fabs(ref.rank_angles[k] - gpu.rank_angles[k]));
  for (int k = 0; k < 6; ++k) worst = std::max(worst, std::fabs(ref.m[k] - gpu.m[k]) / std::max(1.0, std::fabs(ref.m[k])));
  worst = std::max(worst, std::fabs(ref.rms - gpu.rms) / std::max(1.0, std::fabs(ref.rms)));
  std::printf("ranking angles (deg) ref %.9f %.9f %.9f %.9f\n", ref.rank_angles[0], ref.rank_angles[1], ref.rank_angles[2], ref.rank_angles[3]);
  std::printf("decision values      ref %.9f %.9f %.9f %.9f %.9f %.9f  worst metric difference %.3e\n", ref.m[0], ref.m[1], ref.m[2], ref.m[3], ref.m[4], ref.m[5], worst);
  const double eA = rel(gpu.A, ref.A), et = (gpu.t - ref.t).two_norm() / ref.t.two_norm(), ec = rel(gpu.cov, ref.cov), eb = rel(gpu.Ab, ref.Ab), ecb = rel(gpu.cov_b, ref.cov_b), ej = rel(gpu.covJ, ref.covJ);
  std::printf("relative differences: A %.3e t %.3e cov %.3e A_backward %.3e cov_backward %.3e covarianceJ %.3e\n", eA, et, ec, eb, ecb, ej);
  ok = ok && worst < 1e-8 && eA < 1e-5 && et < 1e-5 && ec < 1e-5 && eb < 1e-5 && ecb < 1e-5 && ej < 1e-5;
  std::printf(ok ? "DROPIN OK\n" : "DROPIN MISMATCH\n");
  return ok ? 0 : 1;
}
