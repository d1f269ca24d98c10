// Host-side C++ example and self-check of the full CDC estimator of include/lrr/lrr_estimators.h (cdcl_estimation<3,12>: EM over
// many-to-many matches, cdcl/cdcl_estimation.txx:66-204), driven the way cdcl/examples/estimate.cxx:99-101,113-132 drives it: initialize,
// then one_iteration until converged.  Synthetic data as demo_register.cpp (points on three spheres, unit location covariances);
// exits 0 when the E-step found many-to-many matches in every iteration and the recovered transform maps the moving set onto the fixed one.
#include <cmath>
#include <cstdio>
#include <random>
#include "../../include/lrr/lrr_estimators.h"

using namespace lrr_b200;
typedef cdcl_feature_with_shape<3> feature_type;

int main() {
  const unsigned int dim = 3, dof = 12;
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(0.0, 100.0);
  std::vector<feature_type::sptr> fixed, moving;
  const double A[9] = {1.02, 0.03, -0.02, -0.025, 0.99, 0.04, 0.015, -0.035, 1.01}, t[3] = {53.0, 48.5, 51.0}, c[3] = {50, 50, 50};
  double Ai[9];
  {  // 3x3 inverse
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    Ai[0] = (A[4] * A[8] - A[5] * A[7]) / det; Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det; Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    Ai[3] = (A[5] * A[6] - A[3] * A[8]) / det; Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det; Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    Ai[6] = (A[3] * A[7] - A[4] * A[6]) / det; Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det; Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
  }
  for (int i = 0; i < 3000; ++i) {
    feature_type::sptr q(new feature_type), p(new feature_type);
    const int sph = i % 3;
    const double cs[3][3] = {{30, 40, 50}, {65, 55, 35}, {50, 30, 70}}, rad[3] = {22.0, 18.0, 25.0};
    std::normal_distribution<double> G(0.0, 1.0);
    vec3 n = {G(rng), G(rng), G(rng)};
    const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int d = 0; d < 3; ++d) n[d] /= nn;
    vec3 x = {cs[sph][0] + rad[sph] * n[0], cs[sph][1] + rad[sph] * n[1], cs[sph][2] + rad[sph] * n[2]};
    (void)U;
    q->location_ = x; q->strength_ = 50.0 + U(rng); q->shape_ = feature_type::SHEET; q->directions_.push_back(n);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { q->error_projector_[3 * a + b] = n[a] * n[b]; q->covariance_[3 * a + b] = (a == b) ? 1.0 : 0.0; }
    *p = *q;
    // 1e-4 mm of noise: with exact data the residuals collapse to rounding noise and the reference's scale
    // sqrt(sum r^2 / n) can become exactly 0 (division by zero in weight_BT( r^2 / scale^2 ), cdcl :745-747)
    for (int r = 0; r < 3; ++r) p->location_[r] = Ai[3 * r] * (x[0] - t[0]) + Ai[3 * r + 1] * (x[1] - t[1]) + Ai[3 * r + 2] * (x[2] - t[2]) + c[r] + 1e-4 * G(rng);
    fixed.push_back(q); moving.push_back(p);
  }
  device_context ctx(0);
  cdcl_trans<dim, dof>::sptr init(new cdcl_trans<dim, dof>(mat3{1, 0, 0, 0, 1, 0, 0, 0, 1}, vec3{52.0, 49.5, 50.0}, vec3{50, 50, 50}));
  cdcl_estimation<dim, dof> estimation(ctx, moving, fixed, init);
  estimation.weight_by_strength(true);
  estimation.normalize_matches();
  estimation.initialize();
  bool converged = false, many = true; unsigned int iteration = 0;
  while (!converged && iteration < 12) {
    converged = estimation.one_iteration(0, 0, iteration); ++iteration;
    const lrr_cdc_summary& s = estimation.last_weights();
    many = many && s.entries > s.rows && s.rows > 0;
    if (std::getenv("LRR_DEMO_VERBOSE")) std::printf("it %u rows %lld entries %lld max %d avg_largest_eval %.4f\n", iteration, (long long)s.rows, (long long)s.entries, s.max_matches, s.avg_largest_eval);
  }
  cdcl_trans<dim, dof>::sptr T = estimation.get_transform();
  T->print();
  double err = 0;
  for (size_t i = 0; i < moving.size(); ++i) {
    const vec3 y = T->map_loc(moving[i]->location_);
    for (int d = 0; d < 3; ++d) err = std::fmax(err, std::fabs(y[d] - fixed[i]->location_[d]));
  }
  const int res = estimation.switch_resolution(std::vector<double>{1.0, 2.0}, 1, true);
  std::printf("EM iterations %u converged %d many-to-many %d max map error %.3e resolution %d\n", iteration, (int)converged, (int)many, err, res);
  return (many && err < 0.5) ? 0 : 1;
}
