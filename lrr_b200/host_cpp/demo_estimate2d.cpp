// Counterpart of register_points() in cdcl/examples/estimate.cxx:45-187 for `-method 4` on planar data (BASELINE.json
// configs[0]): cdcl_normalize_data on both sets, transform->normalize, cdcl_estimation_ICP<2,4>::run, transform->unnormalize —
// written against the reference's class and method names (include/lrr/lrr_estimators.h).  Synthetic data: points on two strokes
// and an arc with tangent-aligned covariances; moving = inverse image under a known similarity + noise.  Exits 0 when the
// recovered similarity matches.
#include <cmath>
#include <cstdio>
#include <random>
#include "../../include/lrr/lrr_estimators.h"

using namespace lrr_b200;

int main() {
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::normal_distribution<double> G(0.0, 1.0);
  const double s_gt = 1.07, ang_gt = 0.2, t_gt[2] = {14.0, -9.0};
  std::vector<cdcl_feature2::sptr> fixed, moving;
  for (int i = 0; i < 3000; ++i) {
    // a point with its tangent on one of three curves, in "pixel" units
    double x, y, tx, ty;
    const double u = U(rng);
    switch (i % 3) {
      case 0: x = 100 + 200 * u; y = 120 + 30 * u; tx = 200; ty = 30; break;
      case 1: x = 180 - 20 * u; y = 60 + 220 * u; tx = -20; ty = 220; break;
      default: { const double th = 3.0 * u; x = 220 + 70 * std::cos(th); y = 200 + 70 * std::sin(th); tx = -std::sin(th); ty = std::cos(th); }
    }
    const double tn = std::sqrt(tx * tx + ty * ty); tx /= tn; ty /= tn;
    const double st = 6.0, sn = 0.8;
    const mat2 C = {st * st * tx * tx + sn * sn * ty * ty, (st * st - sn * sn) * tx * ty, (st * st - sn * sn) * tx * ty, st * st * ty * ty + sn * sn * tx * tx};
    fixed.push_back(cdcl_feature2::sptr(new cdcl_feature2(vec2{x + 0.2 * G(rng), y + 0.2 * G(rng)}, 1.0, C)));
    // moving = (sR)^-1 (y - t), covariance rotated back
    const double c = std::cos(ang_gt), s = std::sin(ang_gt);
    const double dx = x - t_gt[0], dy = y - t_gt[1];
    const double mx = (c * dx + s * dy) / s_gt + 0.2 * G(rng), my = (-s * dx + c * dy) / s_gt + 0.2 * G(rng);
    const double mtx = c * tx + s * ty, mty = -s * tx + c * ty;
    const mat2 Cm = {st * st * mtx * mtx + sn * sn * mty * mty, (st * st - sn * sn) * mtx * mty, (st * st - sn * sn) * mtx * mty, st * st * mty * mty + sn * sn * mtx * mtx};
    moving.push_back(cdcl_feature2::sptr(new cdcl_feature2(vec2{mx, my}, 1.0, Cm)));
  }
  cdcl_trans_similarity2d::sptr transform(new cdcl_trans_similarity2d(1.03, 0.26, vec2{10.0, -5.0}));   // initial guess
  vec2 center_fixed, center_moving; double avg_rad_fixed, avg_rad_moving;
  cdcl_normalize_data(fixed, center_fixed, avg_rad_fixed);                                  // estimate.cxx:73-74
  cdcl_normalize_data(moving, center_moving, avg_rad_moving);
  transform->normalize(avg_rad_moving, avg_rad_fixed, center_moving, center_fixed);        // :77
  device_context ctx(0);
  cdcl_estimation_ICP<2, 4> estimation(ctx, moving, fixed, transform);                     // :99-101
  const bool converged = estimation.run();                                                 // :128
  transform = estimation.get_transform();
  const double rms = estimation.RMS_error();
  transform->unnormalize(avg_rad_moving, avg_rad_fixed, center_moving, center_fixed);      // :183
  std::printf("Final transform:\n"); transform->print();
  std::printf("converged %d RMS (normalised units) %.4e  expected s %.4f ang %.4f t %.3f, %.3f\n", (int)converged, rms, s_gt, ang_gt, t_gt[0], t_gt[1]);
  const bool ok = std::fabs(transform->s_ - s_gt) < 5e-3 && std::fabs(transform->ang_ - ang_gt) < 5e-3 && std::fabs(transform->t_[0] - t_gt[0]) < 1.0 &&
                  std::fabs(transform->t_[1] - t_gt[1]) < 1.0;
  return ok ? 0 : 1;
}
