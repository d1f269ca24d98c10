// lrr_ref_rrl.cpp — compiles the reference's rrl estimators UNMODIFIED from /root/reference.  TEST INFRASTRUCTURE ONLY.
// Reference sources pulled in by #include (never copied): rrl/rrl_estimation_ICP_matching_all.txx,
// rrl/rrl_estimation_symmetric_ICP_matching_all.txx, cdcl/cdcl_trans*.txx (+ headers).  A separate translation unit because
// the cdcl and rrl .txx files each define ::rrel_util_median_abs_dev_scale (upstream compiles them apart, too).
// The Voronoi label image of the reference (third-party ITK filter output, Q5) is replaced by the ideal map of
// vxl_shim/itk/itkImage.h whose lookup is the exact nearest neighbour (kd-tree stand-in, ties to the lowest index).
#include "vxl_shim/vcl_shim_all.h"
#include "vxl_shim/vnl/algo/vnl_shim_algo.h"
#include "vxl_shim/rsdl/rsdl_kd_tree.h"
#include "vxl_shim/rrel/rrel_muset_obj.h"
#include "vxl_shim/itk/itkMultiThreader.h"
#include <unordered_map>
#include <cstdlib>
#include <thread>
#define protected public
#define private public
#include <cdcl/cdcl_trans.txx>
#include <cdcl/cdcl_trans_affine.txx>
#include <rrl/rrl_estimation_ICP_matching_all.txx>
#include <rrl/rrl_estimation_symmetric_ICP_matching_all.txx>
#undef protected
#undef private
#include "lrr_ref_impl.hpp"

#include <chrono>
static double g_raster_seconds = 0.0;   // wall time spent building rasterised Voronoi maps (bench.py subtracts it: an input file upstream)
extern "C" double lrrr_voronoi_raster_seconds() { return g_raster_seconds; }

struct RrlPolicy {
  typedef cdcl_feature_ICP<3> feature_type;
  typedef itk::Image<unsigned int, 3> map_type;
  template <unsigned dof> struct fwd { typedef rrl_estimation_ICP_matching_all<3, dof> type; };
  template <unsigned dof> struct sym { typedef rrl_estimation_symmetric_ICP_matching_all<3, dof> type; };
  struct extra { vbl_smart_ptr<rsdl_kd_tree> kd_fixed, kd_moving; map_type::Pointer map_fixed, map_moving; };
  static feature_type::sptr make_feature(const RefFeat& f) {
    std::vector<ref_vec3> dirs;
    int nd = f.shape == 2 ? 1 : (f.shape == 1 ? 2 : 3);
    for (int a = 0; a < nd; ++a) dirs.push_back(ref_vec3(f.dirs[a]));
    return new feature_type(ref_vec3(f.loc), f.strength, ref_mat3(f.P), (feature_type::shape_type)f.shape, dirs);
  }
  template <class V> static map_type::Pointer ideal_map(const V& set, vbl_smart_ptr<rsdl_kd_tree>& kd) {
    std::vector<rsdl_point> pts; rsdl_point pt(3, 0);
    for (size_t i = 0; i < set.size(); ++i) { pt.set_cartesian(set[i]->location_); pts.push_back(pt); }
    kd = new rsdl_kd_tree(pts);
    rsdl_kd_tree* tree = kd.as_pointer();
    map_type::Pointer m = map_type::New();
    if (const char* mode = std::getenv("LRR_REF_VORONOI"); mode && std::string(mode) == "raster") {
      // TIMING ONLY (bench.py's second cpu_baseline figure): the stock reference's O(1) lookup in a rasterised 1 mm label volume
      // (rrl/examples/compute_distance_map.cxx builds it offline; here every voxel gets the exact nearest feature of its centre,
      // all host threads, untimed set-up like the reference's input file)
      const auto t_begin = std::chrono::steady_clock::now();
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (size_t i = 0; i < set.size(); ++i) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], set[i]->location_[d]); hi[d] = std::max(hi[d], set[i]->location_[d]); }
      unsigned long dims[3]; double org[3];
      for (int d = 0; d < 3; ++d) { org[d] = std::floor(lo[d]) - 8.0; dims[d] = (unsigned long)(std::ceil(hi[d]) + 8.0 - org[d]) + 1; }
      m->SetRaster(dims, org, 1.0);
      std::vector<unsigned int>& data = m->Data();
      unsigned nt = std::max(1u, std::thread::hardware_concurrency());
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
        rsdl_kd_tree own(pts);   // (the stand-in's search keeps its scratch in the tree object: one tree per thread)
        rsdl_kd_tree* tree = &own;
        rsdl_point q(3, 0); std::vector<rsdl_point> cp; std::vector<int> ci; double p[3];
        for (unsigned long z = t; z < dims[2]; z += nt) for (unsigned long y = 0; y < dims[1]; ++y) for (unsigned long x = 0; x < dims[0]; ++x) {
          p[0] = org[0] + (double)x; p[1] = org[1] + (double)y; p[2] = org[2] + (double)z;
          q.set_cartesian(p); tree->n_nearest(q, 1, cp, ci);
          data[(z * dims[1] + y) * dims[0] + x] = (unsigned int)ci[0];
        }
      });
      for (auto& x : th) x.join();
      g_raster_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
      return m;
    }
    m->SetLookup([tree](const double* p) -> unsigned int {
      rsdl_point q(3, 0); q.set_cartesian(p);
      std::vector<rsdl_point> cp; std::vector<int> ci;
      tree->n_nearest(q, 1, cp, ci);
      return (unsigned int)ci[0];
    });
    return m;
  }
  template <class F, class V> static void after_construct_fwd(extra& x, F& f, const V& fixed_all) {
    x.map_fixed = ideal_map(fixed_all, x.kd_fixed);
    f.set_fixed_voronoi_map(x.map_fixed);
  }
  template <class S, class V> static void after_construct_sym(extra& x, S& s, const V& moving_all, const V& fixed_all) {
    x.map_fixed = ideal_map(fixed_all, x.kd_fixed); x.map_moving = ideal_map(moving_all, x.kd_moving);
    s.set_voronoi_maps(x.map_moving, x.map_fixed);
  }
  template <unsigned dof, class F> static bool objcov(F&, const double*, double*, double*) { return false; }
  template <unsigned dof, class F> static bool objpar(F&, const double*, double*, double*) { return false; }
};

RefEst* lrr_ref_make_rrl(int kind, int dof, const std::vector<RefFeat>& moving, const std::vector<RefFeat>& fixed) {
  if (dof != 12) return 0;   // Q6: the rrl accumulation hard-codes affine columns
  return new RefEstImpl<RrlPolicy, 12>(kind == LRR_RRL_SYM, moving, fixed);
}
