// Stand-in for VXL contrib/rpl/rsdl: rsdl_point, rsdl_kd_tree::n_nearest, rsdl_dist.  TEST INFRASTRUCTURE ONLY.
// Semantics pinned here (upstream version unpinned): exact Euclidean k-NN, squared distance
// accumulated left to right in fp64 (d0*d0 + d1*d1 + d2*d2), ties broken towards the LOWEST point index.
#ifndef rsdl_kd_tree_shim_h_
#define rsdl_kd_tree_shim_h_
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <utility>
#include <vector>
#include "../vbl/vbl_ref_count.h"
#include "../vnl/vnl_shim_core.h"

class rsdl_point {
 public:
  rsdl_point(unsigned Nc = 0, unsigned Na = 0) : c_(Nc, 0.0), a_(Na, 0.0) {}
  rsdl_point(const vnl_vector<double>& c) : c_(c.begin(), c.end()) {}
  rsdl_point(const vnl_vector<double>& c, const vnl_vector<double>& a) : c_(c.begin(), c.end()), a_(a.begin(), a.end()) {}
  unsigned num_cartesian() const { return (unsigned)c_.size(); }
  unsigned num_angular() const { return (unsigned)a_.size(); }
  void resize(unsigned Nc, unsigned Na) { c_.assign(Nc, 0.0); a_.assign(Na, 0.0); }
  void set_cartesian(const vnl_vector<double>& v) { c_.assign(v.begin(), v.end()); }
  template <unsigned n> void set_cartesian(const vnl_vector_fixed<double, n>& v) { c_.assign(v.begin(), v.end()); }
  void set_cartesian(const double* p) { for (auto& x : c_) x = *p++; }
  double cartesian(unsigned i) const { return c_[i]; }
  double& cartesian(unsigned i) { return c_[i]; }
  double angular(unsigned i) const { return a_[i]; }
  const std::vector<double>& coords() const { return c_; }
 private:
  std::vector<double> c_, a_;
};

inline double rsdl_dist_sq(const rsdl_point& p, const rsdl_point& q) {
  double s = 0; for (unsigned i = 0; i < p.num_cartesian(); ++i) { double d = p.cartesian(i) - q.cartesian(i); s += d * d; } return s;
}
inline double rsdl_dist(const rsdl_point& p, const rsdl_point& q) { return std::sqrt(rsdl_dist_sq(p, q)); }

// axis-aligned box (rsdl_bounding_box): only min/max points are used by the call sites (cdcl/cdcl_utils.txx:724,901)
class rsdl_bounding_box {
 public:
  rsdl_bounding_box(const rsdl_point& mn, const rsdl_point& mx) : min_(mn), max_(mx) {}
  const rsdl_point& min_point() const { return min_; }
  const rsdl_point& max_point() const { return max_; }
 private:
  rsdl_point min_, max_;
};

class rsdl_kd_tree : public vbl_ref_count {
 public:
  rsdl_kd_tree(const std::vector<rsdl_point>& pts, double /*min_angle*/ = 0, int points_per_leaf = 8)
      : dim_(pts.empty() ? 0 : pts[0].num_cartesian()), n_((unsigned)pts.size()), leaf_(points_per_leaf < 1 ? 1 : points_per_leaf) {
    xyz_.resize((size_t)n_ * dim_);
    for (unsigned i = 0; i < n_; ++i) for (unsigned d = 0; d < dim_; ++d) xyz_[(size_t)i * dim_ + d] = pts[i].cartesian(d);
    perm_.resize(n_); std::iota(perm_.begin(), perm_.end(), 0);
    if (n_) build(0, n_);
  }
  // k nearest points of `query`, closest first.  use_heap / max_leaves (approximate search) are ignored: always exact.
  void n_nearest(const rsdl_point& query, int k, std::vector<rsdl_point>& cpoints, std::vector<int>& cindices,
                 bool /*use_heap*/ = false, int /*max_leaves*/ = -1) {
    std::vector<double> q(dim_); for (unsigned d = 0; d < dim_; ++d) q[d] = query.cartesian(d);
    best_.clear(); k_ = (unsigned)std::min<long>(k, n_);
    if (k_) search(0, q.data());
    cindices.resize(best_.size()); cpoints.resize(best_.size());
    for (size_t i = 0; i < best_.size(); ++i) {
      cindices[i] = best_[i].second;
      rsdl_point p(dim_, 0); p.set_cartesian(&xyz_[(size_t)best_[i].second * dim_]); cpoints[i] = p;
    }
  }
  void points_in_radius(const rsdl_point& query, double radius, std::vector<rsdl_point>& points, std::vector<int>& indices) {
    points.clear(); indices.clear();
    rsdl_point p(dim_, 0);
    for (unsigned i = 0; i < n_; ++i) { p.set_cartesian(&xyz_[(size_t)i * dim_]); if (rsdl_dist_sq(query, p) <= radius * radius) { points.push_back(p); indices.push_back((int)i); } }
  }
  void points_in_bounding_box(const rsdl_bounding_box& box, std::vector<rsdl_point>& points, std::vector<int>& indices) {
    points.clear(); indices.clear();
    rsdl_point p(dim_, 0);
    for (unsigned i = 0; i < n_; ++i) {
      bool in = true;
      for (unsigned d = 0; d < dim_ && in; ++d) { const double v = xyz_[(size_t)i * dim_ + d]; in = v >= box.min_point().cartesian(d) && v <= box.max_point().cartesian(d); }
      if (in) { p.set_cartesian(&xyz_[(size_t)i * dim_]); points.push_back(p); indices.push_back((int)i); }
    }
  }
 private:
  struct Node { unsigned lo, hi; int left, right; unsigned axis; double split; };
  int build(unsigned lo, unsigned hi) {
    int id = (int)nodes_.size(); nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.0});
    if (hi - lo <= leaf_) return id;
    unsigned axis = 0; double ext = -1;
    for (unsigned d = 0; d < dim_; ++d) {
      double mn = std::numeric_limits<double>::infinity(), mx = -mn;
      for (unsigned i = lo; i < hi; ++i) { double v = xyz_[(size_t)perm_[i] * dim_ + d]; mn = std::min(mn, v); mx = std::max(mx, v); }
      if (mx - mn > ext) { ext = mx - mn; axis = d; }
    }
    unsigned mid = (lo + hi) / 2;
    std::nth_element(perm_.begin() + lo, perm_.begin() + mid, perm_.begin() + hi,
                     [&](int a, int b) { return xyz_[(size_t)a * dim_ + axis] < xyz_[(size_t)b * dim_ + axis]; });
    double split = xyz_[(size_t)perm_[mid] * dim_ + axis];
    int l = build(lo, mid), r = build(mid, hi);
    nodes_[id].left = l; nodes_[id].right = r; nodes_[id].axis = axis; nodes_[id].split = split;
    return id;
  }
  void offer(double d2, int idx) {
    std::pair<double, int> c(d2, idx);
    if (best_.size() == k_ && !(c < best_.back())) return;
    best_.insert(std::upper_bound(best_.begin(), best_.end(), c), c);   // (distance, index) lexicographic: lowest index wins ties
    if (best_.size() > k_) best_.pop_back();
  }
  void search(int id, const double* q) {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (unsigned i = nd.lo; i < nd.hi; ++i) {
        const double* p = &xyz_[(size_t)perm_[i] * dim_];
        double s = 0; for (unsigned d = 0; d < dim_; ++d) { double e = q[d] - p[d]; s += e * e; }
        offer(s, perm_[i]);
      }
      return;
    }
    double diff = q[nd.axis] - nd.split;
    int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
    search(near, q);
    // '<=' keeps equal-distance candidates on the far side reachable so that the lowest index wins ties
    if (best_.size() < k_ || diff * diff <= best_.back().first) search(far, q);
  }
  unsigned dim_, n_, leaf_, k_ = 0;
  std::vector<double> xyz_;
  std::vector<int> perm_;
  std::vector<Node> nodes_;
  std::vector<std::pair<double, int> > best_;
};
#endif
