// Stand-in for VXL contrib/rpl/rrel: rrel_objective, rrel_muset_obj::scale.  TEST INFRASTRUCTURE ONLY.
// MUSE itself is restated once in ../../lrr_oracle_muse.hpp (third-party, unpinned upstream).
#ifndef rrel_muset_obj_shim_h_
#define rrel_muset_obj_shim_h_
#include <vector>
#include "../../lrr_oracle_muse.hpp"
class rrel_objective {
 public:
  typedef std::vector<double>::const_iterator vect_const_iter;
  typedef std::vector<double>::iterator vect_iter;
  virtual ~rrel_objective() {}
  virtual double scale(vect_const_iter begin, vect_const_iter end) const = 0;
  virtual bool can_estimate_scale() const { return true; }
};
class rrel_muset_obj : public rrel_objective {
 public:
  rrel_muset_obj(int /*max_lookup_table_size*/, bool use_sk_refine = true) : use_sk_refine_(use_sk_refine) {}
  virtual double scale(vect_const_iter begin, vect_const_iter end) const { return lrro::muse_scale(std::vector<double>(begin, end)); }
 private:
  bool use_sk_refine_;
};
#endif
