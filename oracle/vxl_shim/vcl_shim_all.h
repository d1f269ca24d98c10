// vcl_shim_all.h — stand-in for VXL's vcl_* compatibility headers (vcl_vector.h, vcl_iostream.h, ...).
//
// TEST INFRASTRUCTURE ONLY.  VXL is not installed in this image; this directory is a small
// from-scratch implementation of the part of the VXL / ITK *API* that the reference's hot-path sources use,
// so that those sources compile UNMODIFIED from /root/reference (recipe: oracle/Makefile, target _ref).
// Nothing here is copied from VXL; the numerical routines (SVD, eigen-decomposition, k-d tree, MUSE) are
// independent implementations of the published algorithms.
//
// VXL's vcl layer maps vcl_xxx to std::xxx; that is all this header does, plus two hooks:
//  * vcl_cout / vcl_cerr go to a null stream unless LRR_REF_VERBOSE is set (the estimators print per-iteration
//    progress, cdcl/cdcl_estimation_ICP_matching_all.txx:407-408);
//  * vcl_random_shuffle is routed through vxl_shim::shuffle_hook so that the permutation (unseeded rand() in the
//    reference) is an injectable input exactly as in the oracle and the CUDA backend.
#ifndef vcl_shim_all_h_
#define vcl_shim_all_h_

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

namespace vxl_shim {

struct null_buf : std::streambuf { int overflow(int c) override { return c; } };
inline std::ostream& out() {
  // the image links libstdc++ statically into each .so: make sure its stream / locale globals exist before first use
  static std::ios_base::Init ios_init_guard;
  static null_buf nb;
  static std::ostream null_stream(&nb);
  static const bool verbose = std::getenv("LRR_REF_VERBOSE") != nullptr;
  return verbose ? std::cout : null_stream;
}

// shuffle policy: 0 = keep order (identity; the caller pre-permutes), 1 = Fisher-Yates with SplitMix64(seed),
// state advancing across calls.  Only containers of `moving_` features are policy-shuffled; the index-vector
// shuffles in estimate_LS are dead code in the reference (their result is never read) and follow the same hook.
struct shuffle_state { int mode = 0; unsigned long long s = 0; };
inline shuffle_state& shuffle_hook() { static shuffle_state st; return st; }
inline unsigned long long splitmix64(unsigned long long& s) {
  unsigned long long z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <class It> inline void random_shuffle(It b, It e) {
  shuffle_state& st = shuffle_hook();
  if (st.mode == 0) return;
  auto n = e - b;
  for (auto i = n - 1; i > 0; --i) {
    unsigned long long j = splitmix64(st.s) % (unsigned long long)(i + 1);
    std::iter_swap(b + i, b + j);
  }
}

}  // namespace vxl_shim

#define VCL_DEFINE_SPECIALIZATION template <>
#define vcl_vector std::vector
#define vcl_set std::set
#define vcl_map std::map
#define vcl_list std::list
#define vcl_pair std::pair
#define vcl_make_pair std::make_pair
#define vcl_string std::string
#define vcl_ostream std::ostream
#define vcl_istream std::istream
#define vcl_ofstream std::ofstream
#define vcl_ifstream std::ifstream
#define vcl_ostringstream std::ostringstream
#define vcl_istringstream std::istringstream
#define vcl_stringstream std::stringstream
#define vcl_ostream_iterator std::ostream_iterator
#define vcl_back_inserter std::back_inserter
#define vcl_cout (::vxl_shim::out())
#define vcl_cerr (::vxl_shim::out())
#define vcl_endl std::endl
#define vcl_flush std::flush
#define vcl_setw std::setw
#define vcl_setprecision std::setprecision
#define vcl_setfill std::setfill
#define vcl_fixed std::fixed
#define vcl_ios std::ios
#define vcl_ios_fixed std::ios::fixed
#define vcl_sqrt std::sqrt
#define vcl_abs std::abs
#define vcl_fabs std::fabs
#define vcl_acos std::acos
#define vcl_asin std::asin
#define vcl_atan2 std::atan2
#define vcl_cos std::cos
#define vcl_sin std::sin
#define vcl_exp std::exp
#define vcl_log std::log
#define vcl_pow std::pow
#define vcl_floor std::floor
#define vcl_ceil std::ceil
#define vcl_min std::min
#define vcl_max std::max
#define vcl_sort std::sort
#define vcl_copy std::copy
#define vcl_swap std::swap
#define vcl_find std::find
#define vcl_nth_element std::nth_element
#define vcl_random_shuffle ::vxl_shim::random_shuffle
#define vcl_auto_ptr std::unique_ptr
#define vcl_type_info std::type_info
#define vcl_abort std::abort
#define vcl_exit std::exit
#define vcl_rand std::rand
#define vcl_srand std::srand
#define vcl_size_t std::size_t
#define vcl_memcpy std::memcpy

#endif
