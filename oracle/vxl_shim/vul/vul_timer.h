// Stand-in for VXL vul_timer.  Test infrastructure only.
#ifndef vul_timer_shim_h_
#define vul_timer_shim_h_
#include <chrono>
class vul_timer {
 public:
  vul_timer() { mark(); }
  void mark() { t0_ = std::chrono::steady_clock::now(); }
  long real() const { return (long)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0_).count(); }
  long user() const { return real(); }
  long all() const { return real(); }
 private:
  std::chrono::steady_clock::time_point t0_;
};
#endif
