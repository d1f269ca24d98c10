// Stand-in for vnl_random (only constructed / used in disabled sampling code of the reference's hot path).  Test infrastructure only.
#ifndef vnl_random_shim_h_
#define vnl_random_shim_h_
class vnl_random {
 public:
  vnl_random() : s_(9667566ull) {}
  explicit vnl_random(unsigned long seed) : s_(seed) {}
  void reseed(unsigned long seed) { s_ = seed; }
  unsigned long lrand32() { s_ = s_ * 6364136223846793005ull + 1442695040888963407ull; return (unsigned long)(s_ >> 32); }
  int lrand32(int lo, int hi) { return lo + (int)(lrand32() % (unsigned long)(hi - lo + 1)); }
  int lrand32(int hi) { return lrand32(0, hi); }
  double drand32() { return (double)lrand32() / 4294967296.0; }
  double drand32(double lo, double hi) { return lo + (hi - lo) * drand32(); }
  double drand64() { return drand32(); }
  double drand64(double lo, double hi) { return drand32(lo, hi); }
  double normal() { double u = 0; for (int i = 0; i < 12; ++i) u += drand32(); return u - 6.0; }
 private:
  unsigned long long s_;
};
#endif
