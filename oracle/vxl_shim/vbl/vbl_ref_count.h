// Stand-in for VXL vbl_ref_count (intrusive reference count base).  Test infrastructure only, see ../vcl_shim_all.h.
#ifndef vbl_ref_count_shim_h_
#define vbl_ref_count_shim_h_
class vbl_ref_count {
 public:
  vbl_ref_count() : ref_count_(0) {}
  vbl_ref_count(const vbl_ref_count&) : ref_count_(0) {}          // copies start unreferenced
  vbl_ref_count& operator=(const vbl_ref_count&) { return *this; }  // and keep their own count
  virtual ~vbl_ref_count() {}
  void ref() { ++ref_count_; }
  void unref() { if (--ref_count_ <= 0) delete this; }
  int get_references() const { return ref_count_; }
  bool is_referenced() const { return ref_count_ > 0; }
 private:
  int ref_count_;
};
#endif
