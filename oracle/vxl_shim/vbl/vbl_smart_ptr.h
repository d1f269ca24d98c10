// Stand-in for VXL vbl_smart_ptr<T> (intrusive smart pointer over vbl_ref_count).  Test infrastructure only.
#ifndef vbl_smart_ptr_shim_h_
#define vbl_smart_ptr_shim_h_
#include <iosfwd>
template <class T>
class vbl_smart_ptr {
 public:
  vbl_smart_ptr() : p_(0) {}
  vbl_smart_ptr(T* p) : p_(p) { if (p_) p_->ref(); }
  vbl_smart_ptr(const vbl_smart_ptr& o) : p_(o.p_) { if (p_) p_->ref(); }
  template <class U> vbl_smart_ptr(const vbl_smart_ptr<U>& o) : p_(o.as_pointer()) { if (p_) p_->ref(); }
  ~vbl_smart_ptr() { if (p_) p_->unref(); }
  vbl_smart_ptr& operator=(const vbl_smart_ptr& o) { return *this = o.p_; }
  vbl_smart_ptr& operator=(T* p) {
    if (p != p_) { T* old = p_; p_ = p; if (p_) p_->ref(); if (old) old->unref(); }
    return *this;
  }
  T* operator->() const { return p_; }
  T& operator*() const { return *p_; }
  T* ptr() const { return p_; }
  T* as_pointer() const { return p_; }
  explicit operator bool() const { return p_ != 0; }
  bool operator!() const { return p_ == 0; }
  bool operator==(const vbl_smart_ptr& o) const { return p_ == o.p_; }
  bool operator!=(const vbl_smart_ptr& o) const { return p_ != o.p_; }
  bool operator==(const T* o) const { return p_ == o; }
  bool operator!=(const T* o) const { return p_ != o; }
  bool operator<(const vbl_smart_ptr& o) const { return p_ < o.p_; }   // std::set< sptr > orders by address
  bool operator>(const vbl_smart_ptr& o) const { return p_ > o.p_; }
 private:
  T* p_;
};
template <class T> inline std::ostream& operator<<(std::ostream& os, const vbl_smart_ptr<T>& p) { return os << (const void*)p.as_pointer(); }
#endif
