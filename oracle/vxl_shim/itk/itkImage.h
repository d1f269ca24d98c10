// Stand-in for the few itk::Image members the rrl estimators touch (rrl/rrl_estimation_ICP_matching_all.txx:494-543):
#include <cmath>
// Pointer, GetLargestPossibleRegion().GetSize(), TransformPhysicalPointToIndex, GetPixel.  TEST INFRASTRUCTURE ONLY.
//
// The reference looks the nearest fixed feature up in a precomputed 1 mm Voronoi label image (ITK
// DanielssonDistanceMapImageFilter, built by rrl/examples/compute_distance_map.cxx — third-party, approximate, and
// unable to represent feature 0).  The backend and the oracle return the EXACT nearest neighbour instead
// (north_star item 2).  This stand-in therefore models an *ideal* Voronoi map: an unbounded image whose pixel at a
// physical point is the index of that point's exact nearest feature.  TransformPhysicalPointToIndex keeps the
// continuous point alongside the (never clamped) integer index and GetPixel evaluates the lookup function.
//
// TIMING-ONLY second mode (SetRaster): a real rasterised 1 mm label volume with an O(1) voxel lookup, which is what the stock
// reference runs (rrl/rrl_estimation_ICP_matching_all.txx:530-543; USE_KD_TREE is never defined, Q5) — so that the CPU baseline
// bench.py reports is not slower than the stock path.  Its match sets differ from the exact neighbour where a
// 1 mm voxel straddles a Voronoi face, so it is never used for parity.
#ifndef itkImage_shim_h_
#define itkImage_shim_h_
#include <cstddef>
#include <functional>
#include <memory>
#include <vector>
namespace itk {

template <class T> class SmartPointer {
 public:
  SmartPointer() {}
  SmartPointer(T* p) : p_(p) {}
  SmartPointer(int) {}   // "= 0"
  T* operator->() const { return p_.get(); }
  T& operator*() const { return *p_; }
  T* GetPointer() const { return p_.get(); }
  bool IsNull() const { return !p_; }
  bool IsNotNull() const { return (bool)p_; }
  bool operator!() const { return !p_; }
  explicit operator bool() const { return (bool)p_; }
 private:
  std::shared_ptr<T> p_;
};

template <class TCoord, unsigned N> class Point {
 public:
  TCoord& operator[](unsigned i) { return v_[i]; }
  const TCoord& operator[](unsigned i) const { return v_[i]; }
 private:
  TCoord v_[N] = {};
};

template <unsigned N> struct Size {
  typedef unsigned long SizeValueType;
  SizeValueType m_Size[N];
  SizeValueType& operator[](unsigned i) { return m_Size[i]; }
  const SizeValueType& operator[](unsigned i) const { return m_Size[i]; }
};
template <unsigned N> struct Index {
  typedef long IndexValueType;
  IndexValueType m_Index[N];
  double physical[N];            // stand-in only: the continuous point this index was computed from
  IndexValueType& operator[](unsigned i) { return m_Index[i]; }
  const IndexValueType& operator[](unsigned i) const { return m_Index[i]; }
};
template <unsigned N> struct ImageRegion {
  Size<N> size; Index<N> index;
  const Size<N>& GetSize() const { return size; }
  const Index<N>& GetIndex() const { return index; }
};

template <class TPixel, unsigned N> class Image {
 public:
  typedef Image Self;
  typedef SmartPointer<Self> Pointer;
  typedef TPixel PixelType;
  typedef itk::Index<N> IndexType;
  typedef itk::Size<N> SizeType;
  typedef itk::ImageRegion<N> RegionType;
  typedef typename IndexType::IndexValueType IndexValueType;
  typedef itk::Point<double, N> PointType;
  typedef std::function<TPixel(const double*)> LookupType;
  static Pointer New() { return Pointer(new Self); }
  void SetLookup(LookupType f) { lookup_ = f; }
  // rasterised mode: `dims` voxels of `spacing` mm from `origin` (voxel centres), x fastest; the caller fills Data()
  void SetRaster(const unsigned long* dims, const double* origin, double spacing) {
    raster_ = true; spacing_ = spacing;
    size_t n = 1;
    for (unsigned d = 0; d < N; ++d) { region_.size[d] = dims[d]; origin_[d] = origin[d]; n *= dims[d]; }
    data_.assign(n, TPixel());
  }
  std::vector<TPixel>& Data() { return data_; }
  const RegionType& GetLargestPossibleRegion() const { return region_; }
  template <class TCoord> bool TransformPhysicalPointToIndex(const itk::Point<TCoord, N>& p, IndexType& idx) const {
    if (raster_) {   // itk::Image: index = round((p - origin) / spacing); the caller clamps it to the region (:537-540)
      bool inside = true;
      for (unsigned d = 0; d < N; ++d) {
        idx.physical[d] = p[d]; idx[d] = (IndexValueType)std::floor((p[d] - origin_[d]) / spacing_ + 0.5);
        inside = inside && idx[d] >= 0 && idx[d] < (IndexValueType)region_.size[d];
      }
      return inside;
    }
    for (unsigned d = 0; d < N; ++d) { idx.physical[d] = p[d]; idx[d] = half_ + (IndexValueType)std::llround(p[d]); }
    return true;
  }
  TPixel GetPixel(const IndexType& idx) const {
    if (raster_) { size_t o = 0; for (unsigned d = N; d-- > 0;) o = o * region_.size[d] + (size_t)idx[d]; return data_[o]; }
    return lookup_(idx.physical);
  }
 private:
  Image() { for (unsigned d = 0; d < N; ++d) { region_.size[d] = 2ul * (unsigned long)half_; region_.index[d] = 0; } }
  static const long half_ = 1l << 40;   // unbounded for every practical coordinate: the clamp at :537-540 never triggers
  RegionType region_;
  LookupType lookup_;
  bool raster_ = false; double spacing_ = 1.0; double origin_[N] = {}; std::vector<TPixel> data_;
};
}  // namespace itk
#endif
