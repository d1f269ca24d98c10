// Stand-in for itk::MultiThreader as used by rrl/rrl_estimation_symmetric_ICP_matching_all.txx:660-871 (the reference pins
// it to ONE thread, :760).  SingleMethodExecute runs the callback once per "thread" sequentially.  TEST INFRASTRUCTURE ONLY.
#ifndef itkMultiThreader_shim_h_
#define itkMultiThreader_shim_h_
#include "itkImage.h"
#define ITK_THREAD_RETURN_TYPE void*
#define ITK_THREAD_RETURN_VALUE 0
namespace itk {
class MultiThreader {
 public:
  typedef SmartPointer<MultiThreader> Pointer;
  struct ThreadInfoStruct { int ThreadID; int NumberOfThreads; void* UserData; };
  typedef ITK_THREAD_RETURN_TYPE (*ThreadFunctionType)(void*);
  static Pointer New() { return Pointer(new MultiThreader); }
  void SetNumberOfThreads(int n) { n_ = n < 1 ? 1 : n; }
  int GetNumberOfThreads() const { return n_; }
  void SetSingleMethod(ThreadFunctionType f, void* data) { f_ = f; data_ = data; }
  void SingleMethodExecute() { for (int i = 0; i < n_; ++i) { ThreadInfoStruct info{i, n_, data_}; f_(&info); } }
 private:
  int n_ = 1; ThreadFunctionType f_ = 0; void* data_ = 0;
};
}  // namespace itk
#endif
