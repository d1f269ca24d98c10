// lrr_ref_svm.cpp — compiles the LIBSVM sources the reference tree carries (libsvm/svm.cpp) and the reference's interface to them
// (libsvm/svm_interface.cpp: load_model, predict) UNMODIFIED from /root/reference.  TEST INFRASTRUCTURE ONLY: pins
// lrr_b200/csrc/lrr_svm.cpp (tests/test_svm.py).  Nothing is copied.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <libsvm/svm.cpp>
#include <libsvm/svm_interface.cpp>

// model file -> (label, probabilities) for one measurement vector, driven as rrl/itkLocationRegistration.cxx:2092-2109,2512-2518 does:
// node indices index_base + i (the driver: 0), end marker -1, predict() scales in place.
extern "C" int lrrr_test_svm_predict(const char* model_path, const double* x, int size, const double* feature_min, const double* feature_max,
                                     int index_base, int predict_probability, double* label, double* prob, int* nr_class) {
  svm_model* model = svm_load_model(model_path);
  if (!model) return 1;
  const bool has_prob = svm_check_probability_model(model) != 0;
  std::vector<svm_node> nodes((size_t)size + 1);
  for (int i = 0; i < size; ++i) { nodes[(size_t)i].index = index_base + i; nodes[(size_t)i].value = x[i]; }
  nodes[(size_t)size].index = -1; nodes[(size_t)size].value = 0.0;
  std::vector<double> fmin(feature_min, feature_min + size), fmax(feature_max, feature_max + size);
  if (nr_class) *nr_class = svm_get_nr_class(model);
  svm_node* xp = nodes.data();
  *label = predict(model, xp, fmin.data(), fmax.data(), (unsigned)size, predict_probability && has_prob, prob);
  svm_destroy_model(model);
  return 0;
}
