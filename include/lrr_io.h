/* lrr_io.h — on-disk formats either side of the hot path, host only (lrr_b200/liblrr_io.so, no CUDA):
 *   * feature sets: VTK XML PolyData as the reference reads / writes it — points + point-data arrays `errorProjectors` (9 components),
 *     `strengths`, `shapes`, `normals`, `binormals` (reader side cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:150-235 behind
 *     vtkXMLPolyDataReader, rrl/itkLocationRegistration.cxx:809,826; writer side features/vtkITKPointSetToPolyDataFilter.cxx:146-245:
 *     Float32 arrays, UInt8 shapes), and `covariances` (9 components) for the cdcl feature type
 *     (cdcl/io/vtkPolyDataToFeaturesWithShapeFilter.cxx:153).  Reader: ascii, inline binary (base64) and appended (raw or base64) data,
 *     zlib-compressed or not, UInt32 / UInt64 block headers, both byte orders, Float32 / Float64 / integer arrays, several pieces.
 *     Arrays come back as the structure-of-arrays lrr_features_upload takes (include/lrr_gpu.h), directions assembled per shape exactly
 *     as the reference's filter does (:189-224: corner = normal x binormal, binormal, normal; tube = binormal, normal; sheet = normal).
 *   * label volumes: MetaImage header + raw / zlib data (`voronoi.mhd`, the watershed over-segmentation of FeaturesInRegions,
 *     rrl/itkLocationRegistration.cxx:750-800,1407), and the label of every feature's voxel (TransformPhysicalPointToIndex) — the
 *     input of lrr_features_set_labels.
 *   * per-candidate results text (LocationRegistration::WriteResults, rrl/itkLocationRegistration.cxx:1874-2030): same section
 *     labels, same order, for every section the backend produces (the deformation-field sections need the ITK B-spline baseline and
 *     are written only by the reference).
 *   * transforms: ITK transform text file of the driver's CenteredAffineTransform<double,3> (itk::TransformFileWriter, :2001-2018;
 *     parameter order and the cdcl <-> ITK translation convention of GetCDCTransform / UpdateITKTransform, :1071-1150).
 * VTK, ITK and MetaIO are third-party and absent from the reference tree: the formats are restated from their published
 * specifications (parity unpinned; tests/test_io.py checks the reader against files written by an independent numpy / zlib / base64
 * encoder, and the writers by round trip).
 * Plain C ABI; every function returns 0 on success and fills `err` (when given) otherwise. */
#ifndef LRR_IO_H_
#define LRR_IO_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lrr_feature_file lrr_feature_file;

int  lrr_io_read_features_vtp(const char* path, lrr_feature_file** out, char* err, int errlen);
void lrr_io_features_free(lrr_feature_file* f);
int64_t lrr_io_features_count(const lrr_feature_file* f);
/* component-major arrays, valid until lrr_io_features_free: xyz 3n, P6 6n (xx xy xz yy yz zz), C6 6n or NULL (no `covariances`
 * array), strength n, shape n (0 corner, 1 tube, 2 sheet), dirs 9n */
const double*  lrr_io_features_xyz(const lrr_feature_file* f);
const double*  lrr_io_features_P6(const lrr_feature_file* f);
const double*  lrr_io_features_C6(const lrr_feature_file* f);
const double*  lrr_io_features_strength(const lrr_feature_file* f);
const uint8_t* lrr_io_features_shape(const lrr_feature_file* f);
const double*  lrr_io_features_dirs(const lrr_feature_file* f);
/* number of corners / tubes / sheets, the line the reference's filter prints (:231) */
void lrr_io_features_shape_counts(const lrr_feature_file* f, int64_t counts[3]);

/* mode: 0 = ascii, 1 = appended raw, 2 = inline binary (base64) zlib-compressed, 3 = appended base64 zlib-compressed (the
 * vtkXMLPolyDataWriter default the reference's extractor uses).  Float32 arrays + UInt8 shapes like the reference's writer when
 * `as_float32`, else Float64.  C6 may be NULL.  The directions are written back as `normals` / `binormals` (the inverse of the
 * reader's assembly: corner -> normal = d2, binormal = d1; tube -> normal = d1, binormal = d0; sheet -> normal = d0, binormal = 0). */
int  lrr_io_write_features_vtp(const char* path, int64_t n, const double* xyz, const double* P6, const double* C6, const double* strength,
                               const uint8_t* shape, const double* dirs, int mode, int as_float32, char* err, int errlen);

/* ---- MetaImage label volumes --------------------------------------------------------------------------------------------- */
typedef struct lrr_label_volume lrr_label_volume;
int  lrr_io_read_mhd(const char* path, lrr_label_volume** out, char* err, int errlen);
void lrr_io_volume_free(lrr_label_volume* v);
/* dims[3], spacing[3], origin[3]; data as int64 voxels, x fastest */
void lrr_io_volume_geometry(const lrr_label_volume* v, int64_t dims[3], double spacing[3], double origin[3]);
const int64_t* lrr_io_volume_data(const lrr_label_volume* v);
/* labels_out[i] = voxel value at feature i (index = round((x - origin) / spacing), itk::Image::TransformPhysicalPointToIndex with an
 * identity direction; outside the volume: `outside`) — xyz component-major like lrr_features_upload */
int  lrr_io_labels_at(const lrr_label_volume* v, int64_t n, const double* xyz, int32_t outside, int32_t* labels_out);
/* element_type: "MET_UINT" | "MET_USHORT" | "MET_UCHAR" | "MET_INT" | "MET_SHORT"; compressed: zlib.  Data go to <path minus .mhd>.raw / .zraw */
int  lrr_io_write_mhd(const char* path, const int64_t dims[3], const double spacing[3], const double origin[3], const int64_t* data,
                      const char* element_type, int compressed, char* err, int errlen);

/* ---- results text ---------------------------------------------------------------------------------------------------------- */
typedef struct lrr_results_record {
  double A[9], t_centered[3];            /* "Final estimated transform"                       :1942-1943 */
  double covariance[144];                /* "Final estimated covariance"                      :1945-1946 */
  double covarianceJ[9];                 /* "Final estimated covarianceJ from moving points"  :1948-1950 */
  double covarianceJ_evalues[3];         /* ascending (vnl_symmetric_eigensystem order)       :1952-1956; derived on write */
  double fixed_keypoint[3], moving_keypoint[3];
  int32_t converged; int32_t pad;
  double initial_rms, final_rms;
  double query_point[3], query_point_mapped[3];
  double measurements[8];                /* "Classification:" line                            :2021-2024 */
} lrr_results_record;
/* t_centered of the record from the backend's transform (y = A (x - c) + t) and the two keypoints, as WriteResults derives it from
 * the INVERSE transform (:1884-1889): t' + A' f - A' c' - m with (A', t', c') = inverse */
void lrr_io_results_t_centered(const double* A9_inverse, const double* t3_inverse, const double* c3_inverse, const double* fixed_keypoint3,
                               const double* moving_keypoint3, double* t_centered3);
/* precision: significant digits; 0 = the reference's (default ostream formatting, 6) */
int  lrr_io_write_results(const char* path, const lrr_results_record* r, int precision, char* err, int errlen);
/* reads back what lrr_io_write_results wrote (and what the reference wrote, for the sections above; others are skipped) */
int  lrr_io_read_results(const char* path, lrr_results_record* r, char* err, int errlen);
/* file names of the reference: <dir>//<prefix>results%06u_%02u.txt, <dir>//transform%06u_%02u.vtk, <dir>//transform_inverse...  (:1907,2003,2013) */
int  lrr_io_results_name(char* out, int outlen, const char* dir, const char* prefix, const char* stem, unsigned k, unsigned candidate, const char* ext);

/* ---- ITK transform text file, CenteredAffineTransform_double_3_3 -------------------------------------------------------------
 * Parameters = A row-major, centre, translation; the file's translation is T = t - c because ITK maps P' = A (P - C) + (C + T) and
 * cdcl P' = A (P - C) + t (:1102-1104,1138-1146).  The arguments are the backend's (A, t, c). */
int  lrr_io_write_transform(const char* path, const double* A9, const double* t3, const double* c3, char* err, int errlen);
int  lrr_io_read_transform(const char* path, double* A9, double* t3, double* c3, char* err, int errlen);

/* ---- verification of a candidate by the trained classifier (rrl/itkLocationRegistration.cxx:2087-2109, 2508-2528 through
 * libsvm/svm_interface.cpp:18-102) -------------------------------------------------------------------------------------------------
 * The driver loads a LIBSVM model (`training_data_all.model`, not part of the repository: the caller supplies the file), scales the
 * eight decision measurements of a converged, non-oscillating candidate (lrr_decision::measurements, include/lrr_gpu.h) to [-1, 1] and
 * accepts the registration when the probability of the first class exceeds 0.5.  LIBSVM itself is third-party code carried by the
 * reference tree (libsvm/svm.cpp): restated here from its published model format and methods, pinned against that source compiled in
 * oracle/_ref (tests/test_svm.py).  C-SVC / nu-SVC (any number of classes, with or without probability information), one-class and
 * regression models; linear, polynomial, rbf and sigmoid kernels. */
typedef struct lrr_svm_model lrr_svm_model;
int  lrr_svm_load(const char* path, lrr_svm_model** out, char* err, int errlen);   /* svm_load_model (load_model, svm_interface.cpp:79-102) */
void lrr_svm_free(lrr_svm_model* m);
int  lrr_svm_nr_class(const lrr_svm_model* m);
int  lrr_svm_has_probability(const lrr_svm_model* m);
int  lrr_svm_labels(const lrr_svm_model* m, int32_t* labels);
/* predict() (svm_interface.cpp:18-76): x[i] -> -1 + 2 (x[i] - feature_min[i]) / (feature_max[i] - feature_min[i]) (both NULL: no
 * scaling), node indices index_base + i (the driver uses 0, LIBSVM data files 1), then svm_predict_probability when asked and the
 * model supports it (prob[nr_class]), else svm_predict.  *label = the predicted class label (or value, for regression). */
int  lrr_svm_predict(const lrr_svm_model* m, const double* x, int size, const double* feature_min, const double* feature_max, int index_base,
                     int predict_probability, double* label, double* prob);
/* the driver's acceptance test: its scaling ranges (:2108-2109), indices from 0 (:2099), *verified = probability_estimates[0] > 0.5
 * (:2528).  Returns 2 for a model that is not a two-class probability model (the driver exits on those, svm_interface.cpp:82-86). */
int  lrr_svm_verify(const lrr_svm_model* m, const double* measurements8, int* verified, double* probability_good, double* label);

#ifdef __cplusplus
}
#endif
#endif
