"""ctypes harness over lrr_b200/liblrr_io.so (include/lrr_io.h): the on-disk formats either side of the hot path — VTK XML PolyData
feature sets (cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:150-235), MetaImage label volumes, the driver's results text
(rrl/itkLocationRegistration.cxx:1874-2030) and ITK transform files.  Host only; used by tests and scripts, like capi.py."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class IoError(RuntimeError):
    pass


class ResultsRecord(C.Structure):
    _fields_ = [("A", C.c_double * 9), ("t_centered", C.c_double * 3), ("covariance", C.c_double * 144), ("covarianceJ", C.c_double * 9),
                ("covarianceJ_evalues", C.c_double * 3), ("fixed_keypoint", C.c_double * 3), ("moving_keypoint", C.c_double * 3),
                ("converged", C.c_int32), ("pad", C.c_int32), ("initial_rms", C.c_double), ("final_rms", C.c_double),
                ("query_point", C.c_double * 3), ("query_point_mapped", C.c_double * 3), ("measurements", C.c_double * 8)]


SYMBOLS = ["lrr_io_read_features_vtp", "lrr_io_features_free", "lrr_io_features_count", "lrr_io_features_xyz", "lrr_io_features_P6",
           "lrr_io_features_C6", "lrr_io_features_strength", "lrr_io_features_shape", "lrr_io_features_dirs", "lrr_io_features_shape_counts",
           "lrr_io_write_features_vtp", "lrr_io_read_mhd", "lrr_io_volume_free", "lrr_io_volume_geometry", "lrr_io_volume_data",
           "lrr_io_labels_at", "lrr_io_write_mhd", "lrr_io_results_t_centered", "lrr_io_write_results", "lrr_io_read_results",
           "lrr_io_results_name", "lrr_io_write_transform", "lrr_io_read_transform",
           "lrr_svm_load", "lrr_svm_free", "lrr_svm_nr_class", "lrr_svm_has_probability", "lrr_svm_labels", "lrr_svm_predict", "lrr_svm_verify"]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liblrr_io.so")
        if not os.path.exists(path):
            raise IoError(f"{path} is missing: run `make -C lrr_b200/csrc`")
        L = C.CDLL(path)
        dp, vp, cp = C.POINTER(C.c_double), C.c_void_p, C.c_char_p
        L.lrr_io_read_features_vtp.argtypes = [cp, C.POINTER(vp), cp, C.c_int]
        L.lrr_io_features_free.argtypes = [vp]; L.lrr_io_features_free.restype = None
        L.lrr_io_features_count.argtypes = [vp]; L.lrr_io_features_count.restype = C.c_int64
        for name in ("xyz", "P6", "C6", "strength", "dirs"):
            f = getattr(L, "lrr_io_features_" + name); f.argtypes = [vp]; f.restype = dp
        L.lrr_io_features_shape.argtypes = [vp]; L.lrr_io_features_shape.restype = C.POINTER(C.c_uint8)
        L.lrr_io_features_shape_counts.argtypes = [vp, C.POINTER(C.c_int64)]; L.lrr_io_features_shape_counts.restype = None
        L.lrr_io_write_features_vtp.argtypes = [cp, C.c_int64, dp, dp, dp, dp, C.POINTER(C.c_uint8), dp, C.c_int, C.c_int, cp, C.c_int]
        L.lrr_io_read_mhd.argtypes = [cp, C.POINTER(vp), cp, C.c_int]
        L.lrr_io_volume_free.argtypes = [vp]; L.lrr_io_volume_free.restype = None
        L.lrr_io_volume_geometry.argtypes = [vp, C.POINTER(C.c_int64), dp, dp]; L.lrr_io_volume_geometry.restype = None
        L.lrr_io_volume_data.argtypes = [vp]; L.lrr_io_volume_data.restype = C.POINTER(C.c_int64)
        L.lrr_io_labels_at.argtypes = [vp, C.c_int64, dp, C.c_int32, C.POINTER(C.c_int32)]
        L.lrr_io_write_mhd.argtypes = [cp, C.POINTER(C.c_int64), dp, dp, C.POINTER(C.c_int64), cp, C.c_int, cp, C.c_int]
        L.lrr_io_results_t_centered.argtypes = [dp] * 6; L.lrr_io_results_t_centered.restype = None
        L.lrr_io_write_results.argtypes = [cp, C.POINTER(ResultsRecord), C.c_int, cp, C.c_int]
        L.lrr_io_read_results.argtypes = [cp, C.POINTER(ResultsRecord), cp, C.c_int]
        L.lrr_io_results_name.argtypes = [cp, C.c_int, cp, cp, cp, C.c_uint, C.c_uint, cp]
        L.lrr_io_write_transform.argtypes = [cp, dp, dp, dp, cp, C.c_int]
        L.lrr_io_read_transform.argtypes = [cp, dp, dp, dp, cp, C.c_int]
        L.lrr_svm_load.argtypes = [cp, C.POINTER(vp), cp, C.c_int]
        L.lrr_svm_free.argtypes = [vp]; L.lrr_svm_free.restype = None
        L.lrr_svm_nr_class.argtypes = [vp]; L.lrr_svm_has_probability.argtypes = [vp]
        L.lrr_svm_labels.argtypes = [vp, C.POINTER(C.c_int32)]
        L.lrr_svm_predict.argtypes = [vp, dp, C.c_int, dp, dp, C.c_int, C.c_int, dp, dp]
        L.lrr_svm_verify.argtypes = [vp, dp, C.POINTER(C.c_int), dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _call(fn, *args):
    err = C.create_string_buffer(512)
    if fn(*args, err, len(err)) != 0:
        raise IoError(err.value.decode(errors="replace"))


def read_features(path: str) -> dict:
    """VTK XML PolyData -> the dict Context.upload takes: xyz (3,n), P6 (6,n), C6 (6,n)|None, strength (n), shape (n) u8, dirs (9,n)."""
    L = lib(); h = C.c_void_p()
    _call(L.lrr_io_read_features_vtp, os.fsencode(path), C.byref(h))
    try:
        n = L.lrr_io_features_count(h)
        arr = lambda p, k: np.ctypeslib.as_array(p, shape=(k * n,)).reshape(k, n).copy() if n else np.zeros((k, 0))
        c6 = L.lrr_io_features_C6(h)
        out = {"xyz": arr(L.lrr_io_features_xyz(h), 3), "P6": arr(L.lrr_io_features_P6(h), 6), "C6": arr(c6, 6) if c6 else None,
               "strength": arr(L.lrr_io_features_strength(h), 1).reshape(n),
               "shape": np.ctypeslib.as_array(L.lrr_io_features_shape(h), shape=(n,)).copy() if n else np.zeros(0, np.uint8),
               "dirs": arr(L.lrr_io_features_dirs(h), 9)}
        cnt = (C.c_int64 * 3)(); L.lrr_io_features_shape_counts(h, cnt)
        out["shape_counts"] = tuple(cnt)
        return out
    finally:
        L.lrr_io_features_free(h)


def write_features(path: str, feats: dict, mode: int = 3, as_float32: bool = True):
    n = feats["xyz"].shape[1]
    c = lambda k, s: np.ascontiguousarray(feats[k], dtype=np.float64).reshape(s)
    xyz, P6, st, dirs = c("xyz", (3, n)), c("P6", (6, n)), c("strength", (n,)), c("dirs", (9, n))
    C6 = c("C6", (6, n)) if feats.get("C6") is not None else None
    sh = np.ascontiguousarray(feats["shape"], dtype=np.uint8)
    _call(lib().lrr_io_write_features_vtp, os.fsencode(path), n, _dp(xyz), _dp(P6), _dp(C6), _dp(st), sh.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dirs),
          mode, int(as_float32))


def read_mhd(path: str) -> dict:
    L = lib(); h = C.c_void_p()
    _call(L.lrr_io_read_mhd, os.fsencode(path), C.byref(h))
    try:
        dims = (C.c_int64 * 3)(); sp = np.zeros(3); org = np.zeros(3)
        L.lrr_io_volume_geometry(h, dims, _dp(sp), _dp(org))
        d = tuple(dims)
        data = np.ctypeslib.as_array(L.lrr_io_volume_data(h), shape=(d[0] * d[1] * d[2],)).copy().reshape(d[2], d[1], d[0])
        return {"dims": d, "spacing": sp, "origin": org, "data": data}     # data[z, y, x]
    finally:
        L.lrr_io_volume_free(h)


def labels_at(path: str, xyz: np.ndarray, outside: int = 0) -> np.ndarray:
    """labels of the voxels under the features (the input of FeatureSet.set_labels), rrl/itkLocationRegistration.cxx:1407"""
    L = lib(); h = C.c_void_p()
    _call(L.lrr_io_read_mhd, os.fsencode(path), C.byref(h))
    try:
        x = np.ascontiguousarray(xyz, dtype=np.float64); n = x.shape[1]
        out = np.zeros(n, np.int32)
        if L.lrr_io_labels_at(h, n, _dp(x), outside, out.ctypes.data_as(C.POINTER(C.c_int32))) != 0:
            raise IoError("labels_at failed")
        return out
    finally:
        L.lrr_io_volume_free(h)


def write_mhd(path: str, data_zyx: np.ndarray, spacing=(1, 1, 1), origin=(0, 0, 0), element_type: str = "MET_UINT", compressed: bool = False):
    d = np.ascontiguousarray(data_zyx, dtype=np.int64)
    dims = (C.c_int64 * 3)(d.shape[2], d.shape[1], d.shape[0])
    sp, org = np.asarray(spacing, np.float64), np.asarray(origin, np.float64)
    _call(lib().lrr_io_write_mhd, os.fsencode(path), dims, _dp(sp), _dp(org), d.ctypes.data_as(C.POINTER(C.c_int64)), element_type.encode(), int(compressed))


def t_centered(A_inv, t_inv, c_inv, fixed_keypoint, moving_keypoint) -> np.ndarray:
    a = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (A_inv, t_inv, c_inv, fixed_keypoint, moving_keypoint)]
    out = np.zeros(3)
    lib().lrr_io_results_t_centered(*[_dp(v) for v in a], _dp(out))
    return out


_REC = ("A", "t_centered", "covariance", "covarianceJ", "covarianceJ_evalues", "fixed_keypoint", "moving_keypoint", "query_point",
        "query_point_mapped", "measurements")


def write_results(path: str, rec: dict, precision: int = 0):
    r = ResultsRecord()
    for k in _REC:
        if k in rec:
            v = np.asarray(rec[k], np.float64).ravel()
            getattr(r, k)[:] = list(v)
    r.converged = int(bool(rec.get("converged", 0))); r.initial_rms = float(rec.get("initial_rms", 0)); r.final_rms = float(rec.get("final_rms", 0))
    _call(lib().lrr_io_write_results, os.fsencode(path), C.byref(r), precision)


def read_results(path: str) -> dict:
    r = ResultsRecord()
    _call(lib().lrr_io_read_results, os.fsencode(path), C.byref(r))
    out = {k: np.array(getattr(r, k)) for k in _REC}
    out["A"] = out["A"].reshape(3, 3); out["covariance"] = out["covariance"].reshape(12, 12); out["covarianceJ"] = out["covarianceJ"].reshape(3, 3)
    out.update(converged=bool(r.converged), initial_rms=r.initial_rms, final_rms=r.final_rms)
    return out


def results_name(directory: str, k: int, candidate: int, prefix: str = "", stem: str = "results", ext: str = ".txt") -> str:
    buf = C.create_string_buffer(4096)
    if lib().lrr_io_results_name(buf, len(buf), os.fsencode(directory), prefix.encode(), stem.encode(), k, candidate, ext.encode()) != 0:
        raise IoError("name too long")
    return buf.value.decode()


def write_transform(path: str, A, t, c):
    a = [np.ascontiguousarray(v, dtype=np.float64).ravel() for v in (A, t, c)]
    _call(lib().lrr_io_write_transform, os.fsencode(path), *[_dp(v) for v in a])


def read_transform(path: str):
    A, t, c = np.zeros(9), np.zeros(3), np.zeros(3)
    _call(lib().lrr_io_read_transform, os.fsencode(path), _dp(A), _dp(t), _dp(c))
    return A.reshape(3, 3), t, c


class SvmModel:
    """A LIBSVM model file behind lrr_svm_* (the driver's verification step, rrl/itkLocationRegistration.cxx:2087-2109, 2508-2528)."""

    def __init__(self, path: str):
        L = lib(); h = C.c_void_p(); err = C.create_string_buffer(512)
        if L.lrr_svm_load(path.encode(), C.byref(h), err, 512) != 0:
            raise IoError(err.value.decode())
        self.h = h
        self.nr_class = L.lrr_svm_nr_class(h); self.has_probability = bool(L.lrr_svm_has_probability(h))
        lab = np.zeros(max(self.nr_class, 1), np.int32); L.lrr_svm_labels(h, lab.ctypes.data_as(C.POINTER(C.c_int32))); self.labels = lab[: self.nr_class]

    def predict(self, x, feature_min=None, feature_max=None, index_base: int = 1, probability: bool = True):
        """(label, probabilities or None)"""
        x = np.ascontiguousarray(x, np.float64); lab = np.zeros(1); prob = np.zeros(max(self.nr_class, 1))
        fmin = None if feature_min is None else np.ascontiguousarray(feature_min, np.float64)
        fmax = None if feature_max is None else np.ascontiguousarray(feature_max, np.float64)
        rc = lib().lrr_svm_predict(self.h, _dp(x), len(x), None if fmin is None else _dp(fmin), None if fmax is None else _dp(fmax), index_base,
                                   int(probability), _dp(lab), _dp(prob))
        if rc != 0:
            raise IoError(f"lrr_svm_predict -> {rc}")
        return float(lab[0]), (prob[: self.nr_class].copy() if (probability and self.has_probability) else None)

    def verify(self, measurements8):
        """(verified, probability of the first class, label) for the eight decision measurements of one candidate"""
        m = np.ascontiguousarray(measurements8, np.float64); assert m.size == 8
        v = C.c_int(0); p = np.zeros(1); lab = np.zeros(1)
        rc = lib().lrr_svm_verify(self.h, _dp(m), C.byref(v), _dp(p), _dp(lab))
        if rc != 0:
            raise IoError(f"lrr_svm_verify -> {rc} (2: not a two-class probability model)")
        return bool(v.value), float(p[0]), float(lab[0])

    def close(self):
        if self.h:
            lib().lrr_svm_free(self.h); self.h = None
