"""On-disk formats either side of the hot path (include/lrr_io.h).  The reader is checked against files produced
by an INDEPENDENT encoder written here with numpy / zlib / base64 from the published VTK XML layout (VTK itself is third-party and
absent: parity unpinned), in every data mode vtkXMLPolyDataWriter can produce; the writers by round trip; the direction assembly
against cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:189-224 restated in numpy."""
import base64
import os
import re
import struct
import zlib

import numpy as np
import pytest

from lrr_b200 import io as lio
from lrr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------- independent VTP encoder
def _block(raw: bytes, *, compressed, header, big, blocksize=1024, joined=False):
    h = (">" if big else "<") + ("I" if header == "UInt32" else "Q")
    if not compressed:
        return struct.pack(h, len(raw)), raw
    blocks = [raw[i:i + blocksize] for i in range(0, len(raw), blocksize)]
    comp = [zlib.compress(b) for b in blocks]
    last = len(raw) % blocksize
    head = struct.pack(h, len(blocks)) + struct.pack(h, blocksize) + struct.pack(h, last) + b"".join(struct.pack(h, len(c)) for c in comp)
    return head, b"".join(comp)


def encode_vtp(path, pieces, *, fmt="ascii", compressed=False, header="UInt32", big=False, appended_encoding="base64", joined=False,
               f64=False, shape_type="UInt8", drop=()):
    """pieces: list of dicts points (n,3), errorProjectors (n,9), strengths (n,), shapes (n,), normals (n,3), binormals (n,3)[, covariances (n,9)]"""
    end = ">" if big else "<"
    ft = "Float64" if f64 else "Float32"
    fdt = np.dtype(end + ("f8" if f64 else "f4"))
    app = b""
    out = ['<?xml version="1.0"?>', f'<VTKFile type="PolyData" version="0.1" byte_order="{"BigEndian" if big else "LittleEndian"}" header_type="{header}"'
           + (' compressor="vtkZLibDataCompressor"' if compressed else "") + ">", "<!-- written by tests/test_io.py -->", "  <PolyData>"]

    def data_array(name, a, typ, dt, nc):
        nonlocal app
        a = np.ascontiguousarray(a)
        if fmt == "ascii":
            body = " ".join(repr(float(x)) if typ.startswith("Float") else str(int(x)) for x in a.ravel())
            return f'<DataArray type="{typ}" Name="{name}" NumberOfComponents="{nc}" format="ascii">\n{body}\n</DataArray>'
        raw = a.astype(dt).tobytes()
        head, body = _block(raw, compressed=compressed, header=header, big=big)
        if fmt == "binary":
            txt = base64.b64encode(head + body).decode() if joined else base64.b64encode(head).decode() + base64.b64encode(body).decode()
            return f'<DataArray type="{typ}" Name="{name}" NumberOfComponents="{nc}" format="binary">\n  {txt}\n</DataArray>'
        off = len(app)
        app += (base64.b64encode(head) + base64.b64encode(body)) if appended_encoding == "base64" else head + body
        return f'<DataArray type="{typ}" Name="{name}" NumberOfComponents="{nc}" format="appended" offset="{off}"/>'

    sdt = np.dtype(end + {"UInt8": "u1", "Int32": "i4", "Float32": "f4"}[shape_type])
    for p in pieces:
        n = len(p["points"])
        out.append(f'    <Piece NumberOfPoints="{n}" NumberOfVerts="0" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="0">')
        out.append('      <PointData Scalars="shapes" Tensors="errorProjectors">')
        for name, nc in (("strengths", 1), ("shapes", 1), ("normals", 3), ("binormals", 3), ("covariances", 9), ("errorProjectors", 9)):
            if name in p and name not in drop:
                typ, dt = (shape_type, sdt) if name == "shapes" else (ft, fdt)
                out.append("        " + data_array(name, p[name], typ, dt, nc))
        out.append("      </PointData>\n      <CellData>\n      </CellData>\n      <Points>")
        out.append("        " + data_array("Points", p["points"], ft, fdt, 3))
        out.append("      </Points>")
        # the empty cell arrays a vtkPolyData always carries
        for sec in ("Verts", "Lines", "Strips", "Polys"):
            out.append(f"      <{sec}>")
            for nm in ("connectivity", "offsets"):
                out.append("        " + data_array(nm, np.zeros(0), "Int64" if fmt != "ascii" else "Int32", np.dtype(end + "i8"), 1))
            out.append(f"      </{sec}>")
        out.append("    </Piece>")
    out.append("  </PolyData>")
    text = "\n".join(out).encode()
    if fmt == "appended":
        text += f'\n  <AppendedData encoding="{appended_encoding}">\n   _'.encode() + app + b"\n  </AppendedData>"
    text += b"\n</VTKFile>\n"
    with open(path, "wb") as f:
        f.write(text)


def make_piece(n, seed, f32=True, with_cov=False):
    pair = synth.make_pair(seed, n, 10)
    f = pair["fixed"]
    dirs = f["dirs"].T.reshape(n, 3, 3)
    sh = f["shape"].astype(np.uint8)
    normals = np.where(sh[:, None] == 0, dirs[:, 2], np.where(sh[:, None] == 1, dirs[:, 1], dirs[:, 0]))
    binormals = np.where(sh[:, None] == 0, dirs[:, 1], np.where(sh[:, None] == 1, dirs[:, 0], 0.0))
    full = [0, 1, 2, 1, 3, 4, 2, 4, 5]
    p = {"points": f["xyz"].T, "errorProjectors": f["P6"].T[:, full], "strengths": f["strength"], "shapes": sh, "normals": normals, "binormals": binormals}
    if with_cov:
        p["covariances"] = f["C6"].T[:, full]
    if f32:   # what the file can hold
        for k in p:
            if k != "shapes":
                p[k] = p[k].astype(np.float32).astype(np.float64)
    return p


def expected(pieces):
    """the filter's output restated (cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:168-229)"""
    pts = np.concatenate([p["points"] for p in pieces]); n = len(pts)
    ep = np.concatenate([p["errorProjectors"] for p in pieces])
    st = np.concatenate([p.get("strengths", np.zeros(len(p["points"]))) for p in pieces])
    sh = np.concatenate([p.get("shapes", np.zeros(len(p["points"]))) for p in pieces]).astype(np.uint8)
    no = np.concatenate([p["normals"] for p in pieces]); bi = np.concatenate([p["binormals"] for p in pieces])
    dirs = np.zeros((n, 3, 3))
    c, t, s = sh == 0, sh == 1, sh == 2
    dirs[c, 0] = np.cross(no[c], bi[c]); dirs[c, 1] = bi[c]; dirs[c, 2] = no[c]
    dirs[t, 0] = bi[t]; dirs[t, 1] = no[t]
    dirs[s, 0] = no[s]
    return {"xyz": pts.T, "P6": ep[:, [0, 1, 2, 4, 5, 8]].T, "strength": st, "shape": sh, "dirs": dirs.reshape(n, 9).T}


def check(got, exp, tol=0.0):
    for k in ("xyz", "P6", "strength", "dirs"):
        assert got[k].shape == exp[k].shape, k
        assert np.max(np.abs(got[k] - exp[k]), initial=0.0) <= tol, k
    assert np.array_equal(got["shape"], exp["shape"])


CASES = [dict(fmt="ascii"),
         dict(fmt="binary"), dict(fmt="binary", joined=True), dict(fmt="binary", header="UInt64"),
         dict(fmt="binary", compressed=True), dict(fmt="binary", compressed=True, header="UInt64", big=True),
         dict(fmt="appended", appended_encoding="raw"), dict(fmt="appended", appended_encoding="raw", compressed=True, header="UInt64"),
         dict(fmt="appended", appended_encoding="base64"), dict(fmt="appended", appended_encoding="base64", compressed=True),
         dict(fmt="appended", appended_encoding="raw", big=True), dict(fmt="binary", f64=True), dict(fmt="binary", shape_type="Int32", compressed=True),
         dict(fmt="ascii", shape_type="Float32")]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}={v}" for k, v in c.items()))
def test_reader_against_independent_encoder(tmp_path, case):
    f32 = not case.get("f64", False)
    pieces = [make_piece(1537, 5, f32=f32)]       # 1537 x 9 x 4 bytes: several 1 KiB zlib blocks with a partial last one
    path = str(tmp_path / "features_00.vtk")
    encode_vtp(path, pieces, **case)
    got = lio.read_features(path)
    check(got, expected(pieces), tol=0.0 if case["fmt"] != "ascii" else 1e-300)
    assert got["C6"] is None
    sh = expected(pieces)["shape"]
    assert got["shape_counts"] == (int((sh == 0).sum()), int((sh == 1).sum()), int((sh == 2).sum()))


def test_reader_pieces_covariances_and_defaults(tmp_path):
    pieces = [make_piece(300, 6, with_cov=True), make_piece(0, 7, with_cov=True) if False else make_piece(211, 8, with_cov=True)]
    path = str(tmp_path / "two_pieces.vtp")
    encode_vtp(path, pieces, fmt="appended", appended_encoding="base64", compressed=True)
    got = lio.read_features(path)
    check(got, expected(pieces))
    cov = np.concatenate([p["covariances"] for p in pieces])[:, [0, 1, 2, 4, 5, 8]].T
    assert np.array_equal(got["C6"], cov)
    # strengths / shapes absent: 0 and CORNER, as the filter's `if( strengths )` / `if( shapes )` guards leave them (:181-186)
    one = [make_piece(64, 9)]
    encode_vtp(path, one, fmt="binary", drop=("strengths", "shapes"))
    got = lio.read_features(path)
    assert np.all(got["strength"] == 0) and np.all(got["shape"] == 0)
    e = expected([{**one[0], "shapes": np.zeros(64), "strengths": np.zeros(64)}])
    check(got, e)
    # empty file
    encode_vtp(path, [make_piece(10, 3)], fmt="ascii")
    assert lio.read_features(path)["xyz"].shape == (3, 10)


def test_reader_errors(tmp_path):
    path = str(tmp_path / "bad.vtp")
    with pytest.raises(lio.IoError, match="cannot open"):
        lio.read_features(str(tmp_path / "missing.vtp"))
    encode_vtp(path, [make_piece(20, 3)], fmt="binary", drop=("errorProjectors",))
    with pytest.raises(lio.IoError, match="errorProjectors"):
        lio.read_features(path)
    encode_vtp(path, [make_piece(20, 3)], fmt="binary", compressed=True)
    data = open(path, "rb").read().replace(b"vtkZLibDataCompressor", b"vtkLZ4DataCompressor")
    open(path, "wb").write(data)
    with pytest.raises(lio.IoError, match="compressor"):
        lio.read_features(path)
    open(path, "wb").write(b'<?xml version="1.0"?><VTKFile type="UnstructuredGrid"></VTKFile>')
    with pytest.raises(lio.IoError, match="PolyData"):
        lio.read_features(path)
    p = make_piece(20, 3); p["shapes"] = p["shapes"].copy(); p["shapes"][4] = 7
    encode_vtp(path, [p], fmt="ascii")
    with pytest.raises(lio.IoError, match="shape value 7"):
        lio.read_features(path)
    encode_vtp(path, [make_piece(50, 3)], fmt="appended", appended_encoding="raw")
    data = open(path, "rb").read()
    open(path, "wb").write(data[:len(data) - 700])      # truncated appended section
    with pytest.raises(lio.IoError):
        lio.read_features(path)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("f32", [True, False])
def test_writer_round_trip(tmp_path, mode, f32):
    pair = synth.make_pair(11, 2500, 10)
    f = pair["fixed"]
    path = str(tmp_path / f"w{mode}.vtp")
    lio.write_features(path, f, mode=mode, as_float32=f32)
    got = lio.read_features(path)
    dirs = f["dirs"].copy()
    corner = f["shape"] == 0
    d = dirs.T.reshape(-1, 3, 3)
    d[corner, 0] = np.cross(d[corner, 2], d[corner, 1])      # the tangent is not stored: it comes back as normal x binormal
    exp = {"xyz": f["xyz"], "P6": f["P6"], "strength": f["strength"], "shape": f["shape"], "dirs": d.reshape(-1, 9).T}
    if f32:
        exp = {k: (v.astype(np.float32).astype(np.float64) if k != "shape" else v) for k, v in exp.items()}
        d32 = f["dirs"].astype(np.float32).astype(np.float64).T.reshape(-1, 3, 3)
        d32[corner, 0] = np.cross(d32[corner, 2], d32[corner, 1]); exp["dirs"] = d32.reshape(-1, 9).T
    check(got, exp, tol=0.0)
    assert np.array_equal(got["C6"], f["C6"].astype(np.float32).astype(np.float64) if f32 else f["C6"])
    if mode == 3 and f32:      # what our writer calls the reference extractor's default must also be readable by the independent decoder
        txt = open(path, "rb").read()
        assert b'encoding="base64"' in txt and b"vtkZLibDataCompressor" in txt and b'type="UInt8" Name="shapes"' in txt


def test_mhd_label_volume(tmp_path):
    rng = np.random.default_rng(2)
    vol = rng.integers(0, 70000, size=(9, 11, 13)).astype(np.int64)       # z, y, x
    sp, org = (0.7, 0.8, 1.25), (-10.0, 5.0, 100.0)
    for comp, et in ((False, "MET_UINT"), (True, "MET_UINT"), (False, "MET_INT")):
        path = str(tmp_path / f"voronoi_{int(comp)}_{et}.mhd")
        lio.write_mhd(path, vol, sp, org, element_type=et, compressed=comp)
        v = lio.read_mhd(path)
        assert v["dims"] == (13, 11, 9) and np.allclose(v["spacing"], sp) and np.allclose(v["origin"], org)
        assert np.array_equal(v["data"], vol)
    # a header written by hand, MSB data in a LOCAL section, 16-bit
    small = rng.integers(0, 60000, size=(2, 3, 4)).astype(">u2")
    path = str(tmp_path / "local.mha")
    with open(path, "wb") as f:
        f.write(b"ObjectType = Image\nNDims = 3\nDimSize = 4 3 2\nElementSpacing = 1 1 2\nOffset = 0 0 0\nElementType = MET_USHORT\n"
                b"BinaryDataByteOrderMSB = True\nElementDataFile = LOCAL\n" + small.tobytes())
    assert np.array_equal(lio.read_mhd(path)["data"], small.astype(np.int64))
    # labels under the features: index = floor((x - origin) / spacing + 0.5); outside -> the given value
    path = str(tmp_path / "voronoi_0_MET_UINT.mhd")
    q = np.stack([rng.uniform(org[d] - 2 * sp[d], org[d] + (vol.shape[2 - d] + 1) * sp[d], 4000) for d in range(3)])
    q[:, 0] = [org[0] + 2.5 * sp[0], org[1] + 3.49999 * sp[1], org[2] - 0.5 * sp[2]]      # half-way cases
    got = lio.labels_at(path, q, outside=-1)
    idx = np.floor((q - np.array(org)[:, None]) / np.array(sp)[:, None] + 0.5).astype(int)
    inside = np.all((idx >= 0) & (idx < np.array([13, 11, 9])[:, None]), axis=0)
    exp = np.full(q.shape[1], -1); exp[inside] = vol[idx[2, inside], idx[1, inside], idx[0, inside]]
    assert np.array_equal(got, exp) and inside.sum() > 500 and (~inside).sum() > 100
    assert got[0] == vol[0, 3, 3]


def test_results_text(tmp_path):
    rng = np.random.default_rng(4)
    M = rng.normal(size=(3, 3)); covJ = M @ M.T
    rec = {"A": np.eye(3) + 0.05 * rng.normal(size=(3, 3)), "t_centered": rng.normal(size=3), "covariance": rng.normal(size=(12, 12)) * 1e-4,
           "covarianceJ": covJ, "fixed_keypoint": [101.5, 220.25, 87.0], "moving_keypoint": [99.0, 215.5, 90.125], "converged": True,
           "initial_rms": 812.75, "final_rms": 455.5, "query_point": [100, 221, 88], "query_point_mapped": [98.5, 216.25, 91.0],
           "measurements": [0.31, 0.28, 0.0125, 0.011, 7.5, 6.25, 11.0, 9.5]}
    path = str(tmp_path / "results.txt")
    lio.write_results(path, rec, precision=17)
    back = lio.read_results(path)
    for k in ("A", "t_centered", "covariance", "covarianceJ", "fixed_keypoint", "moving_keypoint", "query_point", "query_point_mapped", "measurements"):
        assert np.array_equal(np.asarray(back[k]).ravel(), np.asarray(rec[k], float).ravel()), k
    assert back["converged"] is True and back["initial_rms"] == 812.75 and back["final_rms"] == 455.5
    assert np.allclose(back["covarianceJ_evalues"], np.linalg.eigvalsh(covJ), rtol=1e-12)
    # the reference's own formatting: default stream precision, same labels in the same order (rrl/itkLocationRegistration.cxx:1942-2024)
    lio.write_results(path, rec)
    text = open(path).read()
    labels = ["Final estimated transform: ", "Final estimated covariance: ", "Final estimated covarianceJ from moving points: ",
              "covarianceJ from points evalues: ", "Fixed keypoint location: ", "Moving keypoint location: ", "Converged: 1", "Initial RMS: 812.75",
              "Final RMS: 455.5", "RMS reduced: 1", "Query point: [100, 221, 88]", "Query point mapped: [98.5, 216.25, 91]",
              "Final alignment error : Maximum trace : Sheet angles (deg): Tube angles (deg)", "Classification:\t0.31\t0.28\t0.0125\t0.011\t7.5\t6.25\t11\t9.5\t"]
    pos = -1
    for lab in labels:
        nxt = text.find(lab, pos + 1)
        assert nxt > pos, lab
        pos = nxt
    assert np.allclose(lio.read_results(path)["A"], rec["A"], rtol=1e-5)
    # a file as the reference writes it WITH the deformation-field sections in front: those are skipped
    ref_like = ("Affine Approximation Squared Error: 0.5 scale: 0.25\n\nAffine transform fitted to deformation: \n1 0 0\n0 1 0\n0 0 1\n\n0 0 0\n\n"
                "CDC Initial Error: 3 scale: 1\nCDC Final Error: 1 scale: 0.5\nCDC Error reduced: 1\n") + text
    open(path, "w").write(ref_like)
    again = lio.read_results(path)
    assert np.allclose(again["A"], rec["A"], rtol=1e-5) and np.allclose(again["measurements"], rec["measurements"])
    assert lio.results_name("/out/locreg", 17, 3) == "/out/locreg//results000017_03.txt"
    assert lio.results_name("d", 5, 12, stem="transform_inverse", ext=".vtk") == "d//transform_inverse000005_12.vtk"
    # t_centered as WriteResults derives it from the inverse transform (:1884-1889)
    Ai, ti, ci, fk, mk = rng.normal(size=(3, 3)), rng.normal(size=3), rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
    assert np.allclose(lio.t_centered(Ai, ti, ci, fk, mk), ti + Ai @ fk - Ai @ ci - mk, rtol=1e-14)


def test_transform_file(tmp_path):
    rng = np.random.default_rng(5)
    A, t, c = np.eye(3) + 0.1 * rng.normal(size=(3, 3)), rng.normal(size=3) * 10, rng.normal(size=3) * 100
    path = str(tmp_path / "transform000001_00.vtk")
    lio.write_transform(path, A, t, c)
    text = open(path).read()
    assert text.startswith("#Insight Transform File V1.0\n") and "Transform: CenteredAffineTransform_double_3_3\n" in text
    par = np.array(re.search(r"^Parameters: (.*)$", text, re.M).group(1).split(), float)
    assert par.shape == (15,) and np.array_equal(par[:9], A.ravel()) and np.array_equal(par[9:12], c) and np.array_equal(par[12:], t - c)   # T = t - C (:1138-1146)
    A2, t2, c2 = lio.read_transform(path)
    assert np.array_equal(A2, A) and np.array_equal(c2, c) and np.allclose(t2, t, rtol=1e-15, atol=1e-13)
    # a plain AffineTransform file (matrix + translation, centre in FixedParameters), older "# Transform 0" comment form
    with open(path, "w") as f:
        f.write("#Insight Transform File V1.0\n# Transform 0\nTransform: AffineTransform_double_3_3\nParameters: " + " ".join(repr(float(v)) for v in np.concatenate([A.ravel(), t - c]))
                + "\nFixedParameters: " + " ".join(repr(float(v)) for v in c) + "\n")
    A3, t3, c3 = lio.read_transform(path)
    assert np.array_equal(A3, A) and np.array_equal(c3, c) and np.allclose(t3, t, rtol=1e-15, atol=1e-13)
    open(path, "w").write("#Insight Transform File V1.0\nTransform: BSplineDeformableTransform_double_3_3\nParameters: 1 2 3\nFixedParameters: 0\n")
    with pytest.raises(lio.IoError, match="not a 3-D"):
        lio.read_transform(path)


def test_io_abi_exports_every_declared_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "lrr_io.h")).read()
    names = sorted(set(re.findall(r"\b(lrr_(?:io|svm)_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20 and set(names) == set(lio.SYMBOLS)
    L = ctypes.CDLL(os.path.join(ROOT, "lrr_b200", "liblrr_io.so"))
    for n in names:
        assert hasattr(L, n), n


@pytest.mark.gpu
def test_features_from_disk_through_the_backend(tmp_path):
    """a feature file written in the reference extractor's layout (Float32, appended base64, zlib) flows through the CUDA path and gives
    the registration the in-memory arrays (rounded to what the file holds) give"""
    import lrr_b200
    pair = synth.make_pair(31, 6000, 5000)
    ctx = lrr_b200.lib().context(0)
    res = []
    for via_disk in (False, True):
        sets = {}
        for side in ("fixed", "moving"):
            f = pair[side]
            if via_disk:
                p = str(tmp_path / f"{side}_00.vtk"); lio.write_features(p, f, mode=3, as_float32=True); g = lio.read_features(p)
            else:
                g = {k: (v.astype(np.float32).astype(np.float64) if k not in ("shape",) and v is not None else v) for k, v in f.items()}
                d = g["dirs"].T.reshape(-1, 3, 3); co = g["shape"] == 0
                d[co, 0] = np.cross(d[co, 2], d[co, 1]); g["dirs"] = d.reshape(-1, 9).T
            g["C6"] = None
            sets[side] = ctx.upload(g)
        est = ctx.estimator("rrl_sym", sets["moving"], sets["fixed"], 12)
        A, t, c = pair["init"]; est.set_transform(A, t, c)
        it, conv = 0, False
        while not conv and it < 70:
            conv = est.one_iteration(it); it += 1
        res.append((it, est.get_transform(0), est.get_matches(0)[1], est.metric("sheet_angles"), est.metric("tube_angles")))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][2], res[1][2])
    # (same bits in, same matches out; the sums of the work-queue passes are added in queue order, which differs from run to run by rounding)
    for k in ("A", "t", "cov"):
        assert np.allclose(res[0][1][k], res[1][1][k], rtol=1e-10, atol=1e-14), k
    assert np.isclose(res[0][3], res[1][3], rtol=1e-10) and np.isclose(res[0][4], res[1][4], rtol=1e-10)
    tr = res[1][1]
    p = str(tmp_path / "transform000000_00.vtk"); lio.write_transform(p, tr["A"], tr["t"], tr["c"])
    A2, t2, c2 = lio.read_transform(p)
    assert np.array_equal(A2, tr["A"]) and np.allclose(t2, tr["t"], rtol=1e-15, atol=1e-12)
