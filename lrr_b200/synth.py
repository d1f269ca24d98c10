"""Deterministic synthetic CT-scale feature volumes.

A *scene* is a set of smooth surfaces (spheres -> SHEET features), curved line segments (-> TUBE
features) and corner sites (-> CORNER features) inside a cube.  The fixed and the moving volume sample the
same scene independently (no 1:1 correspondence); the moving volume is then mapped through T_gt^-1, gets
positional noise and a fraction of uniform outliers.  Feature attributes follow the reference's data model
(cdcl/cdcl_feature_with_shape.h): projector P = n n^T (sheet), I - t t^T (tube), I (corner) as built in
features/itkFeatureImageFilter.txx:1050-1070; directions per cdcl/io/vtkPolyDataToFeaturesICPFilter.cxx:191-224
(sheet: normal; tube: binormal, normal; corner: tangent, binormal, normal).
"""
from __future__ import annotations

import numpy as np

CORNER, TUBE, SHEET = 0, 1, 2


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _perp(v, rng):
    a = rng.normal(size=v.shape)
    a -= (a * v).sum(-1, keepdims=True) * v
    return _unit(a)


def rot_xyz(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


class Scene:
    def __init__(self, seed: int, size: float = 300.0, n_spheres: int = 40, n_curves: int = 60, n_corners: int = 4000):
        rng = np.random.default_rng(seed)
        self.size = size
        self.sph_c = rng.uniform(0.1 * size, 0.9 * size, size=(n_spheres, 3))
        self.sph_r = rng.uniform(0.10 * size, 0.35 * size, size=n_spheres)
        self.cur_a = rng.uniform(0, size, size=(n_curves, 3))
        self.cur_b = rng.uniform(0, size, size=(n_curves, 3))
        self.cur_e = _perp(_unit(self.cur_b - self.cur_a), rng)
        self.cur_amp = rng.uniform(0.0, 0.04 * size, size=n_curves)
        self.corners = rng.uniform(0.05 * size, 0.95 * size, size=(n_corners, 3))
        self.corner_frames = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n_corners)])

    def sample(self, n: int, seed: int, frac_sheet=0.55, frac_tube=0.35):
        """n features in scene (fixed-image) coordinates: returns loc(n,3), shape(n), frame(n,3,3) with rows
        (tangent-ish e0, binormal e1, normal e2): sheet normal = e2, tube tangent = e0."""
        rng = np.random.default_rng(seed)
        n_sheet = int(round(n * frac_sheet)); n_tube = int(round(n * frac_tube)); n_corner = n - n_sheet - n_tube
        locs, shapes, frames = [], [], []
        # sheets: rejection-sample sphere points inside the cube
        got = 0
        while got < n_sheet:
            m = int((n_sheet - got) * 2.2) + 64
            s = rng.integers(0, len(self.sph_r), size=m)
            d = _unit(rng.normal(size=(m, 3)))
            p = self.sph_c[s] + self.sph_r[s, None] * d
            ok = np.all((p >= 0) & (p <= self.size), axis=1)
            p, d = p[ok][: n_sheet - got], d[ok][: n_sheet - got]
            e0 = _perp(d, rng); e1 = np.cross(d, e0)
            locs.append(p); shapes.append(np.full(len(p), SHEET, np.uint8)); frames.append(np.stack([e0, e1, d], axis=1))
            got += len(p)
        # tubes
        got = 0
        while got < n_tube:
            m = int((n_tube - got) * 1.5) + 64
            s = rng.integers(0, len(self.cur_amp), size=m)
            u = rng.uniform(0, 1, size=m)
            ab = self.cur_b[s] - self.cur_a[s]
            p = self.cur_a[s] + u[:, None] * ab + (self.cur_amp[s] * np.sin(2 * np.pi * u))[:, None] * self.cur_e[s]
            tan = _unit(ab + (self.cur_amp[s] * 2 * np.pi * np.cos(2 * np.pi * u))[:, None] * self.cur_e[s])
            ok = np.all((p >= 0) & (p <= self.size), axis=1)
            p, tan = p[ok][: n_tube - got], tan[ok][: n_tube - got]
            nrm = _perp(tan, rng); bin_ = np.cross(nrm, tan)     # binormal x normal = tangent
            locs.append(p); shapes.append(np.full(len(p), TUBE, np.uint8)); frames.append(np.stack([tan, bin_, nrm], axis=1))
            got += len(p)
        # corners
        s = rng.integers(0, len(self.corners), size=n_corner)
        locs.append(self.corners[s] + rng.normal(scale=0.05, size=(n_corner, 3)))
        shapes.append(np.full(n_corner, CORNER, np.uint8)); frames.append(self.corner_frames[s])
        loc = np.concatenate(locs); shape = np.concatenate(shapes); frame = np.concatenate(frames)
        perm = rng.permutation(n)
        return loc[perm], shape[perm], frame[perm]


def make_features(loc, shape, frame, seed: int):
    """Assemble the SoA dict consumed by lrr_features_upload."""
    rng = np.random.default_rng(seed)
    n = len(loc)
    e0, e1, e2 = frame[:, 0], frame[:, 1], frame[:, 2]
    I = np.eye(3)[None]
    P = np.where((shape == SHEET)[:, None, None], e2[:, :, None] * e2[:, None, :],
                 np.where((shape == TUBE)[:, None, None], I - e0[:, :, None] * e0[:, None, :], np.broadcast_to(I, (n, 3, 3))))
    # covariance: eigenvalues (ascending) along (normal, binormal, tangent) so that the smallest-eigenvalue
    # eigenvector is the sheet normal and the largest-eigenvalue eigenvector the tube tangent
    s_small = rng.uniform(0.3, 0.6, size=n); s_mid = rng.uniform(0.9, 1.4, size=n); s_big = rng.uniform(1.8, 3.0, size=n)
    Cm = (s_small ** 2)[:, None, None] * e2[:, :, None] * e2[:, None, :] + (s_mid ** 2)[:, None, None] * e1[:, :, None] * e1[:, None, :] \
        + (s_big ** 2)[:, None, None] * e0[:, :, None] * e0[:, None, :]
    Cm = 0.5 * (Cm + Cm.transpose(0, 2, 1)); P = 0.5 * (P + P.transpose(0, 2, 1))
    strength = rng.lognormal(mean=4.6, sigma=0.7, size=n)
    dirs = np.zeros((n, 3, 3))
    sh, tu, co = shape == SHEET, shape == TUBE, shape == CORNER
    dirs[sh, 0] = e2[sh]
    dirs[tu, 0] = e1[tu]; dirs[tu, 1] = e2[tu]
    dirs[co, 0] = e0[co]; dirs[co, 1] = e1[co]; dirs[co, 2] = e2[co]
    iu = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    return {
        "xyz": np.ascontiguousarray(loc.T, dtype=np.float64),
        "P6": np.ascontiguousarray(np.stack([P[:, a, b] for a, b in iu]), dtype=np.float64),
        "C6": np.ascontiguousarray(np.stack([Cm[:, a, b] for a, b in iu]), dtype=np.float64),
        "strength": np.ascontiguousarray(strength, dtype=np.float64),
        "shape": np.ascontiguousarray(shape, dtype=np.uint8),
        "dirs": np.ascontiguousarray(dirs.reshape(n, 9).T, dtype=np.float64),
    }


def random_affine(rng, max_rot_deg, max_scale, max_trans):
    ang = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg, size=3))
    R = rot_xyz(*ang)
    S = np.eye(3) + rng.uniform(-max_scale, max_scale, size=(3, 3))
    t = rng.uniform(-max_trans, max_trans, size=3)
    return R @ S, t


def make_pair(seed: int, n_fixed: int, n_moving: int, size: float = 300.0, noise: float = 0.3, outliers: float = 0.15,
              rigid: bool = False, gt_rot=10.0, gt_scale=0.05, gt_trans=10.0, init_rot=3.0, init_trans=3.0):
    """Returns dict(fixed, moving, gt=(A,t,c), init=(A,t,c)) where T(x) = A (x - c) + t maps moving -> fixed."""
    scene = Scene(seed, size=size)
    rng = np.random.default_rng(seed + 7919)
    c0 = np.full(3, size / 2)
    A_gt, t_gt = random_affine(rng, gt_rot, 0.0 if rigid else gt_scale, gt_trans)
    fl, fs, ff = scene.sample(n_fixed, seed * 31 + 1)
    ml, ms, mf = scene.sample(n_moving, seed * 31 + 2)
    ml = ml + rng.normal(scale=noise, size=ml.shape)
    # moving-image coordinates: x = A_gt^-1 (y - c0 - t_gt) + c0
    Ai = np.linalg.inv(A_gt)
    ml = (ml - c0 - t_gt) @ Ai.T + c0
    # frames: tangent-like vectors map with A^-1, normals with A^T (inverse-transpose of A^-1); re-orthonormalise
    e0 = _unit(mf[:, 0] @ Ai.T); e2 = mf[:, 2] @ A_gt; e2 = _unit(e2 - (e2 * e0).sum(-1, keepdims=True) * e0); e1 = np.cross(e2, e0)
    mf = np.stack([e0, e1, e2], axis=1)
    n_out = int(round(outliers * n_moving))
    if n_out:
        oi = rng.choice(n_moving, size=n_out, replace=False)
        ml[oi] = rng.uniform(0, size, size=(n_out, 3))
        q = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(min(n_out, 512))])
        q = q[rng.integers(0, len(q), size=n_out)]
        q[:, 2] = np.cross(q[:, 0], q[:, 1])
        mf[oi] = q
    dA, dt = random_affine(rng, init_rot, 0.0, init_trans)
    A0 = dA @ A_gt
    if rigid:  # keep the initial transform a rotation
        U, _, Vt = np.linalg.svd(A0); A0 = U @ Vt
    t0 = c0 + t_gt + dt
    return {
        "fixed": make_features(fl, fs, ff, seed * 31 + 3),
        "moving": make_features(ml, ms, mf, seed * 31 + 4),
        "gt": (A_gt, c0 + t_gt, c0.copy()),
        "init": (A0, t0, c0.copy()),
        "size": size,
    }


def make_pair_2d(seed: int, n_fixed: int = 2000, n_moving: int = 2000, noise: float = 0.004, outliers: float = 0.10,
                 gt_scale: float = 1.08, gt_ang_deg: float = 12.0, gt_trans=(0.15, -0.1), init_ang_deg: float = 4.0, init_trans: float = 0.04):
    """BASELINE.json configs[0]: 2-D point sets with normals for cdcl_estimation_ICP<2,4> + cdcl_trans_similarity2d
    (cdcl/examples/estimate.cxx method 4).  Points on an "H"-like set of strokes plus an arc, in the normalised units the
    demo hands to the estimator (cdcl_normalize_data: centred, mean radius ~1).  Each feature carries the 2x2 location
    covariance C = R diag(s_t^2, s_n^2) R^T (tangent spread >> normal spread), from which the estimator takes the normal.
    Returns dict(fixed, moving: dict(xy(2,n), C3(3,n), strength(n)), gt=(s, ang, t), init=(A 3x3, t3))."""
    rng = np.random.default_rng(seed)

    def strokes(n, rng):
        # five straight strokes of an "H" with a slanted bar, and one arc; arclength-uniform sampling
        segs = [((-0.9, -1.2), (-0.9, 1.2)), ((0.9, -1.2), (0.9, 1.2)), ((-0.9, 0.1), (0.9, -0.2)), ((-1.4, 1.2), (-0.4, 1.2)), ((0.4, -1.2), (1.4, -1.2))]
        lens = np.array([np.hypot(b[0] - a[0], b[1] - a[1]) for a, b in segs] + [np.pi * 0.6])
        which = rng.choice(len(lens), size=n, p=lens / lens.sum())
        u = rng.uniform(size=n)
        xy = np.zeros((n, 2)); tan = np.zeros((n, 2))
        for k, (a, b) in enumerate(segs):
            m = which == k
            a = np.array(a); b = np.array(b)
            xy[m] = a + u[m, None] * (b - a); tan[m] = (b - a) / np.linalg.norm(b - a)
        m = which == len(segs)
        th = np.pi * u[m]
        xy[m] = np.stack([0.6 * np.cos(th), 1.2 + 0.6 * np.sin(th) - 0.6], axis=1); tan[m] = np.stack([-np.sin(th), np.cos(th)], axis=1)
        return xy, tan

    def feats(xy, tan, rng):
        n = len(xy)
        st = rng.uniform(0.03, 0.08, size=n); sn = rng.uniform(0.004, 0.012, size=n)
        nx, ny = -tan[:, 1], tan[:, 0]
        cxx = st ** 2 * tan[:, 0] ** 2 + sn ** 2 * nx ** 2
        cxy = st ** 2 * tan[:, 0] * tan[:, 1] + sn ** 2 * nx * ny
        cyy = st ** 2 * tan[:, 1] ** 2 + sn ** 2 * ny ** 2
        return {"xy": np.ascontiguousarray(xy.T), "C3": np.ascontiguousarray(np.stack([cxx, cxy, cyy])),
                "strength": np.exp(rng.normal(0.0, 0.4, size=n))}

    fxy, ftan = strokes(n_fixed, rng)
    mxy, mtan = strokes(n_moving, rng)
    fxy = fxy + rng.normal(scale=noise, size=fxy.shape)
    mxy = mxy + rng.normal(scale=noise, size=mxy.shape)
    ang = np.deg2rad(gt_ang_deg)
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    t = np.asarray(gt_trans, dtype=np.float64)
    # moving coordinates: x = (s R)^-1 (y - t); tangents rotate with R^-1
    mxy = (mxy - t) @ np.linalg.inv(gt_scale * R).T
    mtan = mtan @ R          # (R^-1 v)^T = v^T R
    n_out = int(round(outliers * n_moving))
    if n_out:
        oi = rng.choice(n_moving, size=n_out, replace=False)
        mxy[oi] = rng.uniform(-1.6, 1.6, size=(n_out, 2))
        a = rng.uniform(0, 2 * np.pi, size=n_out); mtan[oi] = np.stack([np.cos(a), np.sin(a)], axis=1)
    a0 = ang + np.deg2rad(init_ang_deg) * (1 if seed % 2 == 0 else -1)
    s0 = gt_scale * 0.97
    A0 = np.eye(3); A0[:2, :2] = s0 * np.array([[np.cos(a0), -np.sin(a0)], [np.sin(a0), np.cos(a0)]])
    t0 = np.array([t[0] + init_trans, t[1] - init_trans, 0.0])
    return {"fixed": feats(fxy, ftan, rng), "moving": feats(mxy, mtan, rng), "gt": (gt_scale, ang, t), "init": (A0, t0)}


def subset(feats: dict, idx) -> dict:
    idx = np.asarray(idx)
    return {k: (np.ascontiguousarray(v[..., idx]) if v is not None else None) for k, v in feats.items()}


def make_jobs(pair: dict, n_locations: int, n_inits: int, seed: int, roi_half: float = 25.0,
              bad_rot=30.0, bad_trans=20.0):
    """C4 job list: n_locations query points x n_inits initialisations (first = near-correct)."""
    from .capi import JOB_DTYPE
    rng = np.random.default_rng(seed)
    size = pair["size"]
    A_gt, t_gt, c_gt = pair["gt"]
    jobs = np.zeros(n_locations * n_inits, dtype=JOB_DTYPE)
    locs_f = rng.uniform(0.25 * size, 0.75 * size, size=(n_locations, 3))
    Ai = np.linalg.inv(A_gt)
    for k in range(n_locations):
        qf = locs_f[k]
        qm = (qf - t_gt) @ Ai.T + c_gt
        for c in range(n_inits):
            if c == 0:
                dA, dt = random_affine(rng, 3.0, 0.0, 3.0)
            else:
                dA, dt = random_affine(rng, bad_rot, 0.0, bad_trans)
            A0 = dA @ A_gt
            # rotate about the query location: T0(x) = A0 (x - qm) + (qf + dt)
            j = jobs[k * n_inits + c]
            j["A"] = A0.reshape(9); j["t"] = qf + dt; j["c"] = qm
            j["roi_center_moving"] = qm; j["roi_center_fixed"] = qf; j["roi_half_size"] = roi_half
    return jobs
