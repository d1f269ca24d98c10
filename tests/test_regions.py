"""FeaturesInRegions (rrl/itkLocationRegistration.cxx:1363-1522): the device selection against the oracle's literal restatement, and
the restatement against an independent numpy statement of the same rules (CPU)."""
import numpy as np
import pytest

from lrr_b200 import synth


def make_case(seed=8, n=60000, size=160.0, n_regions=900):
    pair = synth.make_pair(seed, n, n, size=size)
    rng = np.random.default_rng(seed)
    seeds = rng.uniform(0, size, size=(n_regions, 3))
    xyz = pair["fixed"]["xyz"].T
    # a stand-in for the watershed over-segmentation: label = nearest of a set of random seeds (Voronoi regions)
    from scipy.spatial import cKDTree
    labels = cKDTree(seeds).query(xyz)[1].astype(np.int32) + 1
    return pair, labels


def numpy_in_regions(xyz, labels, x0, x1, ref, min_count):
    inside = np.nonzero(np.all((xyz >= x0) & (xyz <= x1), axis=1))[0]
    d = xyz[inside] - ref
    d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
    order = inside[np.lexsort((inside, d2))]
    lab = labels[order]
    _, first = np.unique(lab, return_index=True)
    out = []
    for f in np.sort(first):
        if len(out) >= min_count:
            break
        out.extend(order[lab == lab[f]].tolist())
    return np.array(out, dtype=np.int32)


CUBES = [((80.0, 75.0, 90.0), 25.0, 1500), ((40.0, 120.0, 60.0), 25.0, 1500), ((80.0, 80.0, 80.0), 12.0, 1500), ((100.0, 60.0, 70.0), 25.0, 400),
         ((-50.0, -50.0, -50.0), 10.0, 1500)]


def test_oracle_restatement_against_numpy(oracle_lib):
    pair, labels = make_case()
    fs = oracle_lib.context().upload(pair["fixed"]); fs.set_labels(labels)
    xyz = pair["fixed"]["xyz"].T
    for c, half, mc in CUBES:
        c = np.array(c); x0, x1 = c - half, c + half
        ref = c + np.array([1.3, -0.7, 2.1])
        got = fs.in_regions(x0, x1, ref, mc)
        exp = numpy_in_regions(xyz, labels, x0, x1, ref, mc)
        assert np.array_equal(got, exp), (c, len(got), len(exp))


@pytest.mark.gpu
def test_device_selection_matches_the_oracle(gpu_lib, oracle_lib):
    pair, labels = make_case()
    g = gpu_lib.context().upload(pair["fixed"]); g.set_labels(labels)
    o = oracle_lib.context().upload(pair["fixed"]); o.set_labels(labels)
    for c, half, mc in CUBES:
        c = np.array(c); x0, x1 = c - half, c + half
        ref = c + np.array([1.3, -0.7, 2.1])
        a, b = g.in_regions(x0, x1, ref, mc), o.in_regions(x0, x1, ref, mc)
        assert np.array_equal(a, b), (c, len(a), len(b))
    # and the selection feeds the estimator like any other inside set (set_features, :1636)
    c = np.array(CUBES[0][0]); Ag, tg, cg = pair["gt"]
    fi = g.in_regions(c - 25.0, c + 25.0, c, 1500)
    assert len(fi) >= 1500 and len(set(fi.tolist())) == len(fi)
