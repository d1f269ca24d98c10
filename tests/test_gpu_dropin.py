"""The drop-in boundary, proven end to end: oracle/_ref/dropin_replay (built in the container where /root/reference exists, by
oracle/Makefile) replays the driver's call sites — rrl/itkLocationRegistration.cxx:1632-1644, 2303-2313, 2418-2446, 2504,
1828-1845, 1952 — VERBATIM, once against the reference's own rrl_estimation_symmetric_ICP_matching_all<3,12> (compiled unmodified
from its sources) and once against lrr_b200::rrl_estimation_symmetric_ICP_matching_all<3,12> of include/lrr/lrr_dropin.h, on the
same reference feature / transform objects, and compares iteration counts, ranking angles, decision values, transforms and
covariances."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_replay")


@pytest.mark.gpu
def test_driver_call_sites_replayed_against_the_drop_in():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_replay is built only where the reference tree is present (oracle/Makefile)")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


def test_drop_in_header_names_the_reference_signatures():
    """CPU check: the drop-in header declares the reference's constructor / method surface (rrl sym .h:78-83 and the call sites)."""
    text = open(os.path.join(ROOT, "include", "lrr", "lrr_dropin.h")).read()
    for needle in ("vcl_vector<feature_sptr_type> const& moving_inside, vcl_vector<feature_sptr_type> const& fixed_inside", "typedef vbl_smart_ptr<rrl_estimation_symmetric_ICP_matching_all> sptr",
                   "void set_features(vcl_vector<feature_sptr_type> const& moving_inside", "bool one_iteration(void* caller, display_callback_type display_points, unsigned int iteration)",
                   "void estimate_LS_backward(bool estimate_parameters)", "void set_voronoi_maps(", "void compute_transfer_error_covariance()", "bool oscillated()"):
        assert needle in text, needle
