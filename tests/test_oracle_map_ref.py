"""Map management pinned by the reference's own execution: tests/golden/ref_map_t80.npz is
the matlab_code/mono_slam.m:50-82 loop run through oracle/mref with the reference's map_management.m,
update_features_info.m, inversedepth_2_cartesian.m, inversedepth2cartesian.m, delete_a_feature.m,
add_features_inverse_depth.m, hinv.m, add_a_feature_covariance_inverse_depth.m and
add_feature_to_info_vector.m (80 frames: the map grows from empty, 25+ features are converted to Cartesian,
flaky features are deleted).  Only `delete_features` (missing upstream) and the toolbox-bound corner search of
`initialize_features` are harness shims (baseline/octave/shims/)."""
import os
import warnings

import numpy as np
import pytest

from oracle.mref import run_ref as R
from tests import helpers as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    return np.load(os.path.join(G, "ref_map_t80.npz"))


def unpack(fx, b):
    Tn = fx["x"].shape[1]
    return dict(det_uv=fx["det_uv"][b], det_tag=fx["det_tag"][b], det_n=fx["det_n"][b],
                zc=[fx["zc"][b, t, :fx["nfeat"][b, t]] for t in range(Tn)],
                has=[fx["has"][b, t, :fx["nfeat"][b, t]] for t in range(Tn)], U=fx["U"][b])


def test_fixture_exercises_every_map_operation():
    fx = load()
    ran = set(str(s) for s in fx["ref_functions_run"])
    for f in ("map_management", "update_features_info", "inversedepth_2_cartesian", "inversedepth2cartesian",
              "delete_a_feature", "add_features_inverse_depth", "hinv", "add_a_feature_covariance_inverse_depth",
              "add_feature_to_info_vector", "hi_cartesian", "calculate_Hi_cartesian"):
        assert f in ran, f
    for b in range(fx["x"].shape[0]):
        assert (fx["types"][b, -1] == 2).sum() >= 10                     # conversions happened
        tags = [set(fx["tags"][b, t, :fx["nfeat"][b, t]]) for t in range(fx["x"].shape[1])]
        assert any(tags[t - 1] - tags[t] for t in range(1, len(tags)))  # deletions happened
        assert any(tags[t] - tags[t - 1] for t in range(1, len(tags)))  # additions after the first frame
        assert fx["nfeat"][b, 0] == 20 and fx["n_after_mm"][b, 0] == 13 + 6 * 20


def test_oracle_map_management_matches_reference_execution():
    warnings.filterwarnings("ignore")
    fx = load()
    B, Tn = fx["x"].shape[0], fx["x"].shape[1]
    for b in range(B):
        inp = unpack(fx, b)
        o = T.oracle_closed_loop(inp["det_uv"], inp["det_tag"], inp["det_n"], inp["zc"], inp["has"], inp["U"], Tn,
                                 int(fx["min_features"]))
        for t in range(Tn):
            nf, n = fx["nfeat"][b, t], fx["nstate"][b, t]
            assert len(o["types"][t]) == nf and o["n_after_mm"][t] == fx["n_after_mm"][b, t], (b, t)
            assert np.array_equal(o["types"][t], fx["types"][b, t, :nf]), (b, t)
            assert np.array_equal(o["tags"][t], fx["tags"][b, t, :nf]), (b, t)
            assert np.array_equal(o["flags"][t], fx["flags"][b, t, :nf]), (b, t)
            assert np.array_equal(o["counters"][t], fx["counters"][b, t, :nf]), (b, t)
            assert o["nhyp"][t] == fx["nhyp"][b, t]
            assert T.rel_err(o["x"][t], fx["x"][b, t, :n]) < 1e-10, (b, t)
        n = fx["nstate"][b, -1]
        assert T.rel_err(o["P"], fx["P"][b, :n, :n]) < 1e-9


@pytest.mark.skipif(not R.reference_available(), reason="REFERENCE ABSENT (/root/reference does not travel)")
def test_map_fixture_regenerates_from_reference_first_frames():
    """Re-runs the reference's map_management.m loop live for the first 8 frames of filter 0."""
    warnings.filterwarnings("ignore")
    import ekf_slam_b200.synth as synth
    from tests.golden.make_ref_steps import MAP_WORLD
    fx = load()
    world = synth.SynthWorld(2, T=80, seed=int(fx["seed"]), **MAP_WORLD)
    I = R.make_interp()
    r = R.run_closed_loop(world, 0, 8, min_features=int(fx["min_features"]), K=int(fx["K"]), I=I)
    for t in range(8):
        n, nf = fx["nstate"][0, t], fx["nfeat"][0, t]
        assert np.array_equal(r["x"][t], fx["x"][0, t, :n])
        assert np.array_equal(r["flags"][t], fx["flags"][0, t, :nf])
    used = {k: v for k, v in I.sources_used.items() if not v.startswith(R.REF_DIR)}
    assert set(used) == {"ref_frame_nomm", "quaternions", "dq3_by_dq1", "delete_features", "initialize_features"}, used
    assert I.sources_used["map_management"].startswith(R.REF_DIR)
