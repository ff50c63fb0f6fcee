"""The oracle against the REFERENCE'S OWN EXECUTION.

tests/golden/ref_*.npz hold what the unmodified /root/reference/matlab_code/*.m sources produced
when run through oracle/mref (mini MATLAB interpreter; tests/golden/make_ref_steps.py).  These tests
pin oracle/ekf_oracle.py (numpy) and oracle/ekf_oracle.c for RANSAC selection, update(), the rescue
gate, the hi update and the Cartesian model: flags bit-exact, hypothesis counts exact, x / P to
1e-11 relative.  Where /root/reference is present (authoring container) the interpreter itself is
re-run and must (a) reproduce the reference-held features_information.mat and (b) regenerate the
committed fixtures bit for bit."""
import os
import warnings

import numpy as np
import pytest

from oracle import ekf_oracle as O
from oracle.mref import run_ref as R
from tests import helpers as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_X, TOL_P = 1e-11, 1e-11


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def initial_state(fx):
    if "x0" in fx:
        return fx["x0"], fx["P0"]
    import ekf_slam_b200.synth as synth
    T_, B, N = fx["zc"].shape[0], fx["zc"].shape[1], fx["zc"].shape[2]
    seq = synth.SynthSequence(B=B, N=N, T=T_, seed=int(fx["seed"]), n_u=int(fx["n_u"]))
    x0, P0, _ = seq.initial_state()
    assert np.array_equal(seq.zc[1:T_ + 1], fx["zc"]), "synthetic generator drifted from the fixture"
    assert np.allclose([x0.sum(), np.abs(P0).sum()], fx["x0_sum"], rtol=1e-13)
    return x0, P0


def oracle_vs_fixture(fx, frames=None, filters=None):
    warnings.filterwarnings("ignore")
    cam = O.initialize_cam()
    x0, P0 = initial_state(fx)
    Tn, B, N = fx["zc"].shape[0], fx["zc"].shape[1], fx["zc"].shape[2]
    frames = Tn if frames is None else min(frames, Tn)
    worst_x = 0.0
    first_flip = None
    for b in (range(B) if filters is None else filters):
        filt = T.oracle_filter(x0[b], P0[b])
        feats = T.oracle_features(fx["types"][b])
        for t in range(frames):
            filt, feats, info = T.oracle_step(filt, feats, cam, fx["zc"][t, b], fx["has"][t, b], fx["U"][b, t])
            fo = T.oracle_flags(feats, N)
            if not np.array_equal(fo, fx["flags"][b, t]) and first_flip is None:
                first_flip = (b, t)
            assert first_flip is None, "flags differ from the reference at (filter, frame) %s" % (first_flip,)
            if (fo & T.F_IC).any():
                assert info["iterations"] == fx["nhyp"][b, t], (b, t)
            worst_x = max(worst_x, T.rel_err(filt.x_k_k, fx["x"][b, t]))
        if frames == Tn:
            if "P" in fx:
                assert T.rel_err(filt.p_k_k, fx["P"][b]) < TOL_P
            else:
                from tests.golden.make_ref_steps import proj_matrix
                V = proj_matrix(filt.p_k_k.shape[0])
                assert T.rel_err(np.diag(filt.p_k_k), fx["P_diag"][b]) < TOL_P
                assert T.rel_err(filt.p_k_k @ V, fx["P_proj"][b]) < TOL_P
    assert worst_x < TOL_X, worst_x
    return worst_x


@pytest.mark.parametrize("name", ["ref_outliers_n16_t6", "ref_mixed_n20_t8", "ref_stale_h", "ref_n100_t3"])
def test_oracle_matches_reference_execution(name):
    fx = load(name)
    ran = set(str(s) for s in fx["ref_functions_run"])
    # the fixture really came out of the reference's own files
    for f in ("ransac_hypotheses", "select_random_match", "compute_hypothesis_support_fast", "update",
              "normJac", "ekf_update_li_inliers", "rescue_hi_inliers", "ekf_update_hi_inliers"):
        assert f in ran, f
    oracle_vs_fixture(fx)


def test_mixed_fixture_runs_cartesian_reference_code():
    fx = load("ref_mixed_n20_t8")
    ran = set(str(s) for s in fx["ref_functions_run"])
    assert "hi_cartesian" in ran and "calculate_Hi_cartesian" in ran
    assert (fx["types"] == 2).sum() >= 8
    assert (fx["flags"] & T.F_HI).any() and (fx["flags"] & T.F_LI).any()


def test_stale_h_fixture_is_the_edge_case():
    """predict_camera_measurements.m:14-16: h of a feature that left the image at x_k_k is the one
    from x_k_km1; the reference's rescue gate used it (feature is IC, not li)."""
    fx = load("ref_stale_h")
    i = int(fx["stale_feature"])
    fl = fx["flags"][0, 0, i]
    assert fl & T.F_HAS_H and fl & T.F_IC and not (fl & T.F_LI)
    # the stored h equals the prediction at x_k_km1, and the feature is NOT visible at x_k_k
    cam = O.initialize_cam()
    filt = T.oracle_filter(fx["x0"][0], fx["P0"][0])
    feats = T.oracle_features(fx["types"][0])
    filt, feats = O.ekf_prediction(filt, feats)
    feats = O.predict_camera_measurements(filt.x_k_km1, cam, feats)
    assert np.allclose(feats[i].h, fx["h"][0, 0, i], rtol=0, atol=1e-12)
    xk = fx["x"][0, 0]
    pos = 13 + 6 * i
    assert O.hi_inverse_depth(xk[pos:pos + 6], xk[0:3], O.q2r(xk[3:7]), cam) is None \
        or fl & T.F_HI  # (if the hi update brought it back the li-state test below still holds)


def test_cfg1_200_frames_oracle_vs_reference():
    """BASELINE configs[0]: ~40 features, 200 frames, free-running: flags exact in every frame."""
    fx = load("ref_cfg1_n40_t200")
    assert fx["zc"].shape[0] == 200
    oracle_vs_fixture(fx, filters=[0])


def test_c_oracle_matches_reference_execution():
    from oracle import c_oracle
    c_oracle.load()
    for name in ("ref_outliers_n16_t6", "ref_mixed_n20_t8"):
        fx = load(name)
        Tn, B, N = fx["zc"].shape[0], fx["zc"].shape[1], fx["zc"].shape[2]
        x, P = np.ascontiguousarray(fx["x0"]).copy(), np.ascontiguousarray(fx["P0"]).copy()
        types = np.ascontiguousarray(fx["types"], dtype=np.uint8)
        nfeat = np.full(B, N, dtype=np.int32)
        for t in range(Tn):
            flags, stats = c_oracle.step_batch(x, P, types, nfeat, np.ascontiguousarray(fx["zc"][t]),
                                               np.ascontiguousarray(fx["has"][t]),
                                               np.ascontiguousarray(fx["U"][:, t]))
            assert np.array_equal(flags & 31, fx["flags"][:, t]), (name, t)
            for b in range(B):
                assert T.rel_err(x[b], fx["x"][b, t]) < 1e-10
        for b in range(B):
            assert T.rel_err(P[b], fx["P"][b]) < 1e-9


# ----------------------------------------------------------------------------------------------
# live: the interpreter on the reference sources (authoring container only)
# ----------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not R.reference_available(),
                               reason="REFERENCE ABSENT (/root/reference does not travel to the GPU box)")


@needs_ref
def test_interpreter_reproduces_reference_held_golden_frame():
    """The interpreter executing hinv.m / add_a_feature_covariance_inverse_depth.m / ekf_prediction.m /
    predict_camera_measurements.m / calculate_derivatives.m reproduces features_information.mat."""
    g = np.load(os.path.join(G, "features_information.npz"))
    I = R.make_interp()
    cam = I.call("initialize_cam")
    x, P = I.call("initialize_x_and_p", nargout=2)
    uv = g["uv_when_initialized"]
    for k in range(uv.shape[0]):
        x, P, yi = I.call("add_features_inverse_depth", uv[k].reshape(2, 1), x, P, cam, 1.0, 1.0, 1.0, nargout=3)
        assert np.abs(yi.reshape(-1) - g["yi"][k]).max() < 1e-15
    filt = R.make_filter(I, x, P)
    fi = I.call("ref_features_info", np.ones((1, uv.shape[0])))
    filt, fi = I.call("ekf_prediction", filt, fi, nargout=2)
    fi = I.call("predict_camera_measurements", filt.get("x_k_km1"), cam, fi)
    fi = I.call("calculate_derivatives", filt.get("x_k_km1"), cam, fi)
    for k, e in enumerate(fi.elems):
        S = e["H"] @ filt.get("p_k_km1") @ e["H"].T + e["R"]
        assert np.abs(e["h"].reshape(-1) - g["h"][k]).max() < 1e-12
        assert T.rel_err(e["H"], g["H"][k]) < 1e-14
        assert T.rel_err(S, g["S"][k]) < 1e-14
    for f in ("hinv", "add_a_feature_covariance_inverse_depth", "hi_inverse_depth", "calculate_Hi_inverse_depth"):
        assert I.sources_used[f].startswith(R.REF_DIR)


@needs_ref
@pytest.mark.parametrize("name", ["ref_outliers_n16_t6", "ref_stale_h"])
def test_fixture_regenerates_from_reference(name):
    warnings.filterwarnings("ignore")
    fx = load(name)
    out = R.run_sequence(fx["x0"][0], fx["P0"][0], fx["types"][0], fx["zc"][:, 0], fx["has"][:, 0], fx["U"][0])
    assert np.array_equal(out["flags"], fx["flags"][0])
    assert np.array_equal(out["nhyp"], fx["nhyp"][0])
    assert np.array_equal(out["x"], fx["x"][0])
    assert np.array_equal(out["P"], fx["P"][0])


@needs_ref
def test_only_missing_functions_are_shimmed():
    """Everything on the path except the two functions the reference does not ship is read from
    /root/reference/matlab_code; `rand` comes from the stored uniform stream."""
    warnings.filterwarnings("ignore")
    fx = load("ref_mixed_n20_t8")
    I = R.make_interp()
    R.run_sequence(fx["x0"][0], fx["P0"][0], fx["types"][0], fx["zc"][:2, 0], fx["has"][:2, 0], fx["U"][0, :2], I=I)
    outside = {k: v for k, v in I.sources_used.items() if not v.startswith(R.REF_DIR)}
    assert set(outside) == {"ref_frame", "ref_features_info", "quaternions", "dq3_by_dq1"}, outside
    for f in ("select_random_match", "generate_state_vector_pattern", "set_as_most_supported_hypothesis",
              "hi_cartesian", "calculate_Hi_cartesian", "update_features_info", "ekf_prediction"):
        assert I.sources_used[f].startswith(R.REF_DIR)
