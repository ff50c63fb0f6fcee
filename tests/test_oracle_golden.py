"""Pins the oracle against the reference's only golden vector
(matlab_code/features_information.mat -> tests/golden/features_information.npz)."""
import numpy as np

from oracle import ekf_oracle as O
from tests.helpers import build_golden_frame


def test_feature_initialisation_matches_fixture(golden):
    cam, filt, features_info = build_golden_frame(golden)
    assert filt.x_k_km1.shape[0] == 91
    for k, fi in enumerate(features_info):
        np.testing.assert_allclose(fi.yi, golden["yi"][k], rtol=0, atol=1e-15)


def test_h_H_S_match_fixture(golden):
    cam, filt, features_info = build_golden_frame(golden)
    features_info = O.predict_and_derive(filt, features_info, cam)
    for k, fi in enumerate(features_info):
        assert fi.h is not None
        np.testing.assert_allclose(fi.h, golden["h"][k], rtol=0, atol=1e-12)
        Href = golden["H"][k]
        scale = np.abs(Href).max()
        np.testing.assert_allclose(fi.H, Href, rtol=0, atol=1e-13 * scale)
        np.testing.assert_allclose(fi.S, golden["S"][k], rtol=1e-13, atol=0)


def test_fixture_ransac_terminates_early(golden):
    """On the fixture frame (real z from the reference's image matcher) the adaptive
    rule stops after a handful of hypotheses; replay with MATLAB rng(0) == MT19937(5489)."""
    cam, filt, features_info = build_golden_frame(golden)
    features_info = O.predict_and_derive(filt, features_info, cam)
    for k, fi in enumerate(features_info):
        fi.z = golden["z"][k].copy()
        fi.individually_compatible = int(golden["individually_compatible"][k])
    u = np.random.RandomState(5489).rand(1000)
    assert abs(u[0] - 0.8147) < 1e-4 and abs(u[1] - 0.9058) < 1e-4
    info = {}
    features_info = O.ransac_hypotheses(filt, features_info, cam, u, info=info)
    assert 1 <= info["iterations"] <= 10
    assert info["max_support"] >= 10
    filt = O.ekf_update_li_inliers(filt, features_info)
    features_info = O.rescue_hi_inliers(filt, features_info, cam)
    filt = O.ekf_update_hi_inliers(filt, features_info)
    assert np.allclose(filt.p_k_k, filt.p_k_k.T, atol=1e-12)
    assert abs(np.linalg.norm(filt.x_k_k[3:7]) - 1) < 1e-14
    assert np.linalg.eigvalsh(filt.p_k_k).min() > -1e-9
