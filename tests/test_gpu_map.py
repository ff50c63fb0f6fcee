"""Device-side map building: mc/hinv.m + mc/add_a_feature_covariance_inverse_depth.m
on the GPU against the reference's golden vector and the oracle."""
import numpy as np
import pytest

from oracle import ekf_oracle as O
from tests import helpers as T

pytestmark = pytest.mark.gpu


def test_device_feature_init_reproduces_the_golden_frame(golden):
    import ekf_slam_b200 as pkg
    N = golden["uv_when_initialized"].shape[0]
    bank = pkg.FilterBank(2, N)
    bank.reset_filters()
    for k in range(N):
        uv = golden["uv_when_initialized"][k]
        bank.add_features_inverse_depth(np.stack([uv, uv]))
    x, P, ns = bank.download_state()
    assert (ns == 13 + 6 * N).all()
    cam, filt, feats = T.build_golden_frame(golden)      # oracle: same recipe, then one prediction
    xo, Po = O.initialize_x_and_p()
    for k in range(N):
        xo, Po, _ = O.add_features_inverse_depth(golden["uv_when_initialized"][k], xo, Po, cam, 1.0, 1.0, 1.0)
    for b in range(2):
        np.testing.assert_allclose(x[b, 13:].reshape(N, 6), golden["yi"], rtol=0, atol=1e-15)   # the REFERENCE's yi
        np.testing.assert_allclose(x[b], xo, rtol=0, atol=1e-15)
        assert T.rel_err(P[b], Po) < 1e-13
        assert np.array_equal(P[b], P[b].T)
    # ... and the frame the reference saved: predict, then h / H / S against the stored values
    bank.begin_frame()
    bank.ekf_prediction()
    bank.measure(1)
    d = bank.download_features()
    for k in range(N):
        np.testing.assert_allclose(d["h"][0, k], golden["h"][k], rtol=0, atol=1e-11)
        np.testing.assert_allclose(d["S"][0, k], golden["S"][k], rtol=1e-12, atol=0)
        Href = golden["H"][k]
        np.testing.assert_allclose(d["Hc"][0, k, :, :7], Href[:, :7], rtol=0, atol=1e-12 * np.abs(Href).max())
        np.testing.assert_allclose(d["Hc"][0, k, :, 7:], Href[:, 13 + 6 * k:19 + 6 * k], rtol=0, atol=1e-12 * np.abs(Href).max())
    bank.close()


def test_device_map_equals_host_closed_form():
    import ekf_slam_b200 as pkg
    import ekf_slam_b200.synth as synth
    seq = synth.SynthSequence(B=5, N=9, T=1, seed=12)
    x0, P0, types = seq.initial_state()
    bank = pkg.FilterBank(5, 9)
    bank.reset_filters()
    for k in range(9):
        bank.add_features_inverse_depth(seq.zc[0, :, k], add=np.array([1, 1, 0, 1, 1], dtype=np.uint8))
    x, P, ns = bank.download_state()
    assert list(ns) == [67, 67, 13, 67, 67]
    for b in (0, 1, 3, 4):
        assert T.rel_err(x[b], x0[b]) < 1e-15 and T.rel_err(P[b], P0[b]) < 1e-13
    # a full map refuses one more feature and says so in the status word
    bank.add_features_inverse_depth(seq.zc[0, :, 0])
    st = bank.download_stats()
    assert (st["status"][[0, 1, 3, 4]] & 4).all() and st["status"][2] == 0
    bank.close()
