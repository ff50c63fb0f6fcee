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


def _oracle_twin(x, P, types):
    feats = T.oracle_features(types)
    return np.array(x), np.array(P), feats


def test_device_conversion_and_deletion_match_oracle():
    """mc/inversedepth_2_cartesian.m (linearity-index rule, one conversion per call) and mc/delete_a_feature.m."""
    import ekf_slam_b200 as pkg
    import ekf_slam_b200.synth as synth
    B, N = 4, 14
    seq = synth.SynthSequence(B=B, N=N, T=12, seed=321)
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    for k in range(N):
        bank.add_features_inverse_depth(seq.zc[0, :, k])
    for t in range(1, 13):                     # let the depth uncertainty shrink
        bank.upload_candidates(*seq.frame(t)); bank.upload_uniforms(seq.uniforms(t)); bank.step()
    x, P, ns = bank.download_state()
    types, nf = bank.download_feature_types()
    assert (ns == 13 + 6 * N).all() and (nf == N).all()
    # --- threshold rule vs oracle (a generous threshold so that something converts)
    thr = 0.5
    conv = bank.inversedepth_2_cartesian(threshold=thr)
    x2, P2, ns2 = bank.download_state()
    types2, _ = bank.download_feature_types()
    for b in range(B):
        feats = T.oracle_features(types[b])
        idx = -1
        for i in range(N):
            if O.linearity_index(x[b], P[b], feats, i) < thr:
                idx = i
                break
        assert conv[b] == idx
        if idx >= 0:
            xo, Po = O.convert_feature_to_cartesian(x[b].copy(), P[b].copy(), feats, idx)
            n = len(xo)
            assert ns2[b] == n and types2[b, idx] == 2
            assert T.rel_err(x2[b, :n], xo) < 1e-13 and T.rel_err(P2[b, :n, :n], Po) < 1e-12
            assert np.array_equal(P2[b], P2[b].T) and not P2[b, n:, :].any() and not x2[b, n:].any()
        else:
            assert np.array_equal(x2[b], x[b]) and np.array_equal(P2[b], P[b])
    assert (conv >= 0).any()
    # --- forced conversion of feature 3 everywhere, then deletion of features 0 and 5
    conv = bank.inversedepth_2_cartesian(force_index=3)
    x3, P3, ns3 = bank.download_state()
    types3, _ = bank.download_feature_types()
    delete = np.zeros((B, N), dtype=np.uint8)
    delete[:, 0] = 1
    delete[:, 5] = 1
    bank.delete_features(delete)
    x4, P4, ns4 = bank.download_state()
    types4, nf4 = bank.download_feature_types()
    for b in range(B):
        feats = T.oracle_features(types3[b])
        xo, Po = x3[b, :ns3[b]].copy(), P3[b, :ns3[b], :ns3[b]].copy()
        for i in (5, 0):
            xo, Po = O.delete_a_feature(xo, Po, i, feats)
            feats.pop(i)
        n = len(xo)
        assert ns4[b] == n and nf4[b] == N - 2
        assert np.array_equal(x4[b, :n], xo) and np.array_equal(P4[b, :n, :n], Po)     # pure data movement: bit exact
        assert [f.type for f in feats] == ["inversedepth" if t == 1 else "cartesian" for t in types4[b, :N - 2]]
        assert types4[b, N - 2:].sum() == 0
    # --- and the shrunken, mixed filter still steps in parity with the oracle
    cam = O.initialize_cam()
    keep = [i for i in range(N) if i not in (0, 5)]
    zc, has = seq.frame(12)
    zc2 = np.zeros_like(zc); has2 = np.zeros_like(has)
    zc2[:, :N - 2] = zc[:, keep]; has2[:, :N - 2] = has[:, keep]
    u = seq.uniforms(12)
    bank.upload_candidates(zc2, has2); bank.upload_uniforms(u); bank.step()
    x5, P5, _ = bank.download_state()
    f5 = bank.download_flags()
    for b in range(B):
        n = ns4[b]
        filt = T.oracle_filter(x4[b, :n], P4[b, :n, :n])
        feats = T.oracle_features(types4[b])
        filt, feats, _ = T.oracle_step(filt, feats, cam, zc2[b], has2[b], u[b])
        assert np.array_equal(f5[b] & 31, T.oracle_flags(feats, N))
        assert T.rel_err(x5[b, :n], filt.x_k_k) < 1e-9 and T.rel_err(P5[b, :n, :n], filt.p_k_k) < 1e-9
    bank.close()
