"""Host-side workload generator vs the oracle (no GPU)."""
import numpy as np

import ekf_slam_b200.synth as synth
from oracle import ekf_oracle as O


def test_initial_map_matches_sequential_augmentation():
    seq = synth.SynthSequence(B=2, N=7, T=3, seed=11)
    x, P, types = seq.initial_state()
    cam = O.initialize_cam()
    for b in range(2):
        xo, Po = O.initialize_x_and_p()
        for k in range(7):
            xo, Po, _ = O.add_features_inverse_depth(seq.zc[0, b, k], xo, Po, cam, 1.0, 1.0, 1.0)
        np.testing.assert_allclose(x[b], xo, rtol=0, atol=1e-14)
        np.testing.assert_allclose(P[b], Po, rtol=1e-12, atol=1e-18)
        assert np.array_equal(P[b], P[b].T)


def test_projection_matches_oracle():
    seq = synth.SynthSequence(B=1, N=5, T=4, seed=3, p_outlier=0.0, noise_px=0.0)
    cam = O.initialize_cam()
    for t in (1, 4):
        r, q = seq.pose_r[0, t], seq.pose_q[0, t]
        for k in range(5):
            # a Cartesian feature at the true point seen from the true pose
            h = O.hi_cartesian(seq.points[0, k], r, O.q2r(q), cam)
            assert h is not None
            np.testing.assert_allclose(seq.zc[t, 0, k], h, rtol=0, atol=1e-9)


def test_convert_to_cartesian_matches_oracle():
    seq = synth.SynthSequence(B=1, N=4, T=1, seed=5)
    x, P, types = seq.initial_state()
    feats = [O.Feature(type="inversedepth") for _ in range(4)]
    xo, Po = x[0].copy(), P[0].copy()
    xo, Po = O.convert_feature_to_cartesian(xo, Po, feats, 1)
    xo, Po = O.convert_feature_to_cartesian(xo, Po, feats, 3)
    xn, Pn, tn = synth.convert_to_cartesian(x[0], P[0], types[0], [1, 3])
    assert list(tn) == [1, 2, 1, 2]
    np.testing.assert_allclose(xn, xo, rtol=0, atol=1e-13)
    np.testing.assert_allclose(Pn, Po, rtol=1e-11, atol=1e-16)


def test_sequences_are_shardable():
    a = synth.SynthSequence(B=4, N=6, T=2, seed=9)
    b = synth.SynthSequence(B=2, N=6, T=2, seed=9, b_offset=2)
    np.testing.assert_array_equal(a.zc[:, 2:4], b.zc)
    np.testing.assert_array_equal(a.uniforms(1, 8)[2:4], b.uniforms(1, 8))
