"""Config-faithful parity runs (BASELINE.json configs at their named shapes, through the C ABI).

  cfg3  N=100, adaptive RANSAC, 64-filter sample of the 4096-filter batch, 20 frames free-running vs the C oracle
  cfg5  N=100 with 40 Cartesian features, 512 fixed hypotheses, rescue + ekf_update_iterated(3) vs the numpy oracle
  (cfg1 — 200 frames, ~40 features — is tests/test_gpu_ref.py::test_gpu_cfg1_200_frames_vs_reference, checked
   against the reference's own execution; cfg2 is tests/test_gpu_parity.py::test_n100_fixed_256; cfg4 is
   test_large_map_n500.)

A tie-break flip (a residual within rounding of a threshold) would show as a flag mismatch; the assertion reports
the first frame at which one occurs."""
import numpy as np
import pytest

from oracle import ekf_oracle as O
from tests import helpers as T

pytestmark = pytest.mark.gpu
TOL = 1e-9
MASK = T.F_HAS_H | T.F_HAS_Z | T.F_IC | T.F_LI | T.F_HI


def test_cfg3_sample_64_filters_20_frames():
    import ekf_slam_b200 as pkg
    import ekf_slam_b200.synth as synth
    from oracle import c_oracle
    B, N, frames = 64, 100, 20
    # filters 1024..1087 of the bench's 4096-filter batch (same seeds as bench.py: seed 2024 + filter index)
    seq = synth.SynthSequence(B=B, N=N, T=frames, seed=2024, b_offset=1024, n_u=64)
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    for k in range(N):
        bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
    xo, Po, types = seq.initial_state()
    xg, Pg, _ = bank.download_state()
    assert T.rel_err(xg, xo) < 1e-12 and T.rel_err(Pg, Po) < 1e-12      # device-built map == closed form
    xo, Po = np.ascontiguousarray(xg), np.ascontiguousarray(Pg)          # identical starting point
    nf = np.full(B, N, dtype=np.int32)
    first_flip = None
    for t in range(1, frames + 1):
        zc, has, u = np.ascontiguousarray(seq.zc[t]), np.ascontiguousarray(seq.has[t]), seq.uniforms(t)
        bank.upload_candidates(zc, has)
        bank.upload_uniforms(u)
        bank.step()
        fl, st = c_oracle.step_batch(xo, Po, types, nf, zc, has, u)
        fg = bank.download_flags()
        sg = bank.download_stats()
        if first_flip is None and not np.array_equal(fg & MASK, fl & MASK):
            bad = np.argwhere((fg & MASK) != (fl & MASK))[0]
            first_flip = (t, int(bad[0]), int(bad[1]))
        assert first_flip is None, "first flag flip at (frame, filter, feature) %s" % (first_flip,)
        assert np.array_equal(sg["ransac_iters"], st[:, 0])
        xg, _, _ = bank.download_state(want_P=False)
        for b in range(B):
            assert T.rel_err(xg[b], xo[b]) < TOL, (t, b)
    _, Pg, _ = bank.download_state()
    for b in range(B):
        assert T.rel_err(Pg[b], Po[b]) < TOL
    bank.close()


def test_cfg5_mixed_n100_fixed512_rescue_iterated():
    import copy
    import ekf_slam_b200 as pkg
    import ekf_slam_b200.synth as synth
    B, N, frames, fixed = 2, 100, 3, 512
    cart = [i for i in range(N) if i % 5 in (1, 3)]          # 40 Cartesian features, interleaved
    assert len(cart) == 40
    seq = synth.SynthSequence(B=B, N=N, T=frames, seed=5150, n_u=fixed)
    x0, P0, types = seq.initial_state()
    n_max = 13 + 6 * N
    xs, Ps, ts = [], [], []
    for b in range(B):
        xb, Pb, tb = synth.convert_to_cartesian(x0[b], P0[b], types[b], cart)
        Pb = 0.5 * (Pb + Pb.T)
        xs.append(np.pad(xb, (0, n_max - len(xb))))
        Ps.append(np.pad(Pb, ((0, n_max - len(xb)), (0, n_max - len(xb)))))
        ts.append(tb)
    x0, P0, types = np.stack(xs), np.stack(Ps), np.stack(ts)
    n = 13 + 6 * 60 + 3 * 40
    cam = O.initialize_cam()
    bank = pkg.FilterBank(B, N, n_max)
    bank.set_params(fixed_hyp=fixed)
    bank.upload_feature_types(types)
    bank.upload_state(x0, P0)
    filts = [T.oracle_filter(x0[b, :n], P0[b, :n, :n]) for b in range(B)]
    feats = [T.oracle_features(types[b]) for b in range(B)]
    n_hi = 0
    for t in range(1, frames + 1):
        zc, has = seq.frame(t)
        u = seq.uniforms(t, fixed)
        bank.upload_candidates(zc, has)
        bank.upload_uniforms(u)
        bank.begin_frame()
        bank.ekf_prediction()
        bank.measure(1)
        bank.gate()
        bank.ransac_hypotheses()
        bank.update_iterated(pkg.F_LI, 1, 3)
        bank.rescue_hi_inliers()
        bank.ekf_update_hi_inliers()
        xg, Pg, _ = bank.download_state()
        fg = bank.download_flags()
        st = bank.download_stats()
        for b in range(B):
            f, fi = filts[b], O.update_features_info(feats[b])
            f, fi = O.ekf_prediction(f, fi)
            fi = O.search_IC_matches(f, fi, cam, (zc[b], has[b]))
            info = {}
            fi = O.ransac_hypotheses(f, fi, cam, u[b], fixed_hypotheses=fixed, info=info)
            f.x_k_k, f.p_k_k = O.update_iterated(f.x_k_km1, f.p_k_km1, fi, cam, "low_innovation_inlier", n_iter=3)
            fi = O.rescue_hi_inliers(f, fi, cam)
            f = O.ekf_update_hi_inliers(f, fi)
            filts[b], feats[b] = f, fi
            assert np.array_equal(fg[b] & MASK, T.oracle_flags(fi, N)), "flag mismatch: frame %d filter %d" % (t, b)
            assert st["ransac_iters"][b] == fixed == info["iterations"]
            assert T.rel_err(xg[b, :n], f.x_k_k) < TOL and T.rel_err(Pg[b, :n, :n], f.p_k_k) < TOL, (t, b)
            n_hi += int(st["n_hi"][b])
    assert (types == 2).sum() == 2 * 40
    bank.close()
