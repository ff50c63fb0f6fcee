"""The reference-named function API (ekf_slam_b200.api) on the GPU against (a) the reference's
own golden vector and (b) the oracle, stage by stage — the test reads like a reference session."""
import copy

import numpy as np
import pytest

from oracle import ekf_oracle as O
from tests.helpers import build_golden_frame, rel_err

pytestmark = pytest.mark.gpu


def _to_api(api, filt, feats):
    f = api.ekf_filter(filt.x_k_k, filt.p_k_k, filt.std_a, filt.std_alpha, filt.std_z, filt.type)
    f.x_k_km1 = None if filt.x_k_km1 is None else filt.x_k_km1.copy()
    f.p_k_km1 = None if filt.p_k_km1 is None else filt.p_k_km1.copy()
    fi = []
    for o in feats:
        a = api.new_feature(o.type, yi=o.yi, uv=o.uv_when_initialized, step=o.init_frame)
        for k in ("h", "z", "H", "S"):
            v = getattr(o, k)
            setattr(a, k, None if v is None else np.array(v))
        a.individually_compatible = o.individually_compatible
        a.low_innovation_inlier = o.low_innovation_inlier
        a.high_innovation_inlier = o.high_innovation_inlier
        fi.append(a)
    return f, fi


def test_golden_frame_through_the_reference_api(golden):
    import ekf_slam_b200.api as api
    cam_o, filt_o, feats_o = build_golden_frame(golden)
    # --- prediction from the augmented state (oracle built the map; both predict)
    filt_o.x_k_km1 = filt_o.p_k_km1 = None
    f, fi = _to_api(api, filt_o, feats_o)
    cam = api.initialize_cam()
    f, fi = api.ekf_prediction(f, fi)
    filt_o, feats_o = O.ekf_prediction(filt_o, feats_o)
    assert rel_err(f.x_k_km1, filt_o.x_k_km1) < 1e-13
    assert rel_err(f.p_k_km1, filt_o.p_k_km1) < 1e-12
    # --- h, H, S against the REFERENCE's stored values
    fi = api.search_IC_matches(f, fi, cam, None)
    for k, a in enumerate(fi):
        np.testing.assert_allclose(a.h, golden["h"][k], rtol=0, atol=1e-11)
        np.testing.assert_allclose(a.H, golden["H"][k], rtol=0, atol=1e-12 * np.abs(golden["H"][k]).max())
        np.testing.assert_allclose(a.S, golden["S"][k], rtol=1e-12, atol=0)
    # --- separate predict / derivative calls give the same
    fi2 = copy.deepcopy(fi)
    for a in fi2:
        a.h = a.H = None
    fi2 = api.predict_camera_measurements(f.x_k_km1, cam, fi2)
    fi2 = api.calculate_derivatives(f.x_k_km1, cam, fi2)
    for a, c in zip(fi, fi2):
        assert np.array_equal(a.h, c.h) and np.array_equal(a.H, c.H)
    # --- the reference's real matches, MATLAB rng(0) uniform stream
    feats_o = O.predict_and_derive(filt_o, feats_o, cam_o)
    for k in range(len(fi)):
        for lst in (fi, feats_o):
            lst[k].z = golden["z"][k].copy()
            lst[k].individually_compatible = int(golden["individually_compatible"][k])
    u = np.random.RandomState(5489).rand(1000)
    info_g, info_o = {}, {}
    fi = api.ransac_hypotheses(f, fi, cam, u=u, info=info_g)
    feats_o = O.ransac_hypotheses(filt_o, feats_o, cam_o, u, info=info_o)
    assert info_g["iterations"] == info_o["iterations"] and info_g["max_support"] == info_o["max_support"]
    assert [a.low_innovation_inlier for a in fi] == [o.low_innovation_inlier for o in feats_o]
    f = api.ekf_update_li_inliers(f, fi)
    filt_o = O.ekf_update_li_inliers(filt_o, feats_o)
    assert rel_err(f.x_k_k, filt_o.x_k_k) < 1e-10 and rel_err(f.p_k_k, filt_o.p_k_k) < 1e-9
    fi = api.rescue_hi_inliers(f, fi, cam)
    feats_o = O.rescue_hi_inliers(filt_o, feats_o, cam_o)
    assert [a.high_innovation_inlier for a in fi] == [o.high_innovation_inlier for o in feats_o]
    for a, o in zip(fi, feats_o):
        assert rel_err(a.h, o.h) < 1e-12 and rel_err(a.H, o.H) < 1e-10
    f = api.ekf_update_hi_inliers(f, fi)
    filt_o = O.ekf_update_hi_inliers(filt_o, feats_o)
    assert rel_err(f.x_k_k, filt_o.x_k_k) < 1e-10 and rel_err(f.p_k_k, filt_o.p_k_k) < 1e-9
    assert abs(np.linalg.norm(f.x_k_k[3:7]) - 1.0) < 1e-15


def test_filter_step_matches_stagewise(golden):
    import ekf_slam_b200.api as api
    import ekf_slam_b200.synth as synth
    seq = synth.SynthSequence(B=1, N=15, T=3, seed=77)
    x0, P0, types = seq.initial_state()
    cam = api.initialize_cam()
    fa = api.ekf_filter(x0[0], P0[0], 0.007, 0.007, 1.0)
    fb = api.ekf_filter(x0[0], P0[0], 0.007, 0.007, 1.0)
    fia = [api.new_feature() for _ in range(15)]
    fib = [api.new_feature() for _ in range(15)]
    for t in range(1, 4):
        zc, has = seq.frame(t)
        u = seq.uniforms(t)[0]
        fia = api.update_features_info(fia)
        fib = api.update_features_info(fib)
        fa, fia = api.filter_step(fa, fia, cam, (zc[0], has[0]), u=u)
        fb, fib = api.ekf_prediction(fb, fib)
        fib = api.search_IC_matches(fb, fib, cam, (zc[0], has[0]))
        fib = api.ransac_hypotheses(fb, fib, cam, u=u)
        fb = api.ekf_update_li_inliers(fb, fib)
        fib = api.rescue_hi_inliers(fb, fib, cam)
        fb = api.ekf_update_hi_inliers(fb, fib)
        assert [a.low_innovation_inlier for a in fia] == [b.low_innovation_inlier for b in fib]
        assert [a.high_innovation_inlier for a in fia] == [b.high_innovation_inlier for b in fib]
        assert rel_err(fa.x_k_k, fb.x_k_k) < 1e-12 and rel_err(fa.p_k_k, fb.p_k_k) < 1e-11


def test_iterated_update_matches_oracle_extension():
    """ekf_update_iterated: the reference names it but ships no implementation; both sides implement the
    standard IEKF (oracle.update_iterated), so this is GPU-vs-oracle consistency, not reference parity."""
    import ekf_slam_b200.api as api
    import ekf_slam_b200.synth as synth
    seq = synth.SynthSequence(B=1, N=18, T=3, seed=88)
    x0, P0, types = seq.initial_state()
    cam = api.initialize_cam()
    cam_o = O.initialize_cam()
    f = api.ekf_filter(x0[0], P0[0], 0.007, 0.007, 1.0)
    fi = [api.new_feature() for _ in range(18)]
    for t in range(1, 3):   # two ordinary frames so the covariance is dense
        fi = api.update_features_info(fi)
        f, fi = api.filter_step(f, fi, cam, (seq.zc[t, 0], seq.has[t, 0]), u=seq.uniforms(t)[0])
    fi = api.update_features_info(fi)
    f, fi = api.ekf_prediction(f, fi)
    fi = api.search_IC_matches(f, fi, cam, (seq.zc[3, 0], seq.has[3, 0]))
    fi = api.ransac_hypotheses(f, fi, cam, u=seq.uniforms(3)[0])
    # oracle twin of the same state
    fo = O.ekf_filter(f.x_k_k, f.p_k_k, 0.007, 0.007, 1.0, "constant_velocity")
    fo.x_k_km1, fo.p_k_km1 = f.x_k_km1.copy(), f.p_k_km1.copy()
    feats_o = []
    for a in fi:
        o = O.Feature(type=a.type, z=None if a.z is None else a.z.copy(), h=None if a.h is None else a.h.copy(),
                      H=None if a.H is None else a.H.copy(), S=None, R=np.eye(2),
                      individually_compatible=a.individually_compatible,
                      low_innovation_inlier=a.low_innovation_inlier, high_innovation_inlier=0)
        feats_o.append(o)
    assert sum(a.low_innovation_inlier for a in fi) >= 4
    for n_iter in (1, 3):
        import copy
        fg = copy.deepcopy(f)
        fg = api.ekf_update_iterated(fg, copy.deepcopy(fi), cam, "low_innovation_inlier", n_iter=n_iter)
        xo, Po = O.update_iterated(fo.x_k_km1, fo.p_k_km1, copy.deepcopy(feats_o), cam_o, "low_innovation_inlier", n_iter=n_iter)
        assert rel_err(fg.x_k_k, xo) < 1e-10 and rel_err(fg.p_k_k, Po) < 1e-9
    # one iteration is the plain update
    fp = api.ekf_update_li_inliers(copy.deepcopy(f), copy.deepcopy(fi))
    f1 = api.ekf_update_iterated(copy.deepcopy(f), copy.deepcopy(fi), cam, "low_innovation_inlier", n_iter=1)
    assert rel_err(f1.x_k_k, fp.x_k_k) < 1e-13 and rel_err(f1.p_k_k, fp.p_k_k) < 1e-13


def test_c_abi_error_behaviour_on_device():
    """Every entry returns a status, never throws across the boundary: bad shapes / ranges / call order give the
    documented negative codes and a message in ekfslam_last_error(); the context stays usable afterwards."""
    import ctypes as C
    import ekf_slam_b200 as pkg
    from ekf_slam_b200 import _lib
    lib = _lib.load()
    bank = pkg.FilterBank(2, 4)
    h = bank._h
    types = np.ones((2, 4), dtype=np.uint8)
    nfeat = np.array([4, 4], dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    # filter range out of bounds
    assert lib.ekfslam_upload_feature_types(h, 1, 2, P(types), P(nfeat)) == -1
    assert b"out of bounds" in lib.ekfslam_last_error()
    # unknown feature type, nfeat beyond N_max
    bad = types.copy(); bad[0, 1] = 7
    assert lib.ekfslam_upload_feature_types(h, 0, 2, P(bad), P(nfeat)) == -1
    assert lib.ekfslam_upload_feature_types(h, 0, 2, P(types), P(np.array([4, 5], dtype=np.int32))) == -1
    # null pointers
    assert lib.ekfslam_upload_candidates(h, 0, 2, None, None) == -1
    assert lib.ekfslam_upload_uniforms(h, 0, 2, None, 8) == -1
    # step before any uniform stream exists: call-order error, not a crash
    assert lib.ekfslam_upload_feature_types(h, 0, 2, P(types), P(nfeat)) == 0
    assert lib.ekfslam_step(h, 1, 1) == -4
    assert b"uniform" in lib.ekfslam_last_error()
    assert lib.ekfslam_step(h, 1, 5) == -1
    # the Python layer raises from the status code
    with pytest.raises(pkg.EkfSlamError) as e:
        bank.upload_feature_types(bad)
    assert e.value.code == -1
    # still usable
    u = np.random.RandomState(0).rand(2, 8)
    bank.upload_uniforms(u)
    bank.upload_candidates(np.zeros((2, 4, 2)), np.zeros((2, 4), dtype=np.uint8))
    bank.step(reset=True, match_mode=1)
    st = bank.download_stats()
    assert (st["status"] == 0).all() and (st["n_li"] == 0).all()
    bank.close()
