"""Map management and the closed loop on the device through the C ABI.

  * ekfslam_map_management + ekfslam_step(reset=0) over the 80-frame fixture tests/golden/ref_map_t80.npz, i.e. against
    what the reference's OWN map_management.m / inversedepth_2_cartesian.m / delete_a_feature.m /
    add_features_inverse_depth.m produced (oracle/mref): map layout, tags, counters and flags exact in every frame,
    x / P within 1e-9;
  * the device-side synthetic world (k_synth_candidates / k_synth_detect / k_synth_uniforms) against its numpy mirror;
  * a 50-frame sequence that runs closed-loop on the GPU with ZERO per-frame host-to-device traffic (detections,
    candidates and RANSAC uniforms are produced on the device), in parity with the oracle driven by the downloaded
    detections / candidates;
  * the reference-named api.map_management(filter, features_info, cam, im, min_n, step)."""
import os

import numpy as np
import pytest

from oracle import ekf_oracle as O
from tests import helpers as T

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9
MASK = T.F_HAS_H | T.F_HAS_Z | T.F_IC | T.F_LI | T.F_HI


@pytest.fixture(scope="module")
def pkg():
    import ekf_slam_b200 as pkg
    return pkg


def test_gpu_map_management_matches_reference_execution(pkg):
    fx = np.load(os.path.join(G, "ref_map_t80.npz"))
    B, Tn, N = fx["x"].shape[0], fx["x"].shape[1], fx["types"].shape[2]
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    minf = int(fx["min_features"])
    for t in range(Tn):
        bank.upload_detections(fx["det_uv"][:, t], fx["det_n"][:, t], tag=fx["det_tag"][:, t])
        bank.map_management(minf)
        _, _, ns = bank.download_state(want_P=False)
        assert np.array_equal(ns, fx["n_after_mm"][:, t]), "state size after map_management, frame %d" % t
        bank.upload_candidates(fx["zc"][:, t], fx["has"][:, t])
        bank.upload_uniforms(np.ascontiguousarray(fx["U"][:, t]))
        bank.step(reset=False, match_mode=1)
        xg, _, ns = bank.download_state(want_P=False)
        ty, nf = bank.download_feature_types()
        tg = bank.download_feature_tags()
        d = bank.download_features()
        st = bank.download_stats()
        for b in range(B):
            k = int(fx["nfeat"][b, t])
            assert nf[b] == k and ns[b] == fx["nstate"][b, t], (t, b)
            assert np.array_equal(ty[b, :k], fx["types"][b, t, :k]), (t, b)
            assert np.array_equal(tg[b, :k], fx["tags"][b, t, :k]), (t, b)
            assert np.array_equal(d["flags"][b, :k] & MASK, fx["flags"][b, t, :k]), "flags, frame %d filter %d" % (t, b)
            assert np.array_equal(d["counters"][b, :k], fx["counters"][b, t, :k]), (t, b)
            if (d["flags"][b, :k] & T.F_IC).any():
                assert st["ransac_iters"][b] == fx["nhyp"][b, t], (t, b)
            assert T.rel_err(xg[b, :ns[b]], fx["x"][b, t, :ns[b]]) < TOL, (t, b)
    _, Pg, ns = bank.download_state()
    for b in range(B):
        assert T.rel_err(Pg[b, :ns[b], :ns[b]], fx["P"][b, :ns[b], :ns[b]]) < TOL
    bank.close()


def _world(B, frames, seed=7):
    import ekf_slam_b200.synth as synth
    from tests.golden.make_ref_steps import MAP_WORLD
    return synth.SynthWorld(B, T=frames, seed=seed, **MAP_WORLD)


def test_device_world_mirrors_numpy(pkg):
    B, N, K = 3, 40, 16
    world = _world(B, 12)
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    bank.world_upload(world)
    tags = np.full((B, N), -1, dtype=np.int32)
    nfeat = np.zeros(B, dtype=np.int32)
    for t in (0, 5, 11):
        bank.world_detect(t, K)
        uv, tg, n = bank.download_detections(K)
        uv_m, tg_m, n_m = world.detections(t, tags, nfeat, K)
        assert np.array_equal(n, n_m) and np.array_equal(tg, tg_m) and np.array_equal(uv, uv_m), t
        bank.map_management(K)                      # grow the maps so that the next frames have features to match
        ty, nfeat = bank.download_feature_types()
        tags = bank.download_feature_tags()
        assert (nfeat > 0).all()
        bank.world_candidates(t + 1)
        zc, has = bank.download_candidates()
        zc_m, has_m = world.candidates(t + 1, tags, nfeat)
        assert np.array_equal(has, has_m), t
        assert np.abs(zc - zc_m).max() < 1e-9, t
        bank.world_uniforms(t + 1, 96)
        bank.step(reset=False, match_mode=1)        # consumes them (and proves u is in place)
        assert (bank.download_stats()["status"] == 0).all()
    # the uniform stream is bit-identical to its numpy mirror
    assert np.array_equal(bank.download_uniforms(), world.uniforms(12, 96))
    bank.close()


def test_closed_loop_on_device_50_frames(pkg):
    """mono_slam.m:50-82 with map growth, conversion and deletion, the whole loop on the GPU: per frame the host only
    issues calls (no host-to-device copies); what it DOWNLOADS (detections, candidates) drives the oracle."""
    B, N, K, frames, minf, n_u = 4, 64, 20, 50, 20, 64
    world = _world(B, frames, seed=2)
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    bank.world_upload(world)
    cam = O.initialize_cam()
    filts, feats = [], []
    for b in range(B):
        x, P = O.initialize_x_and_p()
        filts.append(O.ekf_filter(x, P, 0.007, 0.007, 1.0, "constant_velocity"))
        feats.append([])
    n_conv = n_del = n_add = 0
    for step in range(1, frames + 1):
        bank.world_detect(step - 1, K)              # corners in the previous image            (mono_slam.m:47,53)
        bank.map_management(minf)                   # map_management.m
        bank.world_candidates(step)                 # the new image                            (mono_slam.m:59)
        bank.world_uniforms(step, n_u)
        bank.step(reset=False, match_mode=1)        # mono_slam.m:56-74
        uv, dtag, dn = bank.download_detections(K)
        zc, has = bank.download_candidates()
        U = world.uniforms(step, n_u)
        xg, _, ns = bank.download_state(want_P=False)
        ty, nf = bank.download_feature_types()
        tg = bank.download_feature_tags()
        fg = bank.download_flags()
        for b in range(B):
            before = set(f.feature_when_initialized for f in feats[b])
            ncart = sum(f.type == "cartesian" for f in feats[b])
            filts[b], feats[b] = O.map_management(filts[b], feats[b], cam, (uv[b, :dn[b]], dtag[b, :dn[b]]), minf, step)
            after = set(f.feature_when_initialized for f in feats[b])
            n_del += len(before - after)
            n_add += len(after - before)
            n_conv += sum(f.type == "cartesian" for f in feats[b]) - ncart
            k = len(feats[b])
            filts[b], feats[b] = O.filter_step(filts[b], feats[b], cam, (zc[b, :k], has[b, :k]), U[b])
            assert nf[b] == k and ns[b] == len(filts[b].x_k_k), (step, b)
            assert np.array_equal(tg[b, :k], [f.feature_when_initialized for f in feats[b]]), (step, b)
            assert np.array_equal(ty[b, :k], [1 if f.type == "inversedepth" else 2 for f in feats[b]]), (step, b)
            assert np.array_equal(fg[b, :k] & MASK, T.oracle_flags(feats[b], k)), "flags, frame %d filter %d" % (step, b)
            assert T.rel_err(xg[b, :ns[b]], filts[b].x_k_k) < TOL, (step, b)
    _, Pg, ns = bank.download_state()
    for b in range(B):
        assert T.rel_err(Pg[b, :ns[b], :ns[b]], filts[b].p_k_k) < TOL
    assert n_add > B * minf and n_del > 0, (n_add, n_del, n_conv)
    print("closed loop: %d features added, %d deleted, %d converted to Cartesian over %d frames x %d filters"
          % (n_add, n_del, n_conv, frames, B))
    bank.close()


def test_api_map_management_reference_signature(pkg):
    """api.map_management(filter, features_info, cam, im, min_number_of_features_in_image, step) and the single-purpose
    functions, against the oracle, on the first 30 frames of filter 0 of the reference fixture."""
    import copy
    import ekf_slam_b200.api as api
    fx = np.load(os.path.join(G, "ref_map_t80.npz"))
    cam = api.initialize_cam()
    x, P = api.initialize_x_and_p()
    f = api.ekf_filter(x, P, 0.007, 0.007, 1.0)
    fi = []
    b, minf = 0, int(fx["min_features"])
    for t in range(30):
        nd = int(fx["det_n"][b, t])
        f, fi = api.map_management(f, fi, cam, (fx["det_uv"][b, t, :nd], fx["det_tag"][b, t, :nd]), minf, t + 1)
        k = len(fi)
        assert k == fx["nfeat"][b, t] and len(f.x_k_k) == fx["n_after_mm"][b, t], t
        assert [g.feature_when_initialized for g in fi] == list(fx["tags"][b, t, :k])
        assert [1 if g.type == "inversedepth" else 2 for g in fi] == list(fx["types"][b, t, :k])
        f, fi = api.filter_step(f, fi, cam, (fx["zc"][b, t, :k], fx["has"][b, t, :k]), u=fx["U"][b, t])
        n = fx["nstate"][b, t]
        assert T.rel_err(f.x_k_k, fx["x"][b, t, :n]) < TOL, t
    # single-purpose functions vs the oracle on the state reached
    fo = O.ekf_filter(f.x_k_k.copy(), f.p_k_k.copy(), 0.007, 0.007, 1.0, "constant_velocity")
    feats_o = [O.Feature(type=g.type) for g in fi]
    xo, Po = O.delete_a_feature(fo.x_k_k, fo.p_k_k, 3, feats_o)
    xa, Pa = api.delete_a_feature(f.x_k_k, f.p_k_k, 3, fi)
    assert np.array_equal(xa, xo) and np.array_equal(Pa, 0.5 * (Po + Po.T))
    cam_o = O.initialize_cam()
    uvd = np.array([151.0, 97.0])
    Xo, Pn, newf = O.add_features_inverse_depth(uvd, fo.x_k_k, fo.p_k_k, cam_o, 1.0, 1.0, 1.0)
    Xa, Pna, newa = api.add_features_inverse_depth(uvd, f.x_k_k, f.p_k_k, cam, 1.0, 1.0, 1.0)
    assert T.rel_err(Xa, Xo) < 1e-13 and T.rel_err(Pna, Pn) < 1e-12 and T.rel_err(newa, newf) < 1e-13
    f2, fi2 = api.inversedepth_2_cartesian(copy.deepcopy(f), copy.deepcopy(fi))
    fo2, feats_o2 = O.inversedepth_2_cartesian(copy.deepcopy(fo), copy.deepcopy(feats_o))
    assert [g.type for g in fi2] == [g.type for g in feats_o2]
    assert T.rel_err(f2.x_k_k, fo2.x_k_k) < 1e-12 and T.rel_err(f2.p_k_k, fo2.p_k_k) < 1e-10
