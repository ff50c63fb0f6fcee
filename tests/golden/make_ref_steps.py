"""Generates tests/golden/ref_*.npz by EXECUTING the unmodified reference sources
(/root/reference/matlab_code/*.m) through oracle/mref (mini MATLAB interpreter) over synthetic
filter steps.  Run in the authoring container (the reference does not travel to the GPU box):

    python tests/golden/make_ref_steps.py [case ...]

Each fixture stores the inputs (x0, P0 or the seed that regenerates them, feature types, candidate
pixels, candidate mask, RANSAC uniform streams) and what the reference produced: x_k_k per frame,
p_k_k after the last frame (full, or diagonal + seeded projections for n = 613), the per-feature
flag bytes per frame (HAS_H=1 HAS_Z=2 IC=4 LI=8 HI=16) and the number of hypotheses executed per
frame (= uniforms drawn by the reference's own select_random_match.m).
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.mref import run_ref as R          # noqa: E402
from oracle import ekf_oracle as O            # noqa: E402
import ekf_slam_b200.synth as synth          # noqa: E402
from tests import helpers as T               # noqa: E402

warnings.filterwarnings("ignore")


def proj_matrix(n, k=8, seed=1234):
    return np.random.RandomState(seed).standard_normal((n, k))


def run_case(name, x0, P0, types, zc, has, U, full_P=True, extra=None):
    """x0 [B,n], P0 [B,n,n], types [B,N], zc [T,B,N,2], has [T,B,N], U [B,T,n_u]."""
    B = x0.shape[0]
    I = R.make_interp()
    outs = []
    t0 = time.time()
    for b in range(B):
        outs.append(R.run_sequence(x0[b], P0[b], types[b], zc[:, b], has[:, b], U[b], I=I))
        print("  %s: filter %d/%d done, %.1f s" % (name, b + 1, B, time.time() - t0), flush=True)
    d = dict(types=types, zc=zc, has=has, U=U,
             x=np.array([o["x"] for o in outs]), flags=np.array([o["flags"] for o in outs]),
             nhyp=np.array([o["nhyp"] for o in outs]), h=np.array([o["h"] for o in outs]))
    if full_P:
        d["x0"], d["P0"] = x0, P0
        d["P"] = np.array([o["P"] for o in outs])
    else:
        V = proj_matrix(x0.shape[1])
        d["P_diag"] = np.array([np.diag(o["P"]) for o in outs])
        d["P_proj"] = np.array([o["P"] @ V for o in outs])
        d["P_fro"] = np.array([np.linalg.norm(o["P"]) for o in outs])
        d["x0_sum"] = np.array([x0.sum(), np.abs(P0).sum()])
    if extra:
        d.update(extra)
    d["ref_functions_run"] = np.array(sorted(k for k in I.call_counts if k in I.sources_used and
                                             I.sources_used[k].startswith(R.REF_DIR)))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.0f KB, %.1f s)" % (path, os.path.getsize(path) / 1024, time.time() - t0))


def seq_inputs(B, N, Tn, seed, **kw):
    seq = synth.SynthSequence(B=B, N=N, T=Tn, seed=seed, **kw)
    x0, P0, types = seq.initial_state()
    return seq, x0, P0, types, seq.zc[1:Tn + 1], seq.has[1:Tn + 1], seq.U[:, 1:Tn + 1]


def case_cfg1():
    """BASELINE configs[0]: single filter(s), ~40 features, 200 frames (mono_slam.m:50 loop)."""
    seq, x0, P0, types, zc, has, U = seq_inputs(2, 40, 200, seed=11, n_u=48)
    run_case("ref_cfg1_n40_t200", x0, P0, types, zc, has, U, extra=dict(seed=11, n_u=48))


def case_outliers():
    seq, x0, P0, types, zc, has, U = seq_inputs(3, 16, 6, seed=5, p_outlier=0.5, n_u=128)
    run_case("ref_outliers_n16_t6", x0, P0, types, zc, has, U)


def case_mixed():
    """Mixed inverse-depth / Cartesian map: hi_cartesian.m, calculate_Hi_cartesian.m and the
    z_euc branch of compute_hypothesis_support_fast.m."""
    B, N, Tn = 2, 20, 8
    seq = synth.SynthSequence(B=B, N=N, T=Tn + 4, seed=21, n_u=64)
    x0, P0, types = seq.initial_state()
    cam = O.initialize_cam()
    xs, Ps, ty = [], [], []
    for b in range(B):
        # four frames with the oracle so that depths are observable, then convert 8 features
        filt = T.oracle_filter(x0[b], P0[b])
        feats = T.oracle_features(types[b])
        for t in range(1, 5):
            filt, feats, _ = T.oracle_step(filt, feats, cam, seq.zc[t, b], seq.has[t, b], seq.U[b, t])
        which = list(range(1, N, 5)) + list(range(2, N, 5))
        x, P, tb = synth.convert_to_cartesian(filt.x_k_k, filt.p_k_k, types[b], which)
        xs.append(x)
        Ps.append(0.5 * (P + P.T))
        ty.append(tb)
    run_case("ref_mixed_n20_t8", np.array(xs), np.array(Ps), np.array(ty),
             seq.zc[5:Tn + 5], seq.has[5:Tn + 5], seq.U[:, 5:Tn + 5])


def case_n100():
    seq, x0, P0, types, zc, has, U = seq_inputs(1, 100, 3, seed=0, n_u=64)
    run_case("ref_n100_t3", x0, P0, types, zc, has, U, full_P=False, extra=dict(seed=0, n_u=64))


def find_stale_h():
    """A frame in which a feature is predicted (visible) at x_k_km1, individually compatible, NOT a
    low-innovation inlier, and leaves the image at x_k_k: predict_camera_measurements.m:14-16 then
    keeps its h from x_k_km1 while calculate_Hi_inverse_depth.m:3 re-linearises at x_k_k with that
    stale h, and rescue_hi_inliers.m:11-20 gates it with the mixture.  Found by search with the
    numpy oracle, then executed by the reference."""
    cam = O.initialize_cam()
    N, warm = 14, 5
    for seed in range(400):
        seq = synth.SynthSequence(B=1, N=N, T=warm + 1, seed=1000 + seed, n_u=64, p_outlier=0.1)
        x0, P0, types = seq.initial_state()
        filt = T.oracle_filter(x0[0], P0[0])
        feats = T.oracle_features(types[0])
        for t in range(1, warm + 1):
            filt, feats, _ = T.oracle_step(filt, feats, cam, seq.zc[t, 0], seq.has[t, 0], seq.U[0, t])
        x, P = filt.x_k_k.copy(), filt.p_k_k.copy()
        # move feature i so that its predicted pixel at x_k_km1 sits a hair inside the left border
        i = seed % N
        pos = 13 + 6 * i
        ft = T.oracle_filter(x, P)
        ft, _ = O.ekf_prediction(ft, T.oracle_features(types[0]))
        xp = ft.x_k_km1
        r_wc = O.q2r(xp[3:7])

        def u_of(theta):
            y = xp[pos:pos + 6].copy()
            y[3] = theta
            h = O.hi_inverse_depth(y, xp[0:3], r_wc, cam)
            return None if h is None else h[0]
        th = xp[pos + 3]
        lo, hi_ = th - 1.2, th
        if u_of(lo) is not None:
            continue
        for _ in range(80):
            mid = 0.5 * (lo + hi_)
            if u_of(mid) is None:
                lo = mid
            else:
                hi_ = mid
        target = hi_ + (seed % 7 + 1) * 2e-4          # 0.04 .. 0.3 px inside
        x2 = x.copy()
        x2[pos + 3] = target
        zc = seq.zc[warm + 1, 0].copy() if seq.zc.shape[0] > warm + 1 else seq.zc[warm, 0].copy()
        has = np.ones(N, dtype=np.uint8)
        # candidate of feature i: 1.6 px off vertically -> IC, but not a low-innovation inlier
        ft2 = T.oracle_filter(x2, P)
        fe2 = T.oracle_features(types[0])
        ft2, fe2 = O.ekf_prediction(ft2, fe2)
        fe2 = O.predict_and_derive(ft2, fe2, cam)
        if fe2[i].h is None:
            continue
        for j in range(N):
            if fe2[j].h is not None:
                zc[j] = fe2[j].h + 0.3 * np.array([np.sin(seed + j), np.cos(2 * seed + j)])
        zc[i] = fe2[i].h + np.array([0.05, 1.6])
        # push the update so that the camera estimate moves and feature i drifts out: bias all
        # other candidates to the left
        for j in range(N):
            if j != i:
                zc[j, 0] -= 0.8
        u = seq.U[0, warm]
        filt2 = T.oracle_filter(x2, P)
        feats2 = T.oracle_features(types[0])
        filt2, feats2, _ = T.oracle_step(filt2, feats2, cam, zc, has, u)
        f = feats2[i]
        # visible at x_k_km1 (has h), IC, not li; is it invisible at the state the rescue used?
        chk = T.oracle_filter(x2, P)
        fchk = T.oracle_features(types[0])
        chk, fchk = O.ekf_prediction(chk, fchk)
        fchk = O.search_IC_matches(chk, fchk, cam, (zc, has))
        fchk = O.ransac_hypotheses(chk, fchk, cam, u)
        chk = O.ekf_update_li_inliers(chk, fchk)
        r2 = O.q2r(chk.x_k_k[3:7])
        vis_after = O.hi_inverse_depth(chk.x_k_k[pos:pos + 6], chk.x_k_k[0:3], r2, cam)
        if f.individually_compatible == 1 and f.low_innovation_inlier == 0 and vis_after is None:
            print("stale-h frame found: seed %d feature %d hi=%d" % (seed, i, f.high_innovation_inlier))
            return dict(x0=x2[None], P0=P[None], types=types, zc=zc[None, None], has=has[None, None],
                        U=u[None, None], feature=i)
    raise RuntimeError("no stale-h frame found")


def case_stale_h():
    d = find_stale_h()
    run_case("ref_stale_h", d["x0"], d["P0"], d["types"], d["zc"], d["has"], d["U"],
             extra=dict(stale_feature=d["feature"]))


MAP_WORLD = dict(M=150, depth_range=(1.0, 3.0), speed=(0.05, 0.08), spring=0.004)


def pack_closed_loop(recs, NMAX, K):
    """list (per filter) of closed-loop records -> padded arrays [B, T, ...]."""
    B, Tn = len(recs), len(recs[0]["x"])
    n_max = 13 + 6 * NMAX
    d = dict(x=np.zeros((B, Tn, n_max)), nstate=np.zeros((B, Tn), dtype=np.int32), nfeat=np.zeros((B, Tn), dtype=np.int32),
             n_after_mm=np.zeros((B, Tn), dtype=np.int32),
             types=np.zeros((B, Tn, NMAX), dtype=np.uint8), tags=np.full((B, Tn, NMAX), -1, dtype=np.int32),
             flags=np.zeros((B, Tn, NMAX), dtype=np.uint8), nhyp=np.zeros((B, Tn), dtype=np.int32),
             counters=np.zeros((B, Tn, NMAX, 2), dtype=np.int32), det_uv=np.zeros((B, Tn, K, 2)),
             det_tag=np.full((B, Tn, K), -1, dtype=np.int32), det_n=np.zeros((B, Tn), dtype=np.int32),
             zc=np.zeros((B, Tn, NMAX, 2)), has=np.zeros((B, Tn, NMAX), dtype=np.uint8),
             U=np.zeros((B, Tn, len(recs[0]["U"][0]))))
    Ps = []
    for b, r in enumerate(recs):
        for t in range(Tn):
            n, nf = len(r["x"][t]), len(r["types"][t])
            assert nf <= NMAX
            d["x"][b, t, :n] = r["x"][t]
            d["nstate"][b, t], d["nfeat"][b, t], d["n_after_mm"][b, t] = n, nf, r["n_after_mm"][t]
            d["types"][b, t, :nf] = r["types"][t]
            d["tags"][b, t, :nf] = r["tags"][t]
            d["flags"][b, t, :nf] = r["flags"][t]
            d["nhyp"][b, t] = r["nhyp"][t]
            d["counters"][b, t, :nf] = r["counters"][t]
            d["det_uv"][b, t] = r["det_uv"][t]
            d["det_tag"][b, t] = r["det_tag"][t]
            d["det_n"][b, t] = r["det_n"][t]
            d["zc"][b, t, :nf] = r["zc"][t]
            d["has"][b, t, :nf] = r["has"][t]
            d["U"][b, t] = r["U"][t]
        P = np.zeros((n_max, n_max))
        n = r["P"].shape[0]
        P[:n, :n] = r["P"]
        Ps.append(P)
    d["P"] = np.array(Ps)
    return d


def case_map():
    """mono_slam.m:50-82 closed loop WITH the reference's own map_management.m (feature initialisation,
    inverse-depth -> Cartesian conversion, deletion, per-frame bookkeeping) over 80 frames."""
    B, Tn, K, NMAX, seed = 2, 80, 20, 64, 2
    world = synth.SynthWorld(B, T=Tn, seed=seed, **MAP_WORLD)
    I = R.make_interp()
    t0 = time.time()
    recs = []
    for b in range(B):
        recs.append(R.run_closed_loop(world, b, Tn, min_features=20, K=K, I=I))
        print("  ref_map: filter %d/%d done, %.1f s" % (b + 1, B, time.time() - t0), flush=True)
    d = pack_closed_loop(recs, NMAX, K)
    d.update(seed=seed, min_features=20, K=K)
    d["ref_functions_run"] = np.array(sorted(k for k in I.call_counts if k in I.sources_used and
                                             I.sources_used[k].startswith(R.REF_DIR)))
    path = os.path.join(HERE, "ref_map_t80.npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.0f KB)" % (path, os.path.getsize(path) / 1024))
    print("  conversions:", [(d["types"][b, -1] == 2).sum() for b in range(B)], "final nfeat", d["nfeat"][:, -1])


CASES = dict(cfg1=case_cfg1, outliers=case_outliers, mixed=case_mixed, n100=case_n100, stale_h=case_stale_h, map=case_map)

if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    for c in which:
        CASES[c]()
