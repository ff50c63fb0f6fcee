"""Shared test helpers: rebuild the reference's golden frame with the oracle."""
import numpy as np

from oracle import ekf_oracle as O


def build_golden_frame(golden):
    """Recipe of SURVEY §4: initialize_x_and_p -> add the 13 features from their
    uv_when_initialized (rho0=1, std_rho=1, std_pxl=1) -> one prediction."""
    cam = O.initialize_cam()
    x, P = O.initialize_x_and_p()
    filt = O.ekf_filter(x, P, 0.007, 0.007, 1.0, "constant_velocity")
    features_info = []
    for k in range(golden["uv_when_initialized"].shape[0]):
        uv = golden["uv_when_initialized"][k]
        X, Pn, newf = O.add_features_inverse_depth(uv, filt.x_k_k, filt.p_k_k, cam, filt.std_z, 1.0, 1.0)
        filt.x_k_k, filt.p_k_k = X, Pn
        features_info.append(O.new_feature_info(uv, X, int(golden["init_frame"][k]), newf))
    filt, features_info = O.ekf_prediction(filt, features_info)
    return cam, filt, features_info


# ---------------------------------------------------------------------------------------
# oracle <-> batched-array conversions used by the GPU parity tests
# ---------------------------------------------------------------------------------------
F_HAS_H, F_HAS_Z, F_IC, F_LI, F_HI, F_CAND = 1, 2, 4, 8, 16, 32


def oracle_filter(x, P, std_a=0.007, std_alpha=0.007, std_z=1.0):
    return O.ekf_filter(np.array(x), np.array(P), std_a, std_alpha, std_z, "constant_velocity")


def oracle_features(types):
    feats = []
    for t in types:
        if t == 0:
            continue
        f = O.Feature()
        f.type = "inversedepth" if t == 1 else "cartesian"
        f.individually_compatible = 0
        f.low_innovation_inlier = 0
        f.high_innovation_inlier = 0
        f.z = None
        f.h = None
        f.H = None
        f.S = None
        f.R = np.eye(2)
        f.times_predicted = 0
        f.times_measured = 0
        feats.append(f)
    return feats


def oracle_flags(feats, N):
    fl = np.zeros(N, dtype=np.uint8)
    for i, f in enumerate(feats):
        v = 0
        if f.h is not None:
            v |= F_HAS_H
        if f.z is not None:
            v |= F_HAS_Z
        if f.individually_compatible:
            v |= F_IC
        if f.low_innovation_inlier:
            v |= F_LI
        if f.high_innovation_inlier:
            v |= F_HI
        fl[i] = v
    return fl


def compact_H(feats):
    """[N,2,13] compact Jacobians (7 camera columns + the feature block) of the oracle's dense H."""
    out = np.zeros((len(feats), 2, 13))
    pos = 13
    for i, f in enumerate(feats):
        w = 6 if f.type == "inversedepth" else 3
        if f.H is not None:
            out[i, :, :7] = f.H[:, :7]
            out[i, :, 7:7 + w] = f.H[:, pos:pos + w]
        pos += w
    return out


def oracle_step(filt, feats, cam, zc, has, u, fixed=0):
    """One full reference step incl. the per-frame reset of mc/map_management.m:17."""
    feats = O.update_features_info(feats)
    info = {}
    filt, feats = O.filter_step(filt, feats, cam, (zc, has), u, fixed_hypotheses=fixed, info=info)
    return filt, feats, info


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ---------------------------------------------------------------------------------------
# closed loop with map management (mono_slam.m:50-82) on the oracle, fed with recorded inputs
# ---------------------------------------------------------------------------------------
def oracle_closed_loop(det_uv, det_tag, det_n, zc, has, U, frames, min_features=25, on_frame=None):
    """Drives oracle.map_management + the filter step with the detections / candidates / uniforms
    recorded by oracle/mref/run_ref.run_closed_loop (or downloaded from the GPU).  Yields nothing;
    returns the per-frame record (x, types, tags, flags, nhyp, counters) and the final P."""
    cam = O.initialize_cam()
    x, P = O.initialize_x_and_p()
    filt = O.ekf_filter(x, P, 0.007, 0.007, 1.0, "constant_velocity")
    feats = []
    out = dict(x=[], types=[], tags=[], flags=[], nhyp=[], counters=[], n_after_mm=[])
    for t in range(frames):
        nd = int(det_n[t])
        filt, feats = O.map_management(filt, feats, cam, (det_uv[t][:nd], det_tag[t][:nd]), min_features, t + 1)
        out["n_after_mm"].append(len(filt.x_k_k))
        nf = len(feats)
        info = {}
        filt, feats = O.filter_step(filt, feats, cam, (zc[t][:nf], has[t][:nf]), U[t], info=info)
        out["x"].append(filt.x_k_k.copy())
        out["types"].append(np.array([1 if f.type == "inversedepth" else 2 for f in feats], dtype=np.uint8))
        out["tags"].append(np.array([f.feature_when_initialized for f in feats], dtype=np.int32))
        out["flags"].append(oracle_flags(feats, nf))
        out["nhyp"].append(info.get("iterations", 0))
        out["counters"].append(np.array([[f.times_predicted, f.times_measured] for f in feats], dtype=np.float64))
        if on_frame is not None:
            on_frame(t, filt, feats)
    out["P"] = filt.p_k_k.copy()
    return out


def oracle_closed_loop_world(world, b, frames, min_features=25, K=25, n_u=64, u_seed=0, on_frame=None):
    """Self-driving variant: detections from the image of the previous frame and candidates of the current
    one are produced from `world` for whatever map the oracle holds (same protocol and the same uniform
    streams as oracle/mref/run_ref.run_closed_loop).  Returns the record plus the inputs it generated."""
    cam = O.initialize_cam()
    x, P = O.initialize_x_and_p()
    filt = O.ekf_filter(x, P, 0.007, 0.007, 1.0, "constant_velocity")
    feats = []
    rng = np.random.RandomState(u_seed + 7919 * b)
    out = dict(x=[], types=[], tags=[], flags=[], nhyp=[], counters=[], n_after_mm=[], det_uv=[], det_tag=[],
               det_n=[], zc=[], has=[], U=[])

    def tag_arrays():
        nf = len(feats)
        tg = np.full((world.B, max(nf, 1)), -1, dtype=np.int32)
        nfa = np.zeros(world.B, dtype=np.int32)
        tg[b, :nf] = [f.feature_when_initialized for f in feats]
        nfa[b] = nf
        return tg, nfa

    for step in range(1, frames + 1):
        tg, nfa = tag_arrays()
        uv, dtag, nd = world.detections(step - 1, tg, nfa, K)
        filt, feats = O.map_management(filt, feats, cam, (uv[b, :nd[b]], dtag[b, :nd[b]]), min_features, step)
        out["n_after_mm"].append(len(filt.x_k_k))
        tg, nfa = tag_arrays()
        nf = len(feats)
        zc, has = world.candidates(step, tg, nfa)
        u = rng.rand(n_u)
        info = {}
        filt, feats = O.filter_step(filt, feats, cam, (zc[b, :nf], has[b, :nf]), u, info=info)
        out["x"].append(filt.x_k_k.copy())
        out["types"].append(np.array([1 if f.type == "inversedepth" else 2 for f in feats], dtype=np.uint8))
        out["tags"].append(np.array([f.feature_when_initialized for f in feats], dtype=np.int32))
        out["flags"].append(oracle_flags(feats, nf))
        out["nhyp"].append(info.get("iterations", 0))
        out["counters"].append(np.array([[f.times_predicted, f.times_measured] for f in feats], dtype=np.float64))
        out["det_uv"].append(uv[b].copy()); out["det_tag"].append(dtag[b].copy()); out["det_n"].append(int(nd[b]))
        out["zc"].append(zc[b, :nf].copy()); out["has"].append(has[b, :nf].copy()); out["U"].append(u)
        if on_frame is not None:
            on_frame(step - 1, filt, feats)
    out["P"] = filt.p_k_k.copy()
    return out
