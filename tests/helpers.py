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
