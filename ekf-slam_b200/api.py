"""The reference's function-level API on the GPU path — same names, argument order and struct
field names as diwakar-vsingh/EKF-SLAM ``matlab_code`` so a per-frame driver written against the
reference (mc/mono_slam.m:50-82) runs unchanged:

    cam = initialize_cam()
    filter = ekf_filter(x_k_k, p_k_k, sigma_a, sigma_alpha, sigma_image_noise, 'constant_velocity')
    filter, features_info = ekf_prediction(filter, features_info)
    features_info = search_IC_matches(filter, features_info, cam, im)
    features_info = ransac_hypotheses(filter, features_info, cam)
    filter = ekf_update_li_inliers(filter, features_info)
    features_info = rescue_hi_inliers(filter, features_info, cam)
    filter = ekf_update_hi_inliers(filter, features_info)

Value semantics like MATLAB: every call marshals the structs to the device, runs the stage's
kernels through the C ABI and marshals the result back (for throughput use :class:`FilterBank`,
which keeps B filters resident).  ``features_info`` is a list of :class:`Feature`; "empty"
(``[]``) is ``None``.  ``H`` is stored dense 2 x n like ``full(features_info(i).H)``.

Differences that a drop-in user must know (all stem from gaps in the reference itself):
  * ``im`` of search_IC_matches is not an image but the candidate pixels ``(z_cand [N,2],
    has_cand [N])`` for the synthetic matcher gate (the reference's matcher needs MATLAB's
    Computer Vision Toolbox, mc/matching.m:29-46).
  * ransac_hypotheses takes the uniform stream ``u`` explicitly (``rand(1)`` of
    mc/select_random_match.m:12); if omitted it is drawn from ``numpy.random``.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from . import _lib as L
from .bank import FilterBank
from .synth import default_camera


class Feature(SimpleNamespace):
    """features_info(i) — field names of mc/add_feature_to_info_vector.m:7-32."""


class Filter(SimpleNamespace):
    """The `filter` struct — field names of mc/ekf_filter.m:37-59."""


def initialize_cam():
    """mc/initialize_cam.m."""
    return default_camera()


def initialize_x_and_p():
    """mc/initialize_x_and_p.m:3-24."""
    eps = float(np.finfo(np.float64).eps)
    x = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1e-15, 1e-15, 1e-15], dtype=np.float64)
    p = np.diag([eps] * 7 + [0.025 ** 2] * 6)
    return x, p


def ekf_filter(x_k_k, p_k_k, std_a, std_alpha, std_z, type_="constant_velocity"):
    """mc/ekf_filter.m:37-59."""
    if type_ != "constant_velocity":
        raise NotImplementedError("only the 'constant_velocity' motion model of mc/mono_slam.m:32 is built")
    f = Filter(type=type_, x_k_k=np.array(x_k_k, dtype=np.float64), p_k_k=np.array(p_k_k, dtype=np.float64),
               std_a=float(std_a), std_alpha=float(std_alpha), std_z=float(std_z), x_k_km1=None, p_k_km1=None)
    for name in ("predicted_measurements", "H_predicted", "R_predicted", "S_predicted", "S_matching", "z", "h",
                 "H_matching", "measurements", "R_matching", "x_k_k_mixing_estimate", "p_k_k_mixing_covariance"):
        setattr(f, name, None)
    return f


def new_feature(type_="inversedepth", yi=None, uv=None, step=0):
    """A features_info element as mc/add_feature_to_info_vector.m:7-32 creates it (no image patch)."""
    return Feature(type=type_, yi=None if yi is None else np.array(yi, dtype=np.float64),
                   uv_when_initialized=None if uv is None else np.array(uv, dtype=np.float64),
                   init_frame=step, times_predicted=0, times_measured=0, individually_compatible=0,
                   low_innovation_inlier=0, high_innovation_inlier=0, z=None, h=None, H=None, S=None,
                   state_size=6 if type_ == "inversedepth" else 3, measurement_size=2, R=np.eye(2),
                   half_patch_size_when_initialized=20, half_patch_size_when_matching=6)


# ------------------------------------------------------------------------------------------
# marshalling
# ------------------------------------------------------------------------------------------
_banks = {}


def _bank(N, n, f=None, cam=None):
    key = (N, n)
    bank = _banks.get(key)
    if bank is None:
        if len(_banks) > 8:
            _banks.pop(next(iter(_banks))).close()
        bank = _banks[key] = FilterBank(1, max(N, 1), n)
    if cam is not None:
        bank.set_camera(cam)
    if f is not None:
        bank.set_params(std_a=f.std_a, std_alpha=f.std_alpha, std_z=f.std_z)
    return bank


def _types(features_info):
    t = np.zeros((1, max(len(features_info), 1)), dtype=np.uint8)
    for i, fi in enumerate(features_info):
        if fi.type == "inversedepth":
            t[0, i] = L.FEAT_INVERSEDEPTH
        elif fi.type == "cartesian":
            t[0, i] = L.FEAT_CARTESIAN
        else:
            raise ValueError("feature type %r" % (fi.type,))
    return t


def _state_dim(features_info):
    return 13 + sum(6 if fi.type == "inversedepth" else 3 for fi in features_info)


def _offsets(features_info):
    offs, pos = [], 13
    for fi in features_info:
        offs.append(pos)
        pos += 6 if fi.type == "inversedepth" else 3
    return offs


def _push_features(bank, features_info):
    """features_info -> device arrays (h, compact H, S, z, flag bytes)."""
    N = bank.N
    h = np.zeros((1, N, 2))
    Hc = np.zeros((1, N, 2, 13))
    S = np.zeros((1, N, 2, 2))
    z = np.zeros((1, N, 2))
    fl = np.zeros((1, N), dtype=np.uint8)
    offs = _offsets(features_info)
    for i, fi in enumerate(features_info):
        w = 6 if fi.type == "inversedepth" else 3
        if fi.h is not None:
            h[0, i] = np.asarray(fi.h, dtype=np.float64).reshape(2)
            fl[0, i] |= L.F_HAS_H
        if fi.H is not None:
            H = np.asarray(fi.H.todense() if hasattr(fi.H, "todense") else fi.H, dtype=np.float64)
            Hc[0, i, :, :7] = H[:, :7]
            Hc[0, i, :, 7:7 + w] = H[:, offs[i]:offs[i] + w]
        if fi.S is not None:
            S[0, i] = np.asarray(fi.S, dtype=np.float64)
        if fi.z is not None:
            z[0, i] = np.asarray(fi.z, dtype=np.float64).reshape(2)
            fl[0, i] |= L.F_HAS_Z
        if fi.individually_compatible:
            if fi.z is None:
                raise ValueError("feature %d is individually_compatible but has no z" % i)
            fl[0, i] |= L.F_IC
        if fi.low_innovation_inlier:
            fl[0, i] |= L.F_LI
        if fi.high_innovation_inlier:
            fl[0, i] |= L.F_HI
    bank.upload_features(h=h, Hc=Hc, S=S, z=z, flags=fl)


def _pull_features(bank, features_info, n, fields=("h", "H", "S", "z", "flags")):
    d = bank.download_features()
    offs = _offsets(features_info)
    for i, fi in enumerate(features_info):
        fl = int(d["flags"][0, i])
        w = 6 if fi.type == "inversedepth" else 3
        if "h" in fields:
            fi.h = d["h"][0, i].copy() if fl & L.F_HAS_H else None
        if "H" in fields:
            if fl & L.F_HAS_H:
                H = np.zeros((2, n))
                H[:, :7] = d["Hc"][0, i, :, :7]
                H[:, offs[i]:offs[i] + w] = d["Hc"][0, i, :, 7:7 + w]
                fi.H = H
            else:
                fi.H = None
        if "S" in fields:
            fi.S = d["S"][0, i].copy() if fl & L.F_HAS_H else None
        if "z" in fields:
            fi.z = d["z"][0, i].copy() if fl & L.F_HAS_Z else None
        if "flags" in fields:
            fi.individually_compatible = 1 if fl & L.F_IC else 0
            fi.low_innovation_inlier = 1 if fl & L.F_LI else 0
            fi.high_innovation_inlier = 1 if fl & L.F_HI else 0
    return features_info


def _setup(features_info, x, P=None, which=0, f=None, cam=None):
    n = _state_dim(features_info)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if x.shape[0] != n:
        raise ValueError("state vector has %d entries, features_info implies %d" % (x.shape[0], n))
    bank = _bank(len(features_info), n, f, cam)
    bank.upload_feature_types(_types(features_info), nfeat=np.array([len(features_info)], dtype=np.int32))
    bank.upload_state(x=x[None], P=None if P is None else np.asarray(P, dtype=np.float64)[None], which=which)
    return bank, n


# ------------------------------------------------------------------------------------------
# the hot-path functions
# ------------------------------------------------------------------------------------------
def ekf_prediction(f, features_info):
    """mc/ekf_prediction.m:1-3 -> mc/predict_state_and_covariance.m."""
    bank, n = _setup(features_info, f.x_k_k, f.p_k_k, which=0, f=f)
    bank.ekf_prediction()
    x, P, _ = bank.download_state(which=1)
    f.x_k_km1, f.p_k_km1 = x[0], P[0]
    return f, features_info


def predict_camera_measurements(x_k_k, cam, features_info):
    """mc/predict_camera_measurements.m:1-28 (h is only overwritten for visible features)."""
    bank, n = _setup(features_info, x_k_k, None, which=1, cam=cam)
    _push_features(bank, features_info)
    bank.features(which=1, parts=1)
    return _pull_features(bank, features_info, n, fields=("h",))


def calculate_derivatives(x_k_km1, cam, features_info):
    """mc/calculate_derivatives.m:1-28 (H linearised at the stored h)."""
    bank, n = _setup(features_info, x_k_km1, None, which=1, cam=cam)
    _push_features(bank, features_info)
    bank.features(which=1, parts=2)
    return _pull_features(bank, features_info, n, fields=("H",))


def search_IC_matches(f, features_info, cam, im):
    """mc/search_IC_matches.m:1-17; ``im`` = (z_cand [N,2], has_cand [N]) for the matcher gate."""
    bank, n = _setup(features_info, f.x_k_km1, f.p_k_km1, which=1, f=f, cam=cam)
    _push_features(bank, features_info)
    bank.measure(which=1)
    if im is not None:
        z_cand, has_cand = im
        N = bank.N
        zc = np.zeros((1, N, 2))
        hs = np.zeros((1, N), dtype=np.uint8)
        zc[0, :len(features_info)] = np.asarray(z_cand, dtype=np.float64)[:len(features_info)]
        hs[0, :len(features_info)] = np.asarray(has_cand)[:len(features_info)] != 0
        bank.upload_candidates(zc, hs)
        bank.gate()
    return _pull_features(bank, features_info, n)


def ransac_hypotheses(f, features_info, cam, u=None, fixed_hypotheses=0, info=None):
    """mc/ransac_hypotheses.m:1-47."""
    bank, n = _setup(features_info, f.x_k_km1, f.p_k_km1, which=1, f=f, cam=cam)
    _push_features(bank, features_info)
    if u is None:
        u = np.random.rand(fixed_hypotheses if fixed_hypotheses > 0 else 1000)
    bank.set_params(fixed_hyp=int(fixed_hypotheses))
    bank.upload_uniforms(np.asarray(u, dtype=np.float64)[None])
    bank.ransac_hypotheses()
    st = bank.download_stats()
    bank.set_params(fixed_hyp=0)
    if st["status"][0] & 1:
        raise RuntimeError("uniform stream exhausted after %d hypotheses" % st["ransac_iters"][0])
    if info is not None:
        info["iterations"] = int(st["ransac_iters"][0])
        info["max_support"] = int(st["max_support"][0])
        info["scored"] = int(st["ransac_scored"][0])
    return _pull_features(bank, features_info, n, fields=("flags",))


def _update(f, features_info, mask, which_prior):
    x = f.x_k_km1 if which_prior else f.x_k_k
    P = f.p_k_km1 if which_prior else f.p_k_k
    bank, n = _setup(features_info, x, P, which=which_prior, f=f)
    _push_features(bank, features_info)
    bank.hp(need=mask)
    bank.update_masked(mask, which_prior)
    xs, Ps, _ = bank.download_state(which=0)
    f.x_k_k, f.p_k_k = xs[0], Ps[0]
    return f


def ekf_update_li_inliers(f, features_info):
    """mc/ekf_update_li_inliers.m:1-21 -> mc/update.m."""
    return _update(f, features_info, L.F_LI, 1)


def rescue_hi_inliers(f, features_info, cam):
    """mc/rescue_hi_inliers.m:1-22."""
    bank, n = _setup(features_info, f.x_k_k, f.p_k_k, which=0, f=f, cam=cam)
    _push_features(bank, features_info)
    bank.rescue_hi_inliers()
    return _pull_features(bank, features_info, n, fields=("h", "H", "flags"))


def ekf_update_hi_inliers(f, features_info):
    """mc/ekf_update_hi_inliers.m:1-21 -> mc/update.m."""
    return _update(f, features_info, L.F_HI, 0)


def ekf_update_iterated(f, features_info, cam, flag="low_innovation_inlier", n_iter=3):
    """mc/ekf_update_iterated.m names an ``update_iterated`` that the reference never ships; this is the
    standard iterated EKF over the features whose ``flag`` field is 1 (EXTENSION — no reference
    semantics exist).  Prior: (x_k_km1, p_k_km1) for the li flag, (x_k_k, p_k_k) otherwise."""
    mask = {"low_innovation_inlier": L.F_LI, "high_innovation_inlier": L.F_HI}[flag]
    which_prior = 1 if mask == L.F_LI else 0
    x = f.x_k_km1 if which_prior else f.x_k_k
    P = f.p_k_km1 if which_prior else f.p_k_k
    bank, n = _setup(features_info, x, P, which=which_prior, f=f, cam=cam)
    _push_features(bank, features_info)
    bank.update_iterated(mask, which_prior, n_iter)
    xs, Ps, _ = bank.download_state(which=0)
    f.x_k_k, f.p_k_k = xs[0], Ps[0]
    return f


def update_features_info(features_info):
    """mc/update_features_info.m:1-18 (host-side bookkeeping of the struct array)."""
    for fi in features_info:
        if fi.h is not None:
            fi.times_predicted += 1
        if fi.low_innovation_inlier or fi.high_innovation_inlier:
            fi.times_measured += 1
        fi.individually_compatible = 0
        fi.low_innovation_inlier = 0
        fi.high_innovation_inlier = 0
        fi.h = None
        fi.z = None
        fi.H = None
        fi.S = None
    return features_info


def filter_step(f, features_info, cam, im, u=None, fixed_hypotheses=0, info=None):
    """mc/mono_slam.m:56-74 (without takeImage) in one marshalling round trip."""
    bank, n = _setup(features_info, f.x_k_k, f.p_k_k, which=0, f=f, cam=cam)
    _push_features(bank, features_info)
    z_cand, has_cand = im
    N = bank.N
    zc = np.zeros((1, N, 2))
    hs = np.zeros((1, N), dtype=np.uint8)
    zc[0, :len(features_info)] = np.asarray(z_cand, dtype=np.float64)[:len(features_info)]
    hs[0, :len(features_info)] = np.asarray(has_cand)[:len(features_info)] != 0
    bank.upload_candidates(zc, hs)
    if u is None:
        u = np.random.rand(fixed_hypotheses if fixed_hypotheses > 0 else 1000)
    bank.set_params(fixed_hyp=int(fixed_hypotheses))
    bank.upload_uniforms(np.asarray(u, dtype=np.float64)[None])
    bank.step(reset=False, match_mode=1)
    bank.set_params(fixed_hyp=0)
    xp, _, _ = bank.download_state(which=1, want_P=False)
    x, P, _ = bank.download_state(which=0)
    f.x_k_km1, f.x_k_k, f.p_k_k = xp[0], x[0], P[0]
    f.p_k_km1 = None  # the fused step updates the covariance in place; the intermediate is not kept
    st = bank.download_stats()
    if info is not None:
        info.update({k: int(v[0]) for k, v in st.items()})
    return f, _pull_features(bank, features_info, n)


# ------------------------------------------------------------------------------------------
# map management: mc/map_management.m and the functions it calls, on the device ops
# ekfslam_map_management / ekfslam_add_features / ekfslam_inversedepth_2_cartesian / ekfslam_delete_features
# ------------------------------------------------------------------------------------------
def _q2r(q):
    r, x, y, z = q
    return np.array([[r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
                     [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
                     [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def _setup_map(features_info, x, P, extra, f=None, cam=None):
    """Like _setup, but with room for `extra` new features; uploads (x, P) as (x_k_k, p_k_k)."""
    n = _state_dim(features_info)
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if x.shape[0] != n:
        raise ValueError("state vector has %d entries, features_info implies %d" % (x.shape[0], n))
    N = len(features_info) + int(extra)
    n_max = 13 + 6 * max(N, 1)
    bank = _bank(max(N, 1), n_max, f, cam)
    t = np.zeros((1, bank.N), dtype=np.uint8)
    t[0, :len(features_info)] = _types(features_info)[0, :len(features_info)]
    bank.upload_feature_types(t, nfeat=np.array([len(features_info)], dtype=np.int32))
    xs = np.zeros((1, n_max))
    Ps = np.zeros((1, n_max, n_max))
    xs[0, :n] = x
    Ps[0, :n, :n] = np.asarray(P, dtype=np.float64)
    bank.upload_state(x=xs, P=Ps, which=0)
    return bank, n


def _pull_map(bank):
    x, P, ns = bank.download_state(which=0)
    n = int(ns[0])
    types, nf = bank.download_feature_types()
    return x[0, :n].copy(), P[0, :n, :n].copy(), types[0, :int(nf[0])].copy()


def add_features_inverse_depth(uvd, X, P, cam, std_pxl, initial_rho, std_rho):
    """mc/add_features_inverse_depth.m:1-24 -> mc/hinv.m + mc/add_a_feature_covariance_inverse_depth.m.
    uvd: [2] or [2, nNew] distorted pixels.  Returns (X_RES, P_RES, newFeature) (newFeature of the last one)."""
    uvd = np.asarray(uvd, dtype=np.float64).reshape(2, -1)
    X = np.asarray(X, dtype=np.float64).reshape(-1)
    if uvd.shape[1] == 0:
        return X, np.asarray(P, dtype=np.float64), None
    if (X.shape[0] - 13) % 3:
        raise ValueError("state size %d is not 13 + 6*N_id + 3*N_c" % X.shape[0])
    # only the state SIZE matters for an append: describe the existing map as Cartesian triples
    layout = [new_feature("cartesian") for _ in range((X.shape[0] - 13) // 3)]
    bank, n = _setup_map(layout, X, P, extra=uvd.shape[1], cam=cam)
    for j in range(uvd.shape[1]):
        bank.add_features_inverse_depth(uvd[:, j][None], std_pxl=std_pxl, initial_rho=initial_rho, std_rho=std_rho)
    x, Pn, _ = _pull_map(bank)
    return x, Pn, x[-6:].copy()


def inversedepth_2_cartesian(f, features_info):
    """mc/inversedepth_2_cartesian.m:1-52 (at most one conversion per call, :49)."""
    if not features_info:
        return f, features_info
    bank, n = _setup_map(features_info, f.x_k_k, f.p_k_k, extra=0, f=f)
    conv = bank.inversedepth_2_cartesian(threshold=0.1)
    if conv[0] >= 0:
        f.x_k_k, f.p_k_k, _ = _pull_map(bank)
        features_info[int(conv[0])].type = "cartesian"
    return f, features_info


def delete_a_feature(X_km1_km1, P_km1_km1, featToDelete, features_info):
    """mc/delete_a_feature.m:1-25; featToDelete is the 0-based position in the Python list."""
    bank, n = _setup_map(features_info, X_km1_km1, P_km1_km1, extra=0)
    d = np.zeros((1, bank.N), dtype=np.uint8)
    d[0, int(featToDelete)] = 1
    bank.delete_features(d)
    x, P, _ = _pull_map(bank)
    return x, P


def delete_features(f, features_info):
    """Called by mc/map_management.m:7 but NOT shipped by the reference; rule of the published toolbox it derives
    from: delete a feature predicted more than 5 times and matched in fewer than half of those predictions."""
    dele = [fi.times_measured < 0.5 * fi.times_predicted and fi.times_predicted > 5 for fi in features_info]
    if any(dele):
        bank, n = _setup_map(features_info, f.x_k_k, f.p_k_k, extra=0, f=f)
        d = np.zeros((1, bank.N), dtype=np.uint8)
        d[0, :len(dele)] = dele
        bank.delete_features(d)
        f.x_k_k, f.p_k_k, _ = _pull_map(bank)
        features_info = [fi for fi, dd in zip(features_info, dele) if not dd]
    return f, features_info


def add_feature_to_info_vector(uv, im_k, X_RES, features_info, step, newFeature, init_feature_descriptor):
    """mc/add_feature_to_info_vector.m:1-32 (host bookkeeping; the image patch copy of :7 is dropped)."""
    fi = new_feature("inversedepth", yi=newFeature, uv=uv, step=step)
    X_RES = np.asarray(X_RES, dtype=np.float64).reshape(-1)
    fi.r_wc_when_initialized = X_RES[0:3].copy()
    fi.R_wc_when_initialized = _q2r(X_RES[3:7])
    fi.init_measurement = np.asarray(uv, dtype=np.float64).reshape(2)
    fi.feature_when_initialized = init_feature_descriptor
    return list(features_info) + [fi]


def _detections(im):
    """`im` of the map functions: corner detections uv [K,2] or (uv [K,2], descriptors [K]) — the stand-in for the
    image the reference searches with detectFASTFeatures (mc/initialize_a_feature.m:29)."""
    if isinstance(im, tuple):
        uv, desc = im
    else:
        uv, desc = im, None
    uv = np.asarray(uv, dtype=np.float64).reshape(-1, 2)
    desc = list(range(len(uv))) if desc is None else list(desc)
    return uv, desc


def initialize_features(step, cam, f, features_info, num_features_to_initialize, im):
    """mc/initialize_features.m:1-21 with the corner search of mc/initialize_a_feature.m replaced by the supplied
    detections (one attempt each, at most 50 attempts)."""
    uv, desc = _detections(im)
    k = int(min(num_features_to_initialize, len(uv), 50))
    for j in range(k):
        X_RES, P_RES, newFeature = add_features_inverse_depth(uv[j], f.x_k_k, f.p_k_k, cam, f.std_z, 1.0, 1.0)
        f.x_k_k, f.p_k_k = X_RES, P_RES
        features_info = add_feature_to_info_vector(uv[j], None, X_RES, features_info, step, newFeature, desc[j])
    return f, features_info


def map_management(f, features_info, cam, im, min_number_of_features_in_image, step):
    """mc/map_management.m:1-35 in ONE device call (ekfslam_map_management): delete_features ->
    count measured -> update_features_info -> inversedepth_2_cartesian -> initialize_features.
    ``im`` = corner detections, see :func:`_detections`."""
    uv, desc = _detections(im)
    K = max(len(uv), 1)
    nf0 = len(features_info)
    bank, n = _setup_map(features_info, f.x_k_k, f.p_k_k, extra=min(K, 50), f=f, cam=cam)
    _push_features(bank, features_info)
    counters = np.zeros((1, bank.N, 2), dtype=np.int32)
    tags = np.full((1, bank.N), -1, dtype=np.int32)
    for i, fi in enumerate(features_info):
        counters[0, i] = (fi.times_predicted, fi.times_measured)
        tags[0, i] = i                                   # old features: position in the old list
    bank.upload_feature_meta(counters=counters, tag=tags)
    uvp = np.zeros((1, K, 2))
    uvp[0, :len(uv)] = uv
    dtag = (nf0 + np.arange(K, dtype=np.int32))[None]     # new features: nf0 + detection index
    bank.upload_detections(uvp, np.array([len(uv)], dtype=np.int32), tag=dtag)
    bank.map_management(int(min_number_of_features_in_image))
    x, P, types = _pull_map(bank)
    tg = bank.download_feature_tags()[0, :len(types)]
    cnt = bank.download_features()["counters"][0]
    out = []
    pos = 13
    for i, (ty, tag) in enumerate(zip(types, tg)):
        w = 6 if ty == L.FEAT_INVERSEDEPTH else 3
        if tag < nf0:
            fi = features_info[int(tag)]
            fi.type = "inversedepth" if ty == L.FEAT_INVERSEDEPTH else "cartesian"
        else:
            j = int(tag) - nf0
            fi = add_feature_to_info_vector(uv[j], None, x, [], step, x[pos:pos + 6].copy(), desc[j])[0]
        fi.times_predicted, fi.times_measured = int(cnt[i, 0]), int(cnt[i, 1])
        fi.individually_compatible = fi.low_innovation_inlier = fi.high_innovation_inlier = 0
        fi.h = fi.z = fi.H = fi.S = None
        out.append(fi)
        pos += w
    f.x_k_k, f.p_k_k = x, P
    return f, out
