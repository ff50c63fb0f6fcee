"""CPU oracle (numpy fp64) for the 1-point-RANSAC EKF filter step.

TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``ekf-slam_b200/``) never imports anything from ``oracle/``.

It restates, function by function and in the reference's own operation order,
the MATLAB hot path of diwakar-vsingh/EKF-SLAM (``matlab_code/*.m``; cited as
``mc/<file>:<lines>``).  Parity status:

* PINNED by the reference's own artefact ``matlab_code/features_information.mat``
  (tests/golden/features_information.npz, tests/test_oracle_golden.py):
  ``hinv``, ``add_a_feature_covariance_inverse_depth``,
  ``predict_state_and_covariance``, ``hi_inverse_depth``,
  ``calculate_Hi_inverse_depth`` and ``S_i = H P H' + R`` reproduce the stored
  ``yi``/``h``/``H``/``S`` of all 13 features to <= 1e-12.
* PINNED by the reference's own EXECUTION (round 2): the unmodified ``matlab_code/*.m`` sources are
  run through ``oracle/mref`` (mini MATLAB interpreter) over synthetic frames; its outputs are the
  fixtures ``tests/golden/ref_*.npz`` (``tests/golden/make_ref_steps.py``).  Against them this file
  reproduces the RANSAC selection and adaptive hypothesis count, ``update`` / ``normJac``,
  ``rescue_hi_inliers`` (incl. the stale-``h`` edge), the hi update, the Cartesian model and the map
  management functions: flags / counts exact, x and P <= 1e-11 over 200 free-running frames
  (tests/test_oracle_ref.py, tests/test_oracle_map_ref.py).
* PARITY UNPINNED by construction: ``update_iterated`` only (the reference names it but ships no
  implementation, so there is nothing to execute).

Two functions the reference calls but does not ship are restated from their
published definitions: ``quaternions(v, theta)`` (mc/v2q.m:15) and
``dq3_by_dq1(q)`` (mc/dfv_by_dxv.m:13, mc/func_Q.m:24; derived from the
product rule of mc/qprod.m:8).  ``update_iterated`` (mc/ekf_update_iterated.m:3)
does not exist in the reference at all; ``update_iterated`` below is the
standard IEKF and is an extension with no reference semantics to match.

``features_info`` is a python list of :class:`Feature` objects whose attribute
names equal the MATLAB struct field names (mc/add_feature_to_info_vector.m:7-32);
"empty" (``[]``) is ``None``.  ``filter`` is a :class:`Filter` with the field
names of mc/ekf_filter.m:37-59.  Vectors are 1-D numpy arrays.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

EPS = float(np.finfo(np.float64).eps)
CHI2INV_2_95 = 5.9915  # mc/rescue_hi_inliers.m:3, mc/matching.m:2


# --------------------------------------------------------------------------
# L0  containers
# --------------------------------------------------------------------------
def initialize_cam():
    """mc/initialize_cam.m:3-25."""
    d = 0.0112
    cam = SimpleNamespace()
    cam.k1 = 6.333e-2
    cam.k2 = 1.390e-2
    cam.nRows = 240
    cam.nCols = 320
    cam.Cx = 1.7945 / d
    cam.Cy = 1.4433 / d
    cam.f = 2.1735
    cam.dx = d
    cam.dy = d
    cam.model = "two_distortion_parameters"
    cam.K = np.array([[cam.f / d, 0.0, cam.Cx], [0.0, cam.f / d, cam.Cy], [0.0, 0.0, 1.0]])
    return cam


def initialize_x_and_p():
    """mc/initialize_x_and_p.m:3-24."""
    v_0, std_v_0, w_0, std_w_0 = 0.0, 0.025, 1e-15, 0.025
    x = np.array([0, 0, 0, 1, 0, 0, 0, v_0, v_0, v_0, w_0, w_0, w_0], dtype=np.float64)
    p = np.zeros((13, 13))
    for i in range(7):
        p[i, i] = EPS
    for i in range(7, 10):
        p[i, i] = std_v_0 ** 2
    for i in range(10, 13):
        p[i, i] = std_w_0 ** 2
    return x, p


class Filter(SimpleNamespace):
    """mc/ekf_filter.m:37-59 (a plain struct in the reference)."""


def ekf_filter(x_k_k, p_k_k, std_a, std_alpha, std_z, type_):
    f = Filter()
    f.type = type_
    f.x_k_k = np.array(x_k_k, dtype=np.float64)
    f.p_k_k = np.array(p_k_k, dtype=np.float64)
    f.std_a = std_a
    f.std_alpha = std_alpha
    f.std_z = std_z
    f.x_k_km1 = None
    f.p_k_km1 = None
    for name in ("predicted_measurements", "H_predicted", "R_predicted", "S_predicted",
                 "S_matching", "z", "h", "H_matching", "measurements", "R_matching",
                 "x_k_k_mixing_estimate", "p_k_k_mixing_covariance"):
        setattr(f, name, None)
    return f


class Feature(SimpleNamespace):
    """One element of the features_info struct array."""


def new_feature_info(uv, X_RES, step, newFeature):
    """mc/add_feature_to_info_vector.m:7-32 (image patch / descriptor fields dropped)."""
    fi = Feature()
    fi.r_wc_when_initialized = np.array(X_RES[0:3])
    fi.R_wc_when_initialized = q2r(X_RES[3:7])
    fi.uv_when_initialized = np.array(uv, dtype=np.float64).reshape(2)
    fi.half_patch_size_when_initialized = 20
    fi.half_patch_size_when_matching = 6
    fi.times_predicted = 0
    fi.times_measured = 0
    fi.init_frame = step
    fi.init_measurement = np.array(uv, dtype=np.float64).reshape(2)
    fi.type = "inversedepth"
    fi.yi = np.array(newFeature, dtype=np.float64)
    fi.individually_compatible = 0
    fi.low_innovation_inlier = 0
    fi.high_innovation_inlier = 0
    fi.z = None
    fi.h = None
    fi.H = None
    fi.S = None
    fi.state_size = 6
    fi.measurement_size = 2
    fi.R = np.eye(2)
    return fi


def update_features_info(features_info):
    """mc/update_features_info.m:4-18."""
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


# --------------------------------------------------------------------------
# L1  quaternion / rotation primitives
# --------------------------------------------------------------------------
def q2r(q):
    """mc/q2r.m:3-10."""
    r, x, y, z = q[0], q[1], q[2], q[3]
    return np.array([
        [r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
        [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
        [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def qprod(q, p):
    """mc/qprod.m:3-8."""
    a = q[0]
    v = np.asarray(q[1:4], dtype=np.float64)
    x = p[0]
    u = np.asarray(p[1:4], dtype=np.float64)
    out = np.empty(4)
    out[0] = a * x - v.dot(u)
    out[1:4] = (a * u + x * v) + np.cross(v, u)
    return out


def qconj(q):
    """mc/qconj.m:3-4."""
    qb = -np.asarray(q, dtype=np.float64)
    qb[0] = q[0]
    return qb


def quaternions(v_n, theta):
    """MISSING from the reference (called at mc/v2q.m:15).  Standard axis-angle:
    q = [cos(theta/2), sin(theta/2) * v_n]."""
    s = math.sin(theta / 2.0)
    return np.array([math.cos(theta / 2.0), s * v_n[0], s * v_n[1], s * v_n[2]])


def v2q(v):
    """mc/v2q.m:10-16."""
    theta = float(np.linalg.norm(v))
    if theta < EPS:
        return np.array([1.0, 0.0, 0.0, 0.0])
    v_n = np.asarray(v) / np.linalg.norm(v)
    return quaternions(v_n, theta)


def dq3_by_dq2(q1):
    """mc/dq3_by_dq2.m:8-12  (d(q1*q2)/dq2)."""
    R, X, Y, Z = q1[0], q1[1], q1[2], q1[3]
    return np.array([[R, -X, -Y, -Z],
                     [X, R, Z, -Y],
                     [Y, -Z, R, X],
                     [Z, Y, -X, R]])


def dq3_by_dq1(q2):
    """MISSING from the reference (called at mc/dfv_by_dxv.m:13, mc/func_Q.m:24).
    d(q1*q2)/dq1 for the product of mc/qprod.m:8.  NOTE the reference passes qOld
    (the LEFT factor) here where the chain rule wants d(qOld*qwt)/d(qwt) — i.e.
    the left-multiplication matrix of qOld; Davison's SceneLib convention, which
    this restates: [r -x -y -z; x r -z y; y z r -x; z -y x r]."""
    R, X, Y, Z = q2[0], q2[1], q2[2], q2[3]
    return np.array([[R, -X, -Y, -Z],
                     [X, R, -Z, Y],
                     [Y, Z, R, -X],
                     [Z, -Y, X, R]])


def dqomegadt_by_domega(omega, delta_t):
    """mc/dqomegadt_by_domega.m:6-48."""
    omegamod = float(np.linalg.norm(omega))

    def dq0_by_domegaA(omegaA, om, dt):  # :31-33
        return (-dt / 2.0) * (omegaA / om) * math.sin(om * dt / 2.0)

    def dqA_by_domegaA(omegaA, om, dt):  # :36-40
        return ((dt / 2.0) * omegaA * omegaA / (om * om) * math.cos(om * dt / 2.0)
                + (1.0 / om) * (1.0 - omegaA * omegaA / (om * om)) * math.sin(om * dt / 2.0))

    def dqA_by_domegaB(omegaA, omegaB, om, dt):  # :44-48
        return (omegaA * omegaB / (om * om)) * (
            (dt / 2.0) * math.cos(om * dt / 2.0) - (1.0 / om) * math.sin(om * dt / 2.0))

    o = omega
    J = np.zeros((4, 3))
    J[0, 0] = dq0_by_domegaA(o[0], omegamod, delta_t)
    J[0, 1] = dq0_by_domegaA(o[1], omegamod, delta_t)
    J[0, 2] = dq0_by_domegaA(o[2], omegamod, delta_t)
    J[1, 0] = dqA_by_domegaA(o[0], omegamod, delta_t)
    J[1, 1] = dqA_by_domegaB(o[0], o[1], omegamod, delta_t)
    J[1, 2] = dqA_by_domegaB(o[0], o[2], omegamod, delta_t)
    J[2, 0] = dqA_by_domegaB(o[1], o[0], omegamod, delta_t)
    J[2, 1] = dqA_by_domegaA(o[1], omegamod, delta_t)
    J[2, 2] = dqA_by_domegaB(o[1], o[2], omegamod, delta_t)
    J[3, 0] = dqA_by_domegaB(o[2], o[0], omegamod, delta_t)
    J[3, 1] = dqA_by_domegaB(o[2], o[1], omegamod, delta_t)
    J[3, 2] = dqA_by_domegaA(o[2], omegamod, delta_t)
    return J


def dRq_times_a_by_dq(q, a):
    """mc/dRq_times_a_by_dq.m:1-77."""
    q0, qx, qy, qz = q[0], q[1], q[2], q[3]
    dR0 = np.array([[2 * q0, -2 * qz, 2 * qy], [2 * qz, 2 * q0, -2 * qx], [-2 * qy, 2 * qx, 2 * q0]])
    dRx = np.array([[2 * qx, 2 * qy, 2 * qz], [2 * qy, -2 * qx, -2 * q0], [2 * qz, 2 * q0, -2 * qx]])
    dRy = np.array([[-2 * qy, 2 * qx, 2 * q0], [2 * qx, 2 * qy, 2 * qz], [-2 * q0, 2 * qz, -2 * qy]])
    dRz = np.array([[-2 * qz, -2 * q0, 2 * qx], [2 * q0, -2 * qz, 2 * qy], [2 * qx, 2 * qy, 2 * qz]])
    out = np.zeros((3, 4))
    out[:, 0] = dR0 @ a
    out[:, 1] = dRx @ a
    out[:, 2] = dRy @ a
    out[:, 3] = dRz @ a
    return out


def dqbar_by_dq():
    """mc/dqbar_by_dq.m:3."""
    return np.diag([1.0, -1.0, -1.0, -1.0])


def normJac(q):
    """mc/normJac.m:3-12."""
    r, x, y, z = q[0], q[1], q[2], q[3]
    return (r * r + x * x + y * y + z * z) ** (-3.0 / 2.0) * np.array([
        [x * x + y * y + z * z, -r * x, -r * y, -r * z],
        [-x * r, r * r + y * y + z * z, -x * y, -x * z],
        [-y * r, -y * x, r * r + x * x + z * z, -y * z],
        [-z * r, -z * x, -z * y, r * r + x * x + y * y]])


# --------------------------------------------------------------------------
# L2  camera model
# --------------------------------------------------------------------------
def m(theta, phi):
    """mc/m.m:12-14 (two-argument form) and :6-8 (vectorised form when theta, phi are arrays)."""
    cphi = np.cos(phi)
    return np.array([cphi * np.sin(theta), -np.sin(phi), cphi * np.cos(theta)])


def hu(yi, cam):
    """mc/hu.m:3-13.  yi is 3 or 3xK."""
    u0, v0, f = cam.Cx, cam.Cy, cam.f
    ku, kv = 1 / cam.dx, 1 / cam.dy
    yi = np.asarray(yi, dtype=np.float64)
    return np.array([u0 + (yi[0] / yi[2]) * f * ku,
                     v0 + (yi[1] / yi[2]) * f * kv])


def distort_fm(uv, cam):
    """mc/distort_fm.m:14-38.  uv is 2 or 2xK."""
    Cx, Cy, k1, k2, dx, dy = cam.Cx, cam.Cy, cam.k1, cam.k2, cam.dx, cam.dy
    uv = np.asarray(uv, dtype=np.float64)
    xu = (uv[0] - Cx) * dx
    yu = (uv[1] - Cy) * dy
    ru = np.sqrt(xu * xu + yu * yu)
    rd = ru / (1 + k1 * ru ** 2 + k2 * ru ** 4)
    for _ in range(10):
        f = rd + k1 * rd ** 3 + k2 * rd ** 5 - ru
        f_p = 1 + 3 * k1 * rd ** 2 + 5 * k2 * rd ** 4
        rd = rd - f / f_p
    D = 1 + k1 * rd ** 2 + k2 * rd ** 4
    xd = xu / D
    yd = yu / D
    return np.array([xd / dx + Cx, yd / dy + Cy])


def undistort_fm(uvd, cam):
    """mc/undistort_fm.m:11-27."""
    Cx, Cy, k1, k2, dx, dy = cam.Cx, cam.Cy, cam.k1, cam.k2, cam.dx, cam.dy
    uvd = np.asarray(uvd, dtype=np.float64)
    xd = (uvd[0] - Cx) * dx
    yd = (uvd[1] - Cy) * dy
    rd = np.sqrt(xd * xd + yd * yd)
    D = 1 + k1 * rd ** 2 + k2 * rd ** 4
    xu = xd * D
    yu = yd * D
    return np.array([xu / dx + Cx, yu / dy + Cy])


def jacob_undistor_fm(cam, uvd):
    """mc/jacob_undistor_fm.m:13-34."""
    Cx, Cy, k1, k2, dx, dy = cam.Cx, cam.Cy, cam.k1, cam.k2, cam.dx, cam.dy
    ud, vd = float(uvd[0]), float(uvd[1])
    xd = (ud - Cx) * dx
    yd = (vd - Cy) * dy
    rd2 = xd * xd + yd * yd
    rd4 = rd2 * rd2
    uu_ud = (1 + k1 * rd2 + k2 * rd4) + (ud - Cx) * (k1 + 2 * k2 * rd2) * (2 * (ud - Cx) * dx * dx)
    vu_vd = (1 + k1 * rd2 + k2 * rd4) + (vd - Cy) * (k1 + 2 * k2 * rd2) * (2 * (vd - Cy) * dy * dy)
    uu_vd = (ud - Cx) * (k1 + 2 * k2 * rd2) * (2 * (vd - Cy) * dy * dy)
    vu_ud = (vd - Cy) * (k1 + 2 * k2 * rd2) * (2 * (ud - Cx) * dx * dx)
    return np.array([[uu_ud, uu_vd], [vu_ud, vu_vd]])


# --------------------------------------------------------------------------
# L3  prediction
# --------------------------------------------------------------------------
def fv(X_k_k, delta_t, type_, std_a=None, std_alpha=None):
    """mc/fv.m:3-47."""
    rW, qWR, vW, wW = X_k_k[0:3], X_k_k[3:7], X_k_k[7:10], X_k_k[10:13]
    if type_ == "constant_velocity":  # :42-47
        return np.concatenate([rW + vW * delta_t, qprod(qWR, v2q(wW * delta_t)), vW, wW])
    if type_ == "constant_orientation":  # :8-14
        return np.concatenate([rW + vW * delta_t, qWR, vW, np.zeros(3)])
    if type_ == "constant_position":  # :16-22
        return np.concatenate([rW, qprod(qWR, v2q(wW * delta_t)), np.zeros(3), wW])
    if type_ in ("constant_position_and_orientation",
                 "constant_position_and_orientation_location_noise"):  # :24-40
        return np.concatenate([rW, qWR, np.zeros(3), np.zeros(3)])
    raise ValueError(type_)


def dfv_by_dxv(Xv, u, dt, type_):
    """mc/dfv_by_dxv.m:3-31."""
    omegaOld = Xv[10:13]
    qOld = Xv[3:7]
    F = np.eye(13)
    qwt = v2q(omegaOld * dt)
    F[3:7, 3:7] = dq3_by_dq2(qwt)
    if type_ == "constant_velocity":
        F[0:3, 7:10] = np.eye(3) * dt
        F[3:7, 10:13] = dq3_by_dq1(qOld) @ dqomegadt_by_domega(omegaOld, dt)
    if type_ == "constant_orientation":
        F[3:7, 10:13] = 0
        F[10:13, 10:13] = 0
    if type_ == "constant_position":
        F[0:3, 7:10] = 0
        F[7:10, 7:10] = 0
    if type_ == "constant_position_and_orientation":
        F[3:7, 10:13] = 0
        F[0:3, 7:10] = 0
        F[10:13, 10:13] = 0
        F[7:10, 7:10] = 0
    return F


def func_Q(Xv, u, Pn, delta_t, type_):
    """mc/func_Q.m:13-28 (the '..._location_noise' branch :3-11 needs functions the
    hot path never reaches and is not restated)."""
    if type_ == "constant_position_and_orientation_location_noise":
        raise NotImplementedError("mc/func_Q.m:3-11 is off the hot path")
    omegaOld = Xv[10:13]
    qOld = Xv[3:7]
    G = np.zeros((13, 6))
    G[7:10, 0:3] = np.eye(3)
    G[10:13, 3:6] = np.eye(3)
    G[0:3, 0:3] = np.eye(3) * delta_t
    G[3:7, 3:6] = dq3_by_dq1(qOld) @ dqomegadt_by_domega(omegaOld, delta_t)
    return G @ Pn @ G.T


def predict_state_and_covariance(X_k, P_k, type_, SD_A, SD_alpha):
    """mc/predict_state_and_covariance.m:3-27."""
    X_k = np.asarray(X_k, dtype=np.float64)
    P_k = np.asarray(P_k, dtype=np.float64)
    delta_t = 1
    Xv = fv(X_k[0:13], delta_t, type_, SD_A, SD_alpha)
    X_km1_k = np.concatenate([Xv, X_k[13:]])
    F = dfv_by_dxv(X_k[0:13], np.zeros(6), delta_t, type_)
    la = (SD_A * delta_t) ** 2
    aa = (SD_alpha * delta_t) ** 2
    Pn = np.diag([la, la, la, aa, aa, aa])
    Q = func_Q(X_k[0:13], np.zeros(6), Pn, delta_t, type_)
    n = P_k.shape[0]
    P = np.empty((n, n))
    P[0:13, 0:13] = F @ P_k[0:13, 0:13] @ F.T + Q
    P[0:13, 13:] = F @ P_k[0:13, 13:]
    P[13:, 0:13] = P_k[13:, 0:13] @ F.T
    P[13:, 13:] = P_k[13:, 13:]
    return X_km1_k, P


def ekf_prediction(f, features_info):
    """mc/ekf_prediction.m:3."""
    f.x_k_km1, f.p_k_km1 = predict_state_and_covariance(f.x_k_k, f.p_k_k, f.type, f.std_a, f.std_alpha)
    return f, features_info


# --------------------------------------------------------------------------
# L3  measurement prediction
# --------------------------------------------------------------------------
def _fov_reject(hrl):
    """mc/hi_inverse_depth.m:37-40 == mc/hi_cartesian.m:11-14."""
    ax = math.atan2(hrl[0], hrl[2]) * 180 / math.pi
    ay = math.atan2(hrl[1], hrl[2]) * 180 / math.pi
    return (ax < -60) or (ax > 60) or (ay < -60) or (ay > 60)


def _in_image(uv_d, cam):
    """mc/hi_inverse_depth.m:51."""
    return (uv_d[0] > 0) and (uv_d[0] < cam.nCols) and (uv_d[1] > 0) and (uv_d[1] < cam.nRows)


def hi_inverse_depth(yinit, t_wc, r_wc, cam, features_info=None):
    """mc/hi_inverse_depth.m:7-57.  Returns a 2-vector or None ('[]')."""
    r_cw = r_wc.T
    yi = yinit[0:3]
    theta, phi, rho = yinit[3], yinit[4], yinit[5]
    mi = m(theta, phi)
    hrl = r_cw @ ((yi - t_wc) * rho + mi)
    if _fov_reject(hrl):
        return None
    with np.errstate(divide="ignore", invalid="ignore"):
        uv_u = hu(hrl, cam)
        uv_d = distort_fm(uv_u, cam)
    if _in_image(uv_d, cam):
        return uv_d
    return None


def hi_cartesian(yi, t_wc, r_wc, cam, features_info=None):
    """mc/hi_cartesian.m:7-49."""
    r_cw = np.linalg.inv(r_wc)
    hrl = r_cw @ (yi - t_wc)
    if _fov_reject(hrl):
        return None
    with np.errstate(divide="ignore", invalid="ignore"):
        uv_u = hu(hrl, cam)
        uv_d = distort_fm(uv_u, cam)
    if _in_image(uv_d, cam):
        return uv_d
    return None


def predict_camera_measurements(x_k_k, cam, features_info):
    """mc/predict_camera_measurements.m:4-28.  h is only written when the feature
    is visible; otherwise the previous value is left in place (:14-16, :23-25)."""
    t_wc = x_k_k[0:3]
    r_wc = q2r(x_k_k[3:7])
    pos = 13
    for fi in features_info:
        if fi.type == "cartesian":
            yi = x_k_k[pos:pos + 3]
            pos += 3
            hi = hi_cartesian(yi, t_wc, r_wc, cam, fi)
            if hi is not None:
                fi.h = hi.copy()
        if fi.type == "inversedepth":
            yi = x_k_k[pos:pos + 6]
            pos += 6
            hi = hi_inverse_depth(yi, t_wc, r_wc, cam, fi)
            if hi is not None:
                fi.h = hi.copy()
    return features_info


def _dhd_dhu(cam, zi_d):
    """mc/calculate_Hi_inverse_depth.m:123-126."""
    return np.linalg.inv(jacob_undistor_fm(cam, zi_d))


def _state_size_and_index(features_info, i):
    """mc/calculate_Hi_inverse_depth.m:5-22 (i is 0-based here; returns 0-based index)."""
    n_c = sum(1 for f in features_info if f.type[0] == "c")
    n_i = sum(1 for f in features_info if f.type[0] == "i")
    n_c_before = sum(1 for f in features_info[:i] if f.type[0] == "c")
    n_i_before = sum(1 for f in features_info[:i] if f.type[0] == "i")
    return 13 + 3 * n_c + 6 * n_i, 13 + 3 * n_c_before + 6 * n_i_before


def calculate_Hi_inverse_depth(Xv, yi, cam, i, features_info):
    """mc/calculate_Hi_inverse_depth.m:1-165.  i is 0-based."""
    zi = features_info[i].h
    n, idx = _state_size_and_index(features_info, i)
    Hi = np.zeros((2, n))
    f, ku, kv = cam.f, 1 / cam.dx, 1 / cam.dy
    rw = Xv[0:3]
    qwr = Xv[3:7]
    theta, phi, rho = yi[3], yi[4], yi[5]

    def dhu_dhrl():  # :138-156
        Rrw = np.linalg.inv(q2r(qwr))
        mi = np.array([math.cos(phi) * math.sin(theta), -math.sin(phi), math.cos(phi) * math.cos(theta)])
        hc = Rrw @ ((yi[0:3] - rw) * rho + mi)
        hcx, hcy, hcz = hc[0], hc[1], hc[2]
        return np.array([[+f * ku / hcz, 0.0, -hcx * f * ku / (hcz ** 2)],
                         [0.0, +f * kv / hcz, -hcy * f * kv / (hcz ** 2)]])

    def dh_dhrl():  # :115-117
        return _dhd_dhu(cam, zi) @ dhu_dhrl()

    def dhrl_drw():  # :106-109
        return -(np.linalg.inv(q2r(qwr))) * yi[5]

    def dhrl_dqwr():  # :83-92
        mi = np.array([math.cos(phi) * math.sin(theta), -math.sin(phi), math.cos(phi) * math.cos(theta)])
        return dRq_times_a_by_dq(qconj(qwr), ((yi[0:3] - rw) * rho + mi)) @ dqbar_by_dq()

    def dhrl_dy():  # :43-54
        Rrw = np.linalg.inv(q2r(qwr))
        dmi_dthetai = Rrw @ np.array([math.cos(phi) * math.cos(theta), 0.0, -math.cos(phi) * math.sin(theta)])
        dmi_dphii = Rrw @ np.array([-math.sin(phi) * math.sin(theta), -math.cos(phi), -math.sin(phi) * math.cos(theta)])
        a = np.zeros((3, 6))
        a[:, 0:3] = rho * Rrw
        a[:, 3] = dmi_dthetai
        a[:, 4] = dmi_dphii
        a[:, 5] = Rrw @ (yi[0:3] - rw)
        return a

    Hi[:, 0:3] = dh_dhrl() @ dhrl_drw()      # :98-100
    Hi[:, 3:7] = dh_dhrl() @ dhrl_dqwr()     # :75-77
    Hi[:, idx:idx + 6] = dh_dhrl() @ dhrl_dy()  # :29-31, :23
    return Hi


def calculate_Hi_cartesian(Xv, yi, cam, i, features_info):
    """mc/calculate_Hi_cartesian.m:1-115.  i is 0-based."""
    zi = features_info[i].h
    n, idx = _state_size_and_index(features_info, i)
    Hi = np.zeros((2, n))
    f, ku, kv = cam.f, 1 / cam.dx, 1 / cam.dy
    rw = Xv[0:3]
    qwr = Xv[3:7]

    def dhu_dhrl():  # :101-113
        Rrw = np.linalg.inv(q2r(qwr))
        hrl = Rrw @ (yi - rw)
        return np.array([[f * ku / hrl[2], 0.0, -hrl[0] * f * ku / (hrl[2] ** 2)],
                         [0.0, f * kv / hrl[2], -hrl[1] * f * kv / (hrl[2] ** 2)]])

    def dh_dhrl():  # :85-87
        return _dhd_dhu(cam, zi) @ dhu_dhrl()

    Hi[:, 0:3] = dh_dhrl() @ (-(np.linalg.inv(q2r(qwr))))                                   # :69-79
    Hi[:, 3:7] = dh_dhrl() @ (dRq_times_a_by_dq(qconj(qwr), (yi - rw)) @ dqbar_by_dq())    # :53-63
    Hi[:, idx:idx + 3] = dh_dhrl() @ np.linalg.inv(q2r(qwr))                                # :29-39
    return Hi


def calculate_derivatives(x_k_km1, cam, features_info):
    """mc/calculate_derivatives.m:3-28."""
    x_v = x_k_km1[0:13]
    pos = 13
    for i, fi in enumerate(features_info):
        if fi.h is not None:
            if fi.type == "cartesian":
                y = x_k_km1[pos:pos + 3]
                pos += 3
                fi.H = calculate_Hi_cartesian(x_v, y, cam, i, features_info)
            else:
                y = x_k_km1[pos:pos + 6]
                pos += 6
                fi.H = calculate_Hi_inverse_depth(x_v, y, cam, i, features_info)
        else:
            if fi.type == "cartesian":
                pos += 3
            if fi.type == "inversedepth":
                pos += 6
    return features_info


def predict_and_derive(filter, features_info, cam):
    """mc/search_IC_matches.m:4-10 — everything of search_IC_matches except the
    image matcher (:17), which is out of scope and replaced by a supplied matcher."""
    features_info = predict_camera_measurements(filter.x_k_km1, cam, features_info)
    features_info = calculate_derivatives(filter.x_k_km1, cam, features_info)
    for fi in features_info:
        if fi.h is not None:
            fi.S = fi.H @ filter.p_k_km1 @ fi.H.T + fi.R
    return features_info


def gate_candidates(features_info, z_cand, has_cand):
    """Synthetic stand-in for mc/matching.m: keeps only the matcher's GATING rule.
    A feature with a predicted h and a candidate pixel becomes individually
    compatible iff all(eig(S) < 100) (:16) and nu' inv(S) nu < 5.9915 (:38);
    z and individually_compatible are then written together (:52-53)."""
    for i, fi in enumerate(features_info):
        if fi.h is None or not has_cand[i]:
            continue
        S = fi.S
        if not np.all(np.linalg.eigvals(S).real < 100):
            continue
        nu = np.array([z_cand[i][0] - fi.h[0], z_cand[i][1] - fi.h[1]])
        if nu @ np.linalg.inv(S) @ nu < CHI2INV_2_95:
            fi.individually_compatible = 1
            fi.z = np.array([z_cand[i][0], z_cand[i][1]], dtype=np.float64)
    return features_info


def search_IC_matches(filter, features_info, cam, im):
    """mc/search_IC_matches.m:1-17 with ``im`` = (z_cand [N,2], has_cand [N]) consumed
    by the synthetic matcher above."""
    features_info = predict_and_derive(filter, features_info, cam)
    z_cand, has_cand = im
    return gate_candidates(features_info, z_cand, has_cand)


# --------------------------------------------------------------------------
# L3  1-point RANSAC
# --------------------------------------------------------------------------
def generate_state_vector_pattern(features_info, x):
    """mc/generate_state_vector_pattern.m:3-27."""
    pattern = np.zeros((len(x), 4))
    position = 13
    z_id, z_euc = [], []
    for fi in features_info:
        if fi.type == "inversedepth":
            if fi.z is not None:
                pattern[position:position + 3, 0] = 1
                pattern[position + 3:position + 5, 1] = 1
                pattern[position + 5, 2] = 1
                z_id.append([fi.z[0], fi.z[1]])
            position += 6
        if fi.type == "cartesian":
            if fi.z is not None:
                pattern[position:position + 3, 3] = 1
                z_euc.append([fi.z[0], fi.z[1]])
            position += 3
    z_id = np.array(z_id, dtype=np.float64).T if z_id else np.zeros((2, 0))
    z_euc = np.array(z_euc, dtype=np.float64).T if z_euc else np.zeros((2, 0))
    return pattern, z_id, z_euc


def select_random_match(features_info, u):
    """mc/select_random_match.m:3-20 with rand(1) replaced by the supplied uniform u."""
    ic = [i for i, fi in enumerate(features_info) if fi.individually_compatible]
    num_IC_matches = len(ic)
    random_match_position = int(math.floor(u * num_IC_matches))  # (+1 in 1-based MATLAB)
    position = ic[random_match_position]
    return features_info[position].z, position, num_IC_matches


def compute_hypothesis_support_fast(xi, cam, pattern, z_id, z_euc, threshold):
    """mc/compute_hypothesis_support_fast.m:3-90."""
    support = 0
    u0, v0, f = cam.Cx, cam.Cy, cam.f
    ku, kv = 1 / cam.dx, 1 / cam.dy
    if z_id.shape[1] > 0:
        n_id = z_id.shape[1]
        ri = xi[pattern[:, 0].astype(bool)].reshape(n_id, 3).T
        anglesi = xi[pattern[:, 1].astype(bool)].reshape(n_id, 2).T
        rhoi = xi[pattern[:, 2].astype(bool)]
        mi = m(anglesi[0], anglesi[1])
        rwc = xi[0:3].reshape(3, 1)
        rotcw = q2r(xi[3:7]).T
        ri_minus_rwc = ri - rwc
        ri_minus_rwc_by_rhoi = ri_minus_rwc * rhoi
        hc = rotcw @ (ri_minus_rwc_by_rhoi + mi)
        with np.errstate(divide="ignore", invalid="ignore"):
            h_norm = np.array([hc[0] / hc[2], hc[1] / hc[2]])
            h_image = f * ku * h_norm + np.array([[u0], [v0]])
            h_distorted = distort_fm(h_image, cam)
            nu = z_id - h_distorted
            residuals = np.sqrt(nu[0] ** 2 + nu[1] ** 2)
            pos_id = residuals < threshold
        support += int(pos_id.sum())
    else:
        pos_id = np.zeros(0, dtype=bool)
    if z_euc.shape[1] > 0:
        n_euc = z_euc.shape[1]
        xyz = xi[pattern[:, 3].astype(bool)].reshape(n_euc, 3).T
        rwc = xi[0:3].reshape(3, 1)
        rotcw = q2r(xi[3:7]).T
        hc = rotcw @ (xyz - rwc)
        with np.errstate(divide="ignore", invalid="ignore"):
            h_norm = np.array([hc[0] / hc[2], hc[1] / hc[2]])
            h_image = f * ku * h_norm + np.array([[u0], [v0]])
            h_distorted = distort_fm(h_image, cam)
            nu = z_euc - h_distorted
            residuals = np.sqrt(nu[0] ** 2 + nu[1] ** 2)
            pos_euc = residuals < threshold
        support += int(pos_euc.sum())
    else:
        pos_euc = np.zeros(0, dtype=bool)
    return support, pos_id, pos_euc


def set_as_most_supported_hypothesis(features_info, pos_id, pos_euc):
    """mc/set_as_most_supported_hypothesis.m:3-27."""
    j_id = 0
    j_euc = 0
    for fi in features_info:
        if fi.z is not None:
            if fi.type == "cartesian":
                fi.low_innovation_inlier = 1 if pos_euc[j_euc] else 0
                j_euc += 1
            if fi.type == "inversedepth":
                fi.low_innovation_inlier = 1 if pos_id[j_id] else 0
                j_id += 1
    return features_info


def n_hyp_rule(hypothesis_support, num_IC_matches, p_at_least_one_spurious_free=0.99):
    """mc/ransac_hypotheses.m:40-41, evaluated with the C library's log (math.log)."""
    epsilon = 1 - (hypothesis_support / num_IC_matches)
    den = 1 - (1 - epsilon)
    num = math.log(1 - p_at_least_one_spurious_free)
    if den <= 0.0:
        lden = -math.inf  # MATLAB log(0) = -Inf
    else:
        lden = math.log(den)
    if lden == 0.0:
        return math.inf if num < 0 else -math.inf  # cannot happen for support >= 1
    return math.ceil(num / lden)  # ceil(-0.0) == 0 when lden == -inf


def ransac_hypotheses(filter, features_info, cam, u, fixed_hypotheses=0, info=None):
    """mc/ransac_hypotheses.m:3-47.

    ``u`` is the uniform stream: iteration i (1-based) consumes u[i-1] where the
    reference calls rand(1) (mc/select_random_match.m:12).  ``fixed_hypotheses`` > 0
    is the fixed-budget mode of BASELINE configs 2/5 (exactly that many iterations,
    lines :41-42 and :45 disabled) — an extension, not reference behaviour.  With
    zero individually compatible matches the reference crashes
    (mc/select_random_match.m:16); here that frame is a no-op."""
    threshold = filter.std_z
    n_hyp = 1000
    max_hypothesis_support = 0
    iters = 0
    if not any(fi.individually_compatible for fi in features_info):
        if info is not None:
            info["iterations"] = 0
        return features_info
    pattern, z_id, z_euc = generate_state_vector_pattern(features_info, filter.x_k_km1)
    n_loop = fixed_hypotheses if fixed_hypotheses > 0 else n_hyp
    for i in range(1, n_loop + 1):
        if i - 1 >= len(u):
            raise RuntimeError("uniform stream exhausted at iteration %d" % i)
        zi, position, num_IC_matches = select_random_match(features_info, u[i - 1])
        iters = i
        x_k_km1 = filter.x_k_km1
        p_k_km1 = filter.p_k_km1
        hi = features_info[position].h
        Hi = features_info[position].H
        S = Hi @ p_k_km1 @ Hi.T + features_info[position].R
        K = p_k_km1 @ Hi.T @ np.linalg.inv(S)
        xi = x_k_km1 + K @ (zi - hi)
        support, pos_id, pos_euc = compute_hypothesis_support_fast(xi, cam, pattern, z_id, z_euc, threshold)
        if support > max_hypothesis_support:
            max_hypothesis_support = support
            features_info = set_as_most_supported_hypothesis(features_info, pos_id, pos_euc)
            if fixed_hypotheses <= 0:
                n_hyp = n_hyp_rule(support, num_IC_matches)
                if n_hyp == 0:
                    break
        if fixed_hypotheses <= 0 and i > n_hyp:
            break
    if info is not None:
        info["iterations"] = iters
        info["max_support"] = max_hypothesis_support
    return features_info


# --------------------------------------------------------------------------
# L3  updates
# --------------------------------------------------------------------------
def update(x_km1_k, p_km1_k, H, R, z, h):
    """mc/update.m:3-32."""
    x_km1_k = np.asarray(x_km1_k, dtype=np.float64)
    p_km1_k = np.asarray(p_km1_k, dtype=np.float64)
    if z is not None and len(z) > 0:
        S = H @ p_km1_k @ H.T + R
        K = p_km1_k @ H.T @ np.linalg.inv(S)
        x_k_k = x_km1_k + K @ (z - h)
        p_k_k = p_km1_k - K @ S @ K.T
        p_k_k = 0.5 * p_k_k + 0.5 * p_k_k.T
        Jnorm = normJac(x_k_k[3:7])
        out = p_k_k.copy()
        out[0:3, 3:7] = p_k_k[0:3, 3:7] @ Jnorm.T
        out[3:7, 0:3] = Jnorm @ p_k_k[3:7, 0:3]
        out[3:7, 3:7] = Jnorm @ p_k_k[3:7, 3:7] @ Jnorm.T
        out[3:7, 7:] = Jnorm @ p_k_k[3:7, 7:]
        out[7:, 3:7] = p_k_k[7:, 3:7] @ Jnorm.T
        p_k_k = out
        x_k_k = x_k_k.copy()
        x_k_k[3:7] = x_k_k[3:7] / np.linalg.norm(x_k_k[3:7])
        return x_k_k, p_k_k, K
    return x_km1_k.copy(), p_km1_k.copy(), 0


def _stack(features_info, flag):
    """mc/ekf_update_li_inliers.m:4-18 / mc/ekf_update_hi_inliers.m:4-18."""
    z, h, H = [], [], []
    for fi in features_info:
        if getattr(fi, flag) == 1:
            z += [fi.z[0], fi.z[1]]
            h += [fi.h[0], fi.h[1]]
            H.append(fi.H)
    if not z:
        return None, None, None, None
    z = np.array(z, dtype=np.float64)
    h = np.array(h, dtype=np.float64)
    H = np.vstack(H)
    return z, h, H, np.eye(len(z))


def ekf_update_li_inliers(filter, features_info):
    """mc/ekf_update_li_inliers.m:4-21."""
    z, h, H, R = _stack(features_info, "low_innovation_inlier")
    filter.x_k_k, filter.p_k_k, _ = update(filter.x_k_km1, filter.p_k_km1, H, R, z, h)
    return filter


def rescue_hi_inliers(filter, features_info, cam):
    """mc/rescue_hi_inliers.m:3-22."""
    features_info = predict_camera_measurements(filter.x_k_k, cam, features_info)
    features_info = calculate_derivatives(filter.x_k_k, cam, features_info)
    for fi in features_info:
        if fi.individually_compatible == 1 and fi.low_innovation_inlier == 0:
            hi = fi.h
            Si = fi.H @ filter.p_k_k @ fi.H.T
            nui = fi.z - hi
            if nui @ np.linalg.inv(Si) @ nui < CHI2INV_2_95:
                fi.high_innovation_inlier = 1
            else:
                fi.high_innovation_inlier = 0
    return features_info


def ekf_update_hi_inliers(filter, features_info):
    """mc/ekf_update_hi_inliers.m:4-21."""
    z, h, H, R = _stack(features_info, "high_innovation_inlier")
    filter.x_k_k, filter.p_k_k, _ = update(filter.x_k_k, filter.p_k_k, H, R, z, h)
    return filter


def filter_step(filter, features_info, cam, im, u, fixed_hypotheses=0, info=None):
    """One 'filter-step' = mc/mono_slam.m:56-74 without takeImage (:59); the matcher
    inside search_IC_matches is the synthetic gate."""
    filter, features_info = ekf_prediction(filter, features_info)
    features_info = search_IC_matches(filter, features_info, cam, im)
    features_info = ransac_hypotheses(filter, features_info, cam, u, fixed_hypotheses, info)
    filter = ekf_update_li_inliers(filter, features_info)
    features_info = rescue_hi_inliers(filter, features_info, cam)
    filter = ekf_update_hi_inliers(filter, features_info)
    return filter, features_info


# --------------------------------------------------------------------------
# L4 pieces needed to build maps ("next" rows of SURVEY §8f, pinned by the fixture)
# --------------------------------------------------------------------------
def hinv(uvd, Xv, cam, initial_rho):
    """mc/hinv.m:3-26."""
    fku, fkv, U0, V0 = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]
    uv = undistort_fm(uvd, cam)
    u, v = uv[0], uv[1]
    r_W = Xv[0:3]
    q_WR = Xv[3:7]
    h_LR = np.array([-(U0 - u) / fku, -(V0 - v) / fkv, 1.0])
    n = q2r(q_WR) @ h_LR
    nx, ny, nz = n[0], n[1], n[2]
    return np.array([r_W[0], r_W[1], r_W[2], math.atan2(nx, nz),
                     math.atan2(-ny, math.sqrt(nx * nx + nz * nz)), initial_rho])


def add_a_feature_covariance_inverse_depth(P, uvd, Xv, std_pxl, std_rho, cam):
    """mc/add_a_feature_covariance_inverse_depth.m:3-64."""
    fku, fkv, U0, V0 = cam.K[0, 0], cam.K[1, 1], cam.K[0, 2], cam.K[1, 2]
    q_wc = Xv[3:7]
    R_wc = q2r(q_wc)
    uvu = undistort_fm(uvd, cam)
    uu, vu = uvu[0], uvu[1]
    XYZ_c = np.array([-(U0 - uu) / fku, -(V0 - vu) / fkv, 1.0])
    XYZ_w = R_wc @ XYZ_c
    X_w, Y_w, Z_w = XYZ_w[0], XYZ_w[1], XYZ_w[2]
    dtheta_dgw = np.array([Z_w / (X_w ** 2 + Z_w ** 2), 0.0, -X_w / (X_w ** 2 + Z_w ** 2)])
    dphi_dgw = np.array([
        (X_w * Y_w) / ((X_w ** 2 + Y_w ** 2 + Z_w ** 2) * math.sqrt(X_w ** 2 + Z_w ** 2)),
        -math.sqrt(X_w ** 2 + Z_w ** 2) / (X_w ** 2 + Y_w ** 2 + Z_w ** 2),
        (Z_w * Y_w) / ((X_w ** 2 + Y_w ** 2 + Z_w ** 2) * math.sqrt(X_w ** 2 + Z_w ** 2))])
    dgw_dqwr = dRq_times_a_by_dq(q_wc, XYZ_c)
    dtheta_dqwr = dtheta_dgw @ dgw_dqwr
    dphi_dqwr = dphi_dgw @ dgw_dqwr
    dy_dqwr = np.vstack([np.zeros((3, 4)), dtheta_dqwr, dphi_dqwr, np.zeros((1, 4))])
    dy_drw = np.vstack([np.eye(3), np.zeros((3, 3))])
    dy_dxv = np.hstack([dy_drw, dy_dqwr, np.zeros((6, 6))])
    dyprima_dgw = np.vstack([np.zeros((3, 3)), dtheta_dgw, dphi_dgw])
    dgw_dgc = R_wc
    dgc_dhu = np.array([[1 / fku, 0.0, 0.0], [0.0, 1 / fkv, 0.0]]).T
    dhu_dhd = jacob_undistor_fm(cam, uvd)
    dyprima_dhd = dyprima_dgw @ dgw_dgc @ dgc_dhu @ dhu_dhd
    dy_dhd = np.zeros((6, 3))
    dy_dhd[0:5, 0:2] = dyprima_dhd
    dy_dhd[5, 2] = 1.0
    Padd = np.diag([std_pxl ** 2, std_pxl ** 2, std_rho ** 2])
    n = P.shape[0]
    P_xv = P[0:13, 0:13]
    P_yxv = P[13:, 0:13]
    P_y = P[13:, 13:]
    P_xvy = P[0:13, 13:]
    P_RES = np.zeros((n + 6, n + 6))
    P_RES[0:13, 0:13] = P_xv
    P_RES[0:13, 13:n] = P_xvy
    P_RES[0:13, n:] = P_xv @ dy_dxv.T
    P_RES[13:n, 0:13] = P_yxv
    P_RES[13:n, 13:n] = P_y
    P_RES[13:n, n:] = P_yxv @ dy_dxv.T
    P_RES[n:, 0:13] = dy_dxv @ P_xv
    P_RES[n:, 13:n] = dy_dxv @ P_xvy
    P_RES[n:, n:] = dy_dxv @ P_xv @ dy_dxv.T + dy_dhd @ Padd @ dy_dhd.T
    return P_RES


def add_features_inverse_depth(uvd, X, P, cam, std_pxl, initial_rho, std_rho):
    """mc/add_features_inverse_depth.m:3-24 for a single new feature (uvd is a 2-vector)."""
    Xv = X[0:13]
    newFeature = hinv(uvd, Xv, cam, initial_rho)
    X_RES = np.concatenate([X, newFeature])
    P_RES = add_a_feature_covariance_inverse_depth(P, uvd, Xv, std_pxl, std_rho, cam)
    return X_RES, P_RES, newFeature


def inversedepth2cartesian(inverse_depth):
    """mc/inversedepth2cartesian.m:3-12."""
    rw = inverse_depth[0:3]
    theta, phi, rho = inverse_depth[3], inverse_depth[4], inverse_depth[5]
    cphi = math.cos(phi)
    mm = np.array([cphi * math.sin(theta), -math.sin(phi), cphi * math.cos(theta)])
    return np.array([rw[0] + (1.0 / rho) * mm[0], rw[1] + (1.0 / rho) * mm[1], rw[2] + (1.0 / rho) * mm[2]])


def convert_feature_to_cartesian(X, P, features_info, i):
    """mc/inversedepth_2_cartesian.m:35-48: unconditional conversion of feature i
    (0-based) — the state/covariance transformation without the linearity test."""
    pos = 13
    for j in range(i):
        pos += 3 if features_info[j].type == "cartesian" else 6
    rho = X[pos + 5]
    theta, phi = X[pos + 3], X[pos + 4]
    mi = m(theta, phi)
    p = inversedepth2cartesian(X[pos:pos + 6])
    n_old = X.shape[0]
    Xn = np.concatenate([X[:pos], p, X[pos + 6:]])
    dm_dtheta = np.array([math.cos(phi) * math.cos(theta), 0.0, -math.cos(phi) * math.sin(theta)])
    dm_dphi = np.array([-math.sin(phi) * math.sin(theta), -math.cos(phi), -math.sin(phi) * math.cos(theta)])
    J = np.zeros((3, 6))
    J[:, 0:3] = np.eye(3)
    J[:, 3] = (1 / rho) * dm_dtheta
    J[:, 4] = (1 / rho) * dm_dphi
    J[:, 5] = -mi / (rho ** 2)
    J_all = np.zeros((n_old - 3, n_old))
    J_all[:pos, :pos] = np.eye(pos)
    J_all[pos:pos + 3, pos:pos + 6] = J
    J_all[pos + 3:, pos + 6:] = np.eye(n_old - pos - 6)
    Pn = J_all @ P @ J_all.T
    features_info[i].type = "cartesian"
    return Xn, Pn


def linearity_index(X, P, features_info, i):
    """mc/inversedepth_2_cartesian.m:12-32 for feature i (0-based, inverse-depth)."""
    pos = 13
    for j in range(i):
        pos += 3 if features_info[j].type == "cartesian" else 6
    std_rho = math.sqrt(P[pos + 5, pos + 5])
    rho = X[pos + 5]
    std_d = std_rho / (rho ** 2)
    x_c1 = X[pos:pos + 3]
    x_c2 = X[0:3]
    p = inversedepth2cartesian(X[pos:pos + 6])
    d_c2p = float(np.linalg.norm(p - x_c2))
    cos_alpha = ((p - x_c1) @ (p - x_c2)) / (np.linalg.norm(p - x_c1) * np.linalg.norm(p - x_c2))
    return 4 * std_d * cos_alpha / d_c2p


def inversedepth_2_cartesian(filter, features_info):
    """mc/inversedepth_2_cartesian.m:3-52 (at most one conversion per call, :49)."""
    X = np.asarray(filter.x_k_k, dtype=np.float64)
    P = np.asarray(filter.p_k_k, dtype=np.float64)
    for i, fi in enumerate(features_info):
        if fi.type == "inversedepth":
            if linearity_index(X, P, features_info, i) < 0.1:
                filter.x_k_k, filter.p_k_k = convert_feature_to_cartesian(X, P, features_info, i)
                return filter, features_info
    return filter, features_info


def delete_a_feature(X, P, featToDelete, features_info):
    """mc/delete_a_feature.m:4-25 (featToDelete 0-based)."""
    par = 3 if features_info[featToDelete].type == "cartesian" else 6
    idx = 13
    for j in range(featToDelete):
        idx += 6 if features_info[j].type == "inversedepth" else 3
    keep = np.r_[0:idx, idx + par:X.shape[0]]
    return X[keep].copy(), P[np.ix_(keep, keep)].copy()


def update_iterated(x_km1_k, p_km1_k, features_info, cam, flag, n_iter=3):
    """EXTENSION — mc/ekf_update_iterated.m:3 calls ``update_iterated`` which does not
    exist anywhere in the reference, so there are no reference semantics to match
    (PARITY UNPINNED by construction).  Standard iterated EKF over the features whose
    ``flag`` attribute is 1: x_{j+1} = x^- + K_j (z - h(x_j) - H_j (x^- - x_j)),
    P from the last linearisation, followed by the quaternion fix-up of mc/update.m:18-24.
    h/H of the selected features are re-evaluated at each iterate with the reference's
    own measurement functions."""
    x0 = np.asarray(x_km1_k, dtype=np.float64)
    P0 = np.asarray(p_km1_k, dtype=np.float64)
    sel = [i for i, fi in enumerate(features_info) if getattr(fi, flag) == 1]
    if not sel:
        return x0.copy(), P0.copy()
    xj = x0.copy()
    for _ in range(n_iter):
        features_info = predict_camera_measurements(xj, cam, features_info)
        features_info = calculate_derivatives(xj, cam, features_info)
        z = np.concatenate([features_info[i].z for i in sel])
        h = np.concatenate([features_info[i].h for i in sel])
        H = np.vstack([features_info[i].H for i in sel])
        S = H @ P0 @ H.T + np.eye(len(z))
        K = P0 @ H.T @ np.linalg.inv(S)
        xj = x0 + K @ (z - h - H @ (x0 - xj))
    p = P0 - K @ S @ K.T
    p = 0.5 * p + 0.5 * p.T
    Jn = normJac(xj[3:7])
    J = np.eye(len(xj))
    J[3:7, 3:7] = Jn
    p = J @ p @ J.T
    xj[3:7] = xj[3:7] / np.linalg.norm(xj[3:7])
    return xj, p


# --------------------------------------------------------------------------
# L4  map management: mc/map_management.m and what it calls.
# Pinned by executing the reference's own map_management.m / inversedepth_2_cartesian.m /
# delete_a_feature.m / add_features_inverse_depth.m / add_feature_to_info_vector.m /
# update_features_info.m through oracle/mref (tests/test_oracle_map_ref.py).
# --------------------------------------------------------------------------
def delete_features(filter, features_info):
    """``delete_features`` is called by mc/map_management.m:7 but is NOT shipped by the reference.
    Rule of the published 1-point-RANSAC toolbox the reference derives from (same statement as the
    harness shim baseline/octave/shims/delete_features.m): a feature predicted more than 5 times and
    matched in fewer than half of its predictions is removed; highest index first."""
    for i in range(len(features_info) - 1, -1, -1):
        fi = features_info[i]
        if fi.times_measured < 0.5 * fi.times_predicted and fi.times_predicted > 5:
            filter.x_k_k, filter.p_k_k = delete_a_feature(filter.x_k_k, filter.p_k_k, i, features_info)
            features_info = features_info[:i] + features_info[i + 1:]
    return filter, features_info


def initialize_features(step, cam, filter, features_info, num_features_to_initialize, im):
    """mc/initialize_features.m:4-21 with the corner search of mc/initialize_a_feature.m:14-57 (CV
    Toolbox) replaced by supplied detections ``im = (uv [K,2], tag [K])``, one attempt per detection;
    the rest is mc/initialize_a_feature.m:60-70: add_features_inverse_depth with
    initial_rho = 1, std_rho = 1, std_pxl = std_z, then add_feature_to_info_vector."""
    uv_list, tags = im
    max_attempts, attempts, initialized = 50, 0, 0
    K = len(uv_list)
    while initialized < num_features_to_initialize and attempts < max_attempts and attempts < K:
        uv = np.asarray(uv_list[attempts], dtype=np.float64)
        X_RES, P_RES, newFeature = add_features_inverse_depth(uv, filter.x_k_k, filter.p_k_k, cam,
                                                              filter.std_z, 1.0, 1.0)
        filter.x_k_k, filter.p_k_k = X_RES, P_RES
        fi = new_feature_info(uv, X_RES, step, newFeature)
        fi.feature_when_initialized = int(tags[attempts])
        features_info = features_info + [fi]
        attempts += 1
        initialized += 1
    return filter, features_info


def map_management(filter, features_info, cam, im, min_number_of_features_in_image, step):
    """mc/map_management.m:4-35: delete -> count measured -> update_features_info -> (at most one)
    inverse-depth -> Cartesian conversion -> top up to min_number_of_features_in_image."""
    filter, features_info = delete_features(filter, features_info)                       # :7
    measured = 0
    for fi in features_info:                                                              # :11-14
        if fi.low_innovation_inlier or fi.high_innovation_inlier:
            measured += 1
    features_info = update_features_info(features_info)                                   # :17
    filter, features_info = inversedepth_2_cartesian(filter, features_info)               # :22
    if measured == 0:                                                                     # :27-35
        filter, features_info = initialize_features(step, cam, filter, features_info,
                                                    min_number_of_features_in_image, im)
    elif measured < min_number_of_features_in_image:
        filter, features_info = initialize_features(step, cam, filter, features_info,
                                                    min_number_of_features_in_image - measured, im)
    return filter, features_info
