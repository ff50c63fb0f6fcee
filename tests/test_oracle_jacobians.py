"""Finite-difference checks of every analytic Jacobian the oracle restates — the reference's own
verification idea (the commented-out fsolve blocks at mc/calculate_Hi_inverse_depth.m:33-38,
64-69, 128-133, 158-163), at random NON-identity poses, which the golden vector (identity pose)
cannot discriminate (sign conventions of dq3_by_dq1, qconj chain)."""
import numpy as np
import pytest

from oracle import ekf_oracle as O


def _rand_pose(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    if q[0] < 0:
        q = -q
    q = np.array([0.95, 0.0, 0.0, 0.0]) + 0.15 * q
    q /= np.linalg.norm(q)
    xv = np.concatenate([rng.uniform(-0.3, 0.3, 3), q, rng.uniform(-0.05, 0.05, 3), rng.uniform(-0.05, 0.05, 3)])
    return xv


def _feature_in_view(rng, xv, cam, kind):
    R = O.q2r(xv[3:7])
    for _ in range(100):
        pc = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0), rng.uniform(3.0, 9.0)])
        pw = R @ pc + xv[0:3]
        if kind == "cartesian":
            y = pw
            ok = O.hi_cartesian(y, xv[0:3], R, cam) is not None
        else:
            anchor = xv[0:3] + rng.uniform(-0.2, 0.2, 3)
            d = pw - anchor
            rho = 1.0 / np.linalg.norm(d)
            theta = np.arctan2(d[0], d[2])
            phi = np.arctan2(-d[1], np.hypot(d[0], d[2]))
            y = np.concatenate([anchor, [theta, phi, rho]])
            ok = O.hi_inverse_depth(y, xv[0:3], R, cam) is not None
        if ok:
            return y
    raise RuntimeError("no visible feature found")


def _h_of(x, cam, kind):
    R = O.q2r(x[3:7])
    if kind == "cartesian":
        return O.hi_cartesian(x[13:16], x[0:3], R, cam)
    return O.hi_inverse_depth(x[13:19], x[0:3], R, cam)


@pytest.mark.parametrize("kind", ["inversedepth", "cartesian"])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_measurement_jacobian_fd(kind, seed):
    rng = np.random.RandomState(seed)
    cam = O.initialize_cam()
    xv = _rand_pose(rng)
    y = _feature_in_view(rng, xv, cam, kind)
    x = np.concatenate([xv, y])
    fi = O.Feature(type=kind, h=None, H=None, z=None)
    feats = O.predict_camera_measurements(x, cam, [fi])
    feats = O.calculate_derivatives(x, cam, feats)
    H = feats[0].H
    Hfd = np.zeros_like(H)
    for j in range(len(x)):
        e = np.zeros(len(x))
        step = 1e-6 * max(1.0, abs(x[j]))
        e[j] = step
        Hfd[:, j] = (_h_of(x + e, cam, kind) - _h_of(x - e, cam, kind)) / (2 * step)
    # columns 8..13 (v, w) are structurally zero
    assert np.all(H[:, 7:13] == 0)
    np.testing.assert_allclose(H, Hfd, rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_state_transition_jacobian_fd(seed):
    rng = np.random.RandomState(10 + seed)
    xv = _rand_pose(rng)
    F = O.dfv_by_dxv(xv, np.zeros(6), 1.0, "constant_velocity")
    Ffd = np.zeros((13, 13))
    for j in range(13):
        e = np.zeros(13)
        e[j] = 1e-7
        Ffd[:, j] = (O.fv(xv + e, 1.0, "constant_velocity") - O.fv(xv - e, 1.0, "constant_velocity")) / 2e-7
    np.testing.assert_allclose(F, Ffd, rtol=1e-6, atol=1e-7)


def test_process_noise_is_G_Pn_Gt():
    rng = np.random.RandomState(5)
    xv = _rand_pose(rng)
    Pn = np.diag([1e-4] * 3 + [4e-4] * 3)
    Q = O.func_Q(xv, np.zeros(6), Pn, 1.0, "constant_velocity")
    # impulse (V, Omega) enters as v += V, w += Omega, r += V dt, q via d(q x q(w dt))/dw
    G = np.zeros((13, 6))
    for j in range(6):
        e = np.zeros(13)
        e[7 + j] = 1e-7
        xp, xm = xv + e, xv - e
        G[:, j] = (O.fv(xp, 1.0, "constant_velocity") - O.fv(xm, 1.0, "constant_velocity")) / 2e-7
        G[7 + j, j] = 1.0
    np.testing.assert_allclose(Q, G @ Pn @ G.T, rtol=1e-5, atol=1e-12)
    assert np.allclose(Q, Q.T)


def test_normjac_fd():
    rng = np.random.RandomState(6)
    q = rng.normal(size=4) * 1.3
    J = O.normJac(q)
    Jfd = np.zeros((4, 4))
    for j in range(4):
        e = np.zeros(4)
        e[j] = 1e-7
        Jfd[:, j] = ((q + e) / np.linalg.norm(q + e) - (q - e) / np.linalg.norm(q - e)) / 2e-7
    np.testing.assert_allclose(J, Jfd, rtol=1e-6, atol=1e-8)


def test_feature_init_jacobians_fd():
    """dy/dxv and dy/dhd of mc/add_a_feature_covariance_inverse_depth.m against hinv()."""
    rng = np.random.RandomState(7)
    cam = O.initialize_cam()
    xv = _rand_pose(rng)
    uvd = np.array([101.3, 77.9])
    P = np.zeros((13, 13))
    Pres = O.add_a_feature_covariance_inverse_depth(P, uvd, xv, 1.0, 1.0, cam)
    # with P = 0 the new block is dy_dhd Padd dy_dhd'; rebuild dy_dhd by finite differences
    J = np.zeros((6, 3))
    for j in range(2):
        e = np.zeros(2)
        e[j] = 1e-4
        J[:, j] = (O.hinv(uvd + e, xv, cam, 1.0) - O.hinv(uvd - e, xv, cam, 1.0)) / 2e-4
    J[5, 2] = 1.0
    np.testing.assert_allclose(Pres[13:, 13:], J @ J.T, rtol=1e-5, atol=1e-12)


def test_hinv_roundtrip():
    rng = np.random.RandomState(8)
    cam = O.initialize_cam()
    xv = _rand_pose(rng)
    for uv in ([60.0, 50.0], [250.5, 200.25], [160.0, 120.0]):
        y = O.hinv(np.array(uv), xv, cam, 0.37)
        h = O.hi_inverse_depth(y, xv[0:3], O.q2r(xv[3:7]), cam)
        np.testing.assert_allclose(h, uv, rtol=0, atol=1e-6)  # undistort_fm is the model, distort_fm its Newton inverse
