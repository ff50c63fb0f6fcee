"""The Octave cross-run harness (baseline/octave/): SURVEY §8(c)/(d) ask for the reference's own functions to be run
under GNU Octave on the same inputs wherever an Octave binary exists.  This image has none, so the always-on tests are
static (the harness only calls functions the reference ships, plus the three documented shims; the exported .mat has the
shapes run_ref_step.m indexes); the cross-run itself executes, and is compared with the oracle, when `octave` is on PATH."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest
import scipy.io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OCT = os.path.join(ROOT, "baseline", "octave")
REF = "/root/reference/matlab_code"

BUILTINS = {
    "nargin", "fileparts", "mfilename", "addpath", "fullfile", "load", "double", "zeros", "size", "squeeze", "eye", "isempty",
    "full", "all", "eig", "inv", "tic", "toc", "diag", "printf", "nproc", "save", "end", "if", "for", "function", "global",
    "length", "floor", "sum", "find", "cos", "sin", "norm", "D", "X", "Pdiag", "Plast", "li", "hi", "ic", "nhyp", "P",
    "features_info", "EKFSLAM_U", "x", "v", "q",
}


def _called(path):
    src = open(path).read()
    src = re.sub(r"%.*", "", src)
    src = re.sub(r"'[^'\n]*'", "''", src)
    return set(re.findall(r"(?<![\.\w])([A-Za-z_][A-Za-z0-9_]*)\s*\(", src))


def test_harness_calls_only_reference_functions():
    harness = {"run_ref_step", "ref_frame", "ref_features_info"}
    names = set()
    for h in harness:
        names |= _called(os.path.join(OCT, h + ".m"))
    names = names - BUILTINS - harness - {"zc", "has", "types"}
    shims = {f[:-2] for f in os.listdir(os.path.join(OCT, "shims"))}
    # functions the reference calls but does not ship (+ the toolbox-bound corner search it cannot run)
    assert shims == {"quaternions", "dq3_by_dq1", "delete_features", "initialize_features"}
    assert os.listdir(os.path.join(OCT, "shims_octave")) == ["select_random_match.m"]
    hot = {"ekf_filter", "update_features_info", "ekf_prediction", "predict_camera_measurements", "calculate_derivatives",
           "get_x_k_km1", "get_p_k_km1", "ransac_hypotheses", "ekf_update_li_inliers", "rescue_hi_inliers",
           "ekf_update_hi_inliers", "initialize_cam"}
    assert names == hot, names ^ hot
    if os.path.isdir(REF):
        for n in names:
            assert os.path.exists(os.path.join(REF, n + ".m")), n
        # the two functions the shims supply really are absent upstream, the shadowed one really exists
        assert not os.path.exists(os.path.join(REF, "quaternions.m"))
        assert not os.path.exists(os.path.join(REF, "dq3_by_dq1.m"))
        assert os.path.exists(os.path.join(REF, "select_random_match.m"))


def test_export_shapes(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_inputs", os.path.join(OCT, "export_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = str(tmp_path / "inputs.mat")
    B, N, T = 2, 6, 3
    mod.export(out, B, N, T, seed=3)
    D = scipy.io.loadmat(out)
    n = 13 + 6 * N
    assert D["x0"].shape == (B, n) and D["P0"].shape == (B, n, n) and D["types"].shape == (B, N)
    assert D["zc"].shape == (T, B, N, 2) and D["has"].shape == (T, B, N) and D["U"].shape[:2] == (B, T)
    assert np.allclose(D["P0"], np.transpose(D["P0"], (0, 2, 1)))


def test_shim_quaternion_matrix_matches_oracle():
    """dq3_by_dq1.m / quaternions.m restate the same formulas as the oracle's (parsed, not executed)."""
    from oracle import ekf_oracle as O
    q = np.array([0.9, 0.1, -0.3, 0.2])
    r, x, y, z = q
    M = np.array([[r, -x, -y, -z], [x, r, -z, y], [y, z, r, -x], [z, -y, x, r]])
    assert np.allclose(O.dq3_by_dq1(q), M)
    src = open(os.path.join(OCT, "shims", "dq3_by_dq1.m")).read()
    assert "r -x -y -z" in src and "x  r -z  y" in src and "y  z  r -x" in src and "z -y  x  r" in src


@pytest.mark.skipif(shutil.which("octave") is None or not os.path.isdir(REF),
                    reason="OCTAVE ABSENT - the reference's own CPU execution cannot be run in this image")
def test_octave_crossrun_matches_oracle(tmp_path):
    import importlib.util
    from oracle import ekf_oracle as O
    from tests.helpers import oracle_filter
    spec = importlib.util.spec_from_file_location("export_inputs", os.path.join(OCT, "export_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    inp, outp = str(tmp_path / "inputs.mat"), str(tmp_path / "outputs.mat")
    B, N, T = 2, 20, 4
    seq, x0, P0, types = mod.export(inp, B, N, T, seed=11)
    subprocess.run(["octave", "--no-gui", "--quiet", "--eval",
                    "addpath('%s'); run_ref_step('%s','%s','%s')" % (OCT, inp, outp, REF)], check=True, timeout=1800)
    R = scipy.io.loadmat(outp)
    cam = O.initialize_cam()
    for b in range(B):
        filt = oracle_filter(x0[b], P0[b])
        fi = [O.Feature(type="inversedepth", yi=None, individually_compatible=0, low_innovation_inlier=0,
                        high_innovation_inlier=0, z=None, h=None, H=None, S=None, R=np.eye(2), times_predicted=0,
                        times_measured=0) for _ in range(N)]
        for t in range(1, T + 1):
            fi = O.update_features_info(fi)
            filt, fi = O.filter_step(filt, fi, cam, (seq.zc[t, b], seq.has[t, b]), seq.U[b, t])
            assert [f.low_innovation_inlier for f in fi] == list(R["li"][b, :, t - 1].astype(int))
            assert [f.high_innovation_inlier for f in fi] == list(R["hi"][b, :, t - 1].astype(int))
            np.testing.assert_allclose(filt.x_k_k, R["X"][b, :, t - 1], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(filt.p_k_k, R["Plast"][b], rtol=1e-9, atol=1e-13)
