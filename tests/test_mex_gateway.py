"""The Octave/MATLAB MEX gateway cannot be linked here (no mex.h / mkoctfile in the image): it is
syntax-checked against a declaration-only stub, and its calls are checked against the C ABI header."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mex_gateway_compiles_against_stub():
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    r = subprocess.run([cc, "-fsyntax-only", "-Wall", "-Werror=implicit-function-declaration",
                        "-I" + os.path.join(ROOT, "mex", "stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "mex", "ekfslam_mex.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_mex_gateway_only_calls_declared_abi():
    hdr = open(os.path.join(ROOT, "include", "ekfslam.h")).read()
    declared = set(re.findall(r"\b(ekfslam_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "mex", "ekfslam_mex.c")).read(), flags=re.S)
    used = set(re.findall(r"\b(ekfslam_[a-z0-9_]+)\s*\(", src)) - {"ekfslam_mex"}
    assert used and used <= declared, used - declared


def test_every_hot_path_function_has_a_shim():
    shims = {f[:-2] for f in os.listdir(os.path.join(ROOT, "mex", "shims")) if f.endswith(".m")}
    assert shims == {"ekf_prediction", "search_IC_matches", "ransac_hypotheses", "ekf_update_li_inliers",
                     "rescue_hi_inliers", "ekf_update_hi_inliers"}
    src = open(os.path.join(ROOT, "mex", "ekfslam_mex.c")).read()
    for s in shims:
        assert '"%s"' % s in src


def test_octave_probe_is_reported():
    """north_star asks for an Octave cross-run; the image has none — record that explicitly."""
    if shutil.which("octave") is None:
        pytest.skip("OCTAVE ABSENT - the reference's own CPU execution cannot be run in this image")
