"""Runs the UNMODIFIED reference sources (/root/reference/matlab_code/*.m) through the mini
interpreter over synthetic filter steps and returns / stores what they produce.

TEST INFRASTRUCTURE (oracle/).  The reference cannot travel to the GPU box, so the outputs are
committed as fixtures under tests/golden/ (tests/golden/make_ref_steps.py is the generating script).

What is NOT reference code on this path (and why):
  baseline/octave/ref_frame.m, ref_features_info.m, ref_map.m  harness: per-frame call order of
      matlab_code/mono_slam.m:50-82 with the image matcher replaced by its own gating rule on
      supplied candidate pixels (matching.m needs the CV Toolbox and an absent image sequence)
  baseline/octave/shims/quaternions.m, dq3_by_dq1.m            functions the reference calls but
      does not ship (v2q.m:15, dfv_by_dxv.m:13, func_Q.m:24)
`rand` is a builtin of the interpreter fed from the stored uniform stream, so the reference's own
select_random_match.m runs as shipped.
"""
import os

import numpy as np

from .minterp import Interp, StructArr, EMPTY

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_DIR = os.environ.get("EKFSLAM_REFERENCE_DIR", "/root/reference/matlab_code")
HARNESS_DIR = os.path.join(REPO, "baseline", "octave")

F_HAS_H, F_HAS_Z, F_IC, F_LI, F_HI = 1, 2, 4, 8, 16


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "ransac_hypotheses.m"))


def make_interp():
    if not reference_available():
        raise RuntimeError("reference sources not found under %s" % REF_DIR)
    # harness first (it holds no function named like a reference function), then the two missing
    # functions, then the reference itself
    return Interp([HARNESS_DIR, os.path.join(HARNESS_DIR, "shims"), REF_DIR])


def make_filter(I, x, P, std_a=0.007, std_alpha=0.007, std_z=1.0):
    return I.call("ekf_filter", np.asarray(x, dtype=np.float64).reshape(-1, 1), np.asarray(P, dtype=np.float64),
                  std_a, std_alpha, std_z, "constant_velocity")


def flags_of(features_info, N=None):
    """features_info (StructArr) -> the flag byte per feature used by the GPU/oracle comparisons."""
    n = len(features_info.elems)
    out = np.zeros(N or n, dtype=np.uint8)
    for i, e in enumerate(features_info.elems):
        v = 0
        if e["h"].size:
            v |= F_HAS_H
        if e["z"].size:
            v |= F_HAS_Z
        if float(np.asarray(e["individually_compatible"]).reshape(-1)[0]):
            v |= F_IC
        if float(np.asarray(e["low_innovation_inlier"]).reshape(-1)[0]):
            v |= F_LI
        if float(np.asarray(e["high_innovation_inlier"]).reshape(-1)[0]):
            v |= F_HI
        out[i] = v
    return out


def types_of(features_info):
    return np.array([1 if e["type"] == "inversedepth" else 2 for e in features_info.elems], dtype=np.uint8)


def run_sequence(x0, P0, types, zc, has, U, I=None, keep_P="last", on_frame=None):
    """One filter over T frames through the reference.

    x0 [n], P0 [n,n], types [N] (1 id / 2 cartesian), zc [T,N,2], has [T,N], U [T,n_u].
    Returns dict(x [T,n], P [n,n] (after the last frame) or [T,n,n], flags [T,N] u8, nhyp [T] =
    number of uniforms the reference drew = hypotheses executed, h [T,N,2] (NaN = not predicted)).
    """
    I = I or make_interp()
    T, N = zc.shape[0], zc.shape[1]
    cam = I.call("initialize_cam")
    filt = make_filter(I, x0, P0)
    fi = I.call("ref_features_info", np.asarray(types, dtype=np.float64).reshape(1, -1))
    xs, Ps, fls, nh, hs = [], [], [], [], []
    for t in range(T):
        I.rand_stream = iter(U[t])
        I.rand_drawn = 0
        filt, fi, _nic = I.call("ref_frame", filt, fi, cam, np.ascontiguousarray(zc[t].T),
                                has[t].astype(np.float64).reshape(1, -1), nargout=3)
        xs.append(filt.get("x_k_k").reshape(-1).copy())
        if keep_P == "all":
            Ps.append(np.array(filt.get("p_k_k")))
        fls.append(flags_of(fi, N))
        nh.append(I.rand_drawn)
        h = np.full((N, 2), np.nan)
        for i, e in enumerate(fi.elems):
            if e["h"].size:
                h[i] = e["h"].reshape(-1)
        hs.append(h)
        if on_frame is not None:
            on_frame(t, filt, fi)
    out = dict(x=np.array(xs), flags=np.array(fls), nhyp=np.array(nh, dtype=np.int32), h=np.array(hs))
    out["P"] = np.array(Ps) if keep_P == "all" else np.array(filt.get("p_k_k"))
    return out


# ------------------------------------------------------------------------------------------
# closed loop with map management (mono_slam.m:50-82): the reference's own map_management.m
# ------------------------------------------------------------------------------------------
def tags_of(features_info):
    return np.array([int(np.asarray(e["feature_when_initialized"]).reshape(-1)[0]) for e in features_info.elems],
                    dtype=np.int32)


def run_closed_loop(world, b, frames, min_features=25, K=25, n_u=64, I=None, u_seed=0):
    """Filter b of a SynthWorld through `frames` iterations of the mono_slam.m loop, reference code:
    map_management.m (delete_features shim, update_features_info, inversedepth_2_cartesian,
    initialize_features shim -> add_features_inverse_depth + add_feature_to_info_vector) then the
    filter step of ref_frame.m.  Detections come from the image of the previous frame and the
    candidates from the current one, exactly like `im` in mono_slam.m:47,53,59.
    Returns per frame: x, types, tags, flags, nhyp, counters + the inputs that were fed (detections,
    candidates, uniforms) so that the oracle / the GPU can be driven with the very same data."""
    from .minterp import StructArr, mat
    I = I or make_interp()
    cam = I.call("initialize_cam")
    x, P = I.call("initialize_x_and_p", nargout=2)
    filt = make_filter(I, x, P)
    fi = np.zeros((0, 0))
    rng = np.random.RandomState(u_seed + 7919 * b)
    out = dict(x=[], P=None, types=[], tags=[], flags=[], nhyp=[], counters=[], det_uv=[], det_tag=[], det_n=[],
               zc=[], has=[], U=[], n_after_mm=[])
    blank = np.zeros((int(world.cam.nRows), int(world.cam.nCols)))
    for step in range(1, frames + 1):
        nf = len(fi.elems) if isinstance(fi, StructArr) else 0
        tags = tags_of(fi) if nf else np.zeros(0, dtype=np.int32)
        # NOTE: the detector must not re-detect features already in the map; features deleted by
        # map_management in THIS call are still excluded (they were in the map when the image was searched)
        tg_in = np.full((world.B, max(nf, 1)), -1, dtype=np.int32)
        nfa = np.zeros(world.B, dtype=np.int32)
        tg_in[b, :nf] = tags
        nfa[b] = nf
        uv, tg, nd = world.detections(step - 1, tg_in, nfa, K)
        im = StructArr(["uv", "tag", "img"], [{"uv": np.ascontiguousarray(uv[b, :nd[b]].T),
                                                "tag": tg[b, :nd[b]].astype(np.float64).reshape(1, -1),
                                                "img": blank}])
        filt, fi = I.call("map_management", filt, fi, cam, im, float(min_features), float(step), nargout=2)
        nf = len(fi.elems)
        tags = tags_of(fi)
        tg_in = np.full((world.B, max(nf, 1)), -1, dtype=np.int32)
        nfa = np.zeros(world.B, dtype=np.int32)
        tg_in[b, :nf] = tags
        nfa[b] = nf
        zc, has = world.candidates(step, tg_in, nfa)
        u = rng.rand(n_u)
        I.rand_stream = iter(u)
        I.rand_drawn = 0
        out["n_after_mm"].append(filt.get("x_k_k").shape[0])
        filt, fi, _nic = I.call("ref_frame_nomm", filt, fi, cam, np.ascontiguousarray(zc[b, :nf].T),
                                has[b, :nf].astype(np.float64).reshape(1, -1), nargout=3)
        out["x"].append(filt.get("x_k_k").reshape(-1).copy())
        out["types"].append(types_of(fi))
        out["tags"].append(tags.copy())
        out["flags"].append(flags_of(fi))
        out["nhyp"].append(I.rand_drawn)
        out["counters"].append(np.array([[float(np.asarray(e["times_predicted"]).reshape(-1)[0]),
                                          float(np.asarray(e["times_measured"]).reshape(-1)[0])] for e in fi.elems]))
        out["det_uv"].append(uv[b].copy())
        out["det_tag"].append(tg[b].copy())
        out["det_n"].append(int(nd[b]))
        out["zc"].append(zc[b, :nf].copy())
        out["has"].append(has[b, :nf].copy())
        out["U"].append(u)
    out["P"] = np.array(filt.get("p_k_k"))
    return out
