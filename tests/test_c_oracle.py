"""The C restatement (oracle/ekf_oracle.c) against the numpy oracle — two independent CPU statements
of the reference must agree before either is trusted as a checker or timed as a baseline."""
import numpy as np
import pytest

import ekf_slam_b200.synth as synth
from oracle import c_oracle, ekf_oracle as O
from tests import helpers as T


@pytest.mark.parametrize("N,frames,fixed,cart", [(10, 5, 0, None), (24, 4, 0, [1, 2, 7, 20]), (16, 3, 40, None)])
def test_c_oracle_matches_numpy_oracle(N, frames, fixed, cart):
    B = 3
    n_u = max(64, fixed)
    seq = synth.SynthSequence(B=B, N=N, T=frames, seed=900 + N, n_u=n_u)
    x0, P0, types = seq.initial_state()
    n_max = 13 + 6 * N
    if cart:
        xs, Ps, ts = [], [], []
        for b in range(B):
            xb, Pb, tb = synth.convert_to_cartesian(x0[b], P0[b], types[b], cart)
            Pb = np.tril(Pb) + np.tril(Pb, -1).T
            xs.append(np.pad(xb, (0, n_max - len(xb))))
            Ps.append(np.pad(Pb, ((0, n_max - len(xb)), (0, n_max - len(xb)))))
            ts.append(tb)
        x0, P0, types = np.stack(xs), np.stack(Ps), np.stack(ts)
    cam = O.initialize_cam()
    nfeat = np.full(B, N, dtype=np.int32)
    nstate = np.array([13 + sum(6 if t == 1 else 3 for t in types[b]) for b in range(B)])
    xc, Pc = x0.copy(), np.ascontiguousarray(P0.copy())
    filts = [T.oracle_filter(x0[b, :nstate[b]], P0[b, :nstate[b], :nstate[b]]) for b in range(B)]
    feats = [T.oracle_features(types[b]) for b in range(B)]
    for t in range(1, frames + 1):
        zc, has = seq.frame(t)
        u = seq.uniforms(t, n_u)
        flags, stats = c_oracle.step_batch(xc, Pc, types, nfeat, np.ascontiguousarray(zc), np.ascontiguousarray(has), u,
                                           fixed_hyp=fixed, nthreads=2)
        for b in range(B):
            filts[b], feats[b], info = T.oracle_step(filts[b], feats[b], cam, zc[b], has[b], u[b], fixed)
            n = nstate[b]
            assert np.array_equal(flags[b], T.oracle_flags(feats[b], N)), (t, b)
            assert stats[b, 0] == info["iterations"] and stats[b, 3] == 0
            assert T.rel_err(xc[b, :n], filts[b].x_k_k) < 1e-10
            assert T.rel_err(Pc[b, :n, :n], filts[b].p_k_k) < 1e-9


def test_c_oracle_is_multithreaded_and_deterministic():
    seq = synth.SynthSequence(B=6, N=8, T=2, seed=5)
    x0, P0, types = seq.initial_state()
    nfeat = np.full(6, 8, dtype=np.int32)
    outs = []
    for nt in (1, 4):
        x, P = x0.copy(), P0.copy()
        zc, has = seq.frame(1)
        fl, st = c_oracle.step_batch(x, P, types, nfeat, np.ascontiguousarray(zc), np.ascontiguousarray(has), seq.uniforms(1), nthreads=nt)
        outs.append((x, P, fl))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert c_oracle.max_threads() >= 1
