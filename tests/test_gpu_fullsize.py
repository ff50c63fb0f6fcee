"""BASELINE.json's full size (configs[2]: 4096 filters x N=100, n=613) through properties that do not need
the oracle at that size, plus oracle parity on a random sample of the same batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

B, N, FRAMES = 4096, 100, 3


@pytest.fixture(scope="module")
def run():
    import ekf_slam_b200 as pkg
    import ekf_slam_b200.synth as synth
    seq = synth.SynthSequence(B=B, N=N, T=FRAMES, seed=4242, n_u=64)

    def go(order=None):
        idx = np.arange(B) if order is None else order
        bank = pkg.FilterBank(B, N)
        bank.reset_filters()
        for k in range(N):
            bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, idx, k]))
        out = []
        for t in range(1, FRAMES + 1):
            bank.upload_candidates(np.ascontiguousarray(seq.zc[t, idx]), np.ascontiguousarray(seq.has[t, idx]))
            bank.upload_uniforms(np.ascontiguousarray(seq.U[idx, t]))
            bank.step()
            out.append((bank.download_flags(), bank.download_stats()))
        x, _, _ = bank.download_state(want_P=False)
        # covariances of a strided subset (the full 12 GB stays on the device)
        Psub = {}
        for b in range(0, B, 512):
            _, Pb, _ = bank.download_state(b0=b, nb=1)
            Psub[b] = Pb[0]
        bank.close()
        return x, out, Psub

    return seq, go


def test_fullsize_properties_and_determinism(run):
    import ekf_slam_b200 as pkg
    seq, go = run
    x, out, Psub = go()
    # unit quaternions, finite state
    assert np.isfinite(x).all()
    np.testing.assert_allclose(np.linalg.norm(x[:, 3:7], axis=1), 1.0, rtol=0, atol=4e-16)
    for flags, st in out:
        ic, li, hi = (flags & pkg.F_IC) != 0, (flags & pkg.F_LI) != 0, (flags & pkg.F_HI) != 0
        assert not (li & ~ic).any() and not (hi & ~ic).any() and not (li & hi).any()
        assert ((flags & pkg.F_HAS_Z) != 0).sum() == ic.sum()
        assert np.array_equal(st["n_ic"], ic.sum(1)) and np.array_equal(st["n_li"], li.sum(1))
        assert np.array_equal(st["n_hi"], hi.sum(1)) and (st["status"] == 0).all()
        assert (st["ransac_iters"] >= 1).all() and (st["ransac_scored"] <= st["ransac_iters"]).all()
        assert (st["max_support"] == li.sum(1)).all()
        assert li.sum() > 30 * B  # the filters are tracking
    for b, P in Psub.items():
        assert np.array_equal(P, P.T)                       # exactly symmetric
        assert np.linalg.eigvalsh(P).min() > -1e-10 * np.abs(P).max()   # PSD to rounding
    # bit-exact determinism and permutation equivariance over the batch index
    rng = np.random.RandomState(0)
    perm = rng.permutation(B)
    x2, out2, Psub2 = go(perm)
    assert np.array_equal(x2, x[perm])
    for (f1, s1), (f2, s2) in zip(out, out2):
        assert np.array_equal(f2, f1[perm])
        for k in s1:
            assert np.array_equal(s2[k], s1[k][perm])
    inv = np.argsort(perm)
    for b, P in Psub.items():
        j = inv[b]
        if j % 512 == 0:
            assert np.array_equal(Psub2[j], P)


def test_fullsize_sample_matches_c_oracle(run):
    """32 filters of the 4096, free-running over the same frames, against the C restatement."""
    from oracle import c_oracle
    import ekf_slam_b200.synth as synth
    seq, go = run
    x, out, Psub = go()
    sample = np.arange(0, B, 128)
    xs, Ps, types = seq.initial_state(0, 1)           # shapes only
    n = 13 + 6 * N
    xo = np.zeros((len(sample), n))
    Po = np.zeros((len(sample), n, n))
    for i, b in enumerate(sample):
        xb, Pb, _ = seq.initial_state(int(b), 1)
        xo[i], Po[i] = xb[0], Pb[0]
    ty = np.ones((len(sample), N), dtype=np.uint8)
    nf = np.full(len(sample), N, dtype=np.int32)
    for t in range(1, FRAMES + 1):
        fl, st = c_oracle.step_batch(xo, Po, ty, nf, np.ascontiguousarray(seq.zc[t, sample]),
                                     np.ascontiguousarray(seq.has[t, sample]), np.ascontiguousarray(seq.U[sample, t]))
        assert np.array_equal(fl, out[t - 1][0][sample] & 31)
        assert np.array_equal(st[:, 0], out[t - 1][1]["ransac_iters"][sample])
    err = np.abs(x[sample] - xo).max() / np.abs(xo).max()
    assert err < 1e-9
    for i, b in enumerate(sample):
        if int(b) in Psub:
            assert np.abs(Psub[int(b)] - Po[i]).max() / np.abs(Po[i]).max() < 1e-9
