"""Probe: does splitting the batch over S concurrent streams (S FilterBanks of B/S filters, stepped back to back
each frame) beat one bank of B filters?  Latency-bound kernels of one part could overlap heavy kernels of another.
usage: split_probe.py [B] [S ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ekf_slam_b200 as pkg
import ekf_slam_b200.synth as synth


def run(B, S, N=100, warm=4, timed=8):
    T = warm + timed
    nb = B // S
    parts = []
    for i in range(S):
        seq = synth.SynthSequence(B=nb, N=N, T=T, seed=1, b_offset=i * nb, n_u=64)
        bank = pkg.FilterBank(nb, N)
        st = torch.cuda.Stream()
        bank.set_stream(st.cuda_stream)
        bank.set_params(fixed_hyp=0)
        bank.reset_filters()
        for k in range(N):
            bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
        zc = torch.from_numpy(seq.zc).cuda(); fl = torch.from_numpy((seq.has * pkg.F_CAND).astype(np.uint8)).cuda()
        u = torch.from_numpy(np.ascontiguousarray(np.transpose(seq.U, (1, 0, 2)))).cuda()
        parts.append((bank, st, zc, fl, u))

    def step(t):
        for bank, st, zc, fl, u in parts:
            bank.bind_frame(zc[t].data_ptr(), fl[t].data_ptr(), u[t].data_ptr(), 64)
            bank.step(reset=True, match_mode=1)

    for t in range(1, warm + 1):
        step(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(warm + 1, T + 1):
        step(t)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / timed
    for p in parts:
        p[0].close()
    return ms


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for S in [int(a) for a in sys.argv[2:]] or [1, 2, 4]:
        ms = run(B, S)
        print("B=%d split over %d streams: %.2f ms/step, %.0f filter-steps/s" % (B, S, ms, B / ms * 1e3), flush=True)
