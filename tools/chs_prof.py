"""Phase breakdown of the 112-row resident Cholesky (debug build -DCHS_PROF): cycles of thread 0 per phase, mean per CTA.
usage (on the GPU box): bash ekf-slam_b200/csrc/build.sh -DCHS_PROF && python tools/chs_prof.py"""
import ctypes, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ekf_slam_b200 as pkg
import ekf_slam_b200.synth as synth
B, N, T = 1024, 100, 8
seq = synth.SynthSequence(B=B, N=N, T=T, seed=1, n_u=64)
bank = pkg.FilterBank(B, N)
bank.set_params(fixed_hyp=0); bank.reset_filters()
for k in range(N): bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
zc = torch.from_numpy(seq.zc).cuda(); fl = torch.from_numpy((seq.has * pkg.F_CAND).astype(np.uint8)).cuda()
u = torch.from_numpy(np.ascontiguousarray(np.transpose(seq.U, (1, 0, 2)))).cuda()
lib = bank.lib
out = (ctypes.c_ulonglong * 16)()
names = ["load S", "copy D", "diag block: inverse (warp 0) || inverse row p-1", "panel solve", "trailing update", "-", "last inverse row", "store X, y, cv", "diag block: factor (warp 0)"]
for t in range(1, T + 1):
    bank.bind_frame(zc[t].data_ptr(), fl[t].data_ptr(), u[t].data_ptr(), 64)
    if t == 5: lib.ekfslam_debug_chs_prof(None, 1)
    bank.step(reset=True, match_mode=1)
lib.ekfslam_debug_chs_prof(out, 0)
n = max(out[15], 1); tot = sum(out[i] for i in range(9))
print("CTAs", out[15], "mean cycles per CTA", round(tot / n))
for i, nm in enumerate(names): print("  %-24s %8.0f  %5.1f%%" % (nm, out[i] / n, 100.0 * out[i] / max(tot, 1)))
