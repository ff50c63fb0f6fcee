"""Analysis only: times the downdate with parts of the pipeline disabled (results are garbage)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ekf_slam_b200 as pkg, ekf_slam_b200.synth as synth
from ekf_slam_b200 import _lib
B, N = 2048, 100
seq = synth.SynthSequence(B=B, N=N, T=8, seed=1)
lib = _lib.load()
lib.ekfslam_debug_flag.argtypes = [ctypes.c_int]
for flag in (0, 1, 2, 4, 3, 7):
    bank = pkg.FilterBank(B, N)
    bank.reset_filters()
    for k in range(N):
        bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
    for t in range(1, 5):
        zc, has = seq.frame(t)
        bank.upload_candidates(zc, has); bank.upload_uniforms(seq.uniforms(t)); bank.step()
    # one more frame up to the li downdate, with timing; only the flagged launch is perturbed
    zc, has = seq.frame(5)
    bank.upload_candidates(zc, has); bank.upload_uniforms(seq.uniforms(5))
    bank.begin_frame(); bank.ekf_prediction(); bank.measure(1); bank.gate(); bank.ransac_hypotheses()
    bank.enable_timing(True)
    lib.ekfslam_debug_flag(flag)
    bank.ekf_update_li_inliers()
    lib.ekfslam_debug_flag(0)
    kt = bank.kernel_times()
    st = bank.download_stats()
    print("flag", flag, "k_li mean", 2 * st["n_li"].mean(), "downdate ms (B=%d)" % B, round(kt["k_downdate"][0], 3), flush=True)
    bank.close()
