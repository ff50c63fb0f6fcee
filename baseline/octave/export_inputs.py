"""Writes the synthetic workload of bench.py / the parity tests as a .mat file for run_ref_step.m.

    python baseline/octave/export_inputs.py --filters 4 --features 100 --frames 5 --out inputs.mat

Uses only the product package's generator (ekf-slam_b200/synth.py); the same seeds give the GPU, the CPU
oracle and Octave identical states, candidate pixels and RANSAC uniform streams.
"""
import argparse
import os
import sys

import numpy as np
import scipy.io

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ekf_slam_b200.synth import SynthSequence  # noqa: E402


def export(path, B, N, T, seed=0, p_outlier=0.2, n_u=64, std_a=0.007, std_alpha=0.007, std_z=1.0):
    seq = SynthSequence(B, N, T, seed=seed, p_outlier=p_outlier, n_u=n_u)
    x0, P0, types = seq.initial_state()
    scipy.io.savemat(path, {
        "B": B, "N": N, "T": T, "x0": x0, "P0": P0, "types": types.astype(np.float64),
        "zc": seq.zc[1:T + 1], "has": seq.has[1:T + 1].astype(np.float64), "U": seq.U[:, 1:T + 1],
        "std_a": std_a, "std_alpha": std_alpha, "std_z": std_z}, do_compression=True)
    return seq, x0, P0, types


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=4)
    ap.add_argument("--features", type=int, default=100)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="inputs.mat")
    a = ap.parse_args()
    export(a.out, a.filters, a.features, a.frames, seed=a.seed)
    print("wrote", a.out)
