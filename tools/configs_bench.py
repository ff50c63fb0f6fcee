"""Times the other BASELINE.json configs (parity-test shapes, not bench lines) on one GPU and writes
gpurun_out/configs.json:  cfg2 single filter N=100, 256 hypotheses (latency); cfg4 N=500 large map;
cfg5 mixed inverse-depth/Cartesian, 512 hypotheses, batch 1024, with and without the iterated update."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ekf_slam_b200 as pkg
import ekf_slam_b200.synth as synth

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "fp64_peak.json")))["dmma_m8n8k4_tflops"]


def run(name, B, N, fixed, frames_warm, frames_timed, cart_frac=0.0, iterated=False, seed=1, graph=False):
    n_u = max(64, fixed)
    T = frames_warm + frames_timed
    seq = synth.SynthSequence(B=B, N=N, T=T, seed=seed, n_u=n_u)
    bank = pkg.FilterBank(B, N)
    stream = torch.cuda.Stream()
    bank.set_stream(stream.cuda_stream)
    bank.set_params(fixed_hyp=fixed)
    bank.reset_filters()
    for k in range(N):
        bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
    if cart_frac > 0:
        for i in range(int(cart_frac * N)):
            bank.inversedepth_2_cartesian(force_index=i)
    zc = torch.from_numpy(seq.zc).cuda(); fl = torch.from_numpy((seq.has * pkg.F_CAND).astype(np.uint8)).cuda()
    u = torch.from_numpy(np.ascontiguousarray(np.transpose(seq.U, (1, 0, 2)))).cuda()

    def step(t):
        if graph:   # latency path: the frame is copied into the context's own buffers, the step replays a CUDA graph
            bank.stage_frame(zc[t].data_ptr(), fl[t].data_ptr(), u[t].data_ptr(), n_u)
            bank.step(reset=True, match_mode=1, graph=True)
            return
        bank.bind_frame(zc[t].data_ptr(), fl[t].data_ptr(), u[t].data_ptr(), n_u)
        if not iterated:
            bank.step(reset=True, match_mode=1)
        else:
            bank.begin_frame(); bank.ekf_prediction(); bank.measure(1); bank.gate(); bank.ransac_hypotheses()
            bank.update_iterated(pkg.F_LI, 1, 3)
            bank.rescue_hi_inliers(); bank.ekf_update_hi_inliers()

    for t in range(1, frames_warm + 1):
        step(t)
    torch.cuda.synchronize()
    if not graph:
        bank.enable_timing(True)   # per-kernel event pairs (a timed step is not graph-captured)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record(stream)
    for t in range(frames_warm + 1, T + 1):
        step(t)
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - w0
    ms = e0.elapsed_time(e1) / frames_timed
    kt = {} if graph else {k: v[0] / frames_timed for k, v in bank.kernel_times().items() if v[1] > 0}
    st = bank.download_stats()
    _, _, ns = bank.download_state(want_P=False)
    n = float(ns.mean())
    k_li, k_hi = 2 * st["n_li"].mean(), 2 * st["n_hi"].mean()
    out = {"config": name, "B": B, "N": N, "n_mean": n, "fixed_hyp": fixed, "ms_per_step": ms,
           "wall_ms_per_step": 1e3 * wall / frames_timed, "filter_steps_per_s": B / (ms * 1e-3),
           "mean_k_li": float(k_li), "mean_k_hi": float(k_hi), "mean_hyp_drawn": float(st["ransac_iters"].mean()),
           "mean_hyp_scored": float(st["ransac_scored"].mean()),
           "ransac_hyps_per_s_kernel": float(st["ransac_iters"].mean() * B / (kt["k_ransac"] * 1e-3)) if "k_ransac" in kt else None,
           "kernels_ms_per_step": {k: round(v, 4) for k, v in sorted(kt.items(), key=lambda kv: -kv[1])},
           "status_flags": int((st["status"] != 0).sum())}
    if "k_downdate" in kt:
        tf = B * n * n * k_li / (kt["k_downdate"] * 1e-3) / 1e12
        out["li_downdate_tflops_fp64"] = tf
        out["li_downdate_frac_of_dmma_peak"] = tf / PEAK
    bank.close()
    return out


if __name__ == "__main__":
    res = [run("cfg2: single filter, N=100, 256 hypotheses/frame (latency)", 1, 100, 256, 6, 20),
           run("cfg2 through the captured step graph (ekfslam_step_graph + ekfslam_stage_frame)", 1, 100, 256, 6, 40, graph=True),
           run("cfg4: large map N=500 (n=3013), batch 8", 8, 500, 0, 3, 4),
           run("cfg4: large map N=500 (n=3013), batch 32", 32, 500, 0, 3, 4),
           run("cfg5: mixed 40% Cartesian, 512 hypotheses/frame, batch 1024", 1024, 100, 512, 4, 8, cart_frac=0.4),
           run("cfg5 + iterated li update (3 iterations, extension)", 1024, 100, 512, 4, 8, cart_frac=0.4, iterated=True),
           run("cfg3 shape with 256 fixed hypotheses (RANSAC hyps/s)", 4096, 100, 256, 3, 6)]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/configs.json", "w"), indent=1)
    for r in res:
        print(r["config"], "| ms/step", round(r["ms_per_step"], 3), "| steps/s", round(r["filter_steps_per_s"]),
              "| hyps/s", round(r["ransac_hyps_per_s_kernel"] or 0), "| li TF", round(r.get("li_downdate_tflops_fp64", 0), 1))
