"""Device-resident vs host-buffer step on neighbouring frames of the same run (is the e2e overhead real?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ekf_slam_b200 as pkg
import ekf_slam_b200.synth as synth
B, N, n_u = 4096, 100, 64
T = 60
seq = synth.SynthSequence(B=B, N=N, T=T, seed=1, n_u=n_u)
bank = pkg.FilterBank(B, N)
stream = torch.cuda.Stream(); bank.set_stream(stream.cuda_stream)
bank.reset_filters()
for k in range(N):
    bank.add_features_inverse_depth(np.ascontiguousarray(seq.zc[0, :, k]))
fl_np = (seq.has * pkg.F_CAND).astype(np.uint8)
zc_pin = torch.from_numpy(seq.zc).pin_memory(); fl_pin = torch.from_numpy(fl_np).pin_memory()
u_pin = torch.from_numpy(np.ascontiguousarray(np.transpose(seq.U, (1, 0, 2)))).pin_memory()
zc_dev, fl_dev, u_dev = zc_pin.cuda(), fl_pin.cuda(), u_pin.cuda()
n = 13 + 6 * N
x_out = torch.empty((B, n), dtype=torch.float64).pin_memory(); f_out = torch.empty((B, N), dtype=torch.uint8).pin_memory()
s_out = torch.empty((B, 8), dtype=torch.int32).pin_memory()
zc_h, fl_h, u_h = zc_pin.numpy(), fl_pin.numpy(), u_pin.numpy()
xo, fo, so = x_out.numpy(), f_out.numpy(), s_out.numpy()
def dev_steps(t0, t1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); w = time.perf_counter(); e0.record(stream)
    for t in range(t0, t1):
        bank.bind_frame(zc_dev[t].data_ptr(), fl_dev[t].data_ptr(), u_dev[t].data_ptr(), n_u)
        bank.step(reset=True, match_mode=1)
    e1.record(stream); torch.cuda.synchronize()
    bank.unbind_frame()
    return e0.elapsed_time(e1) / (t1 - t0), 1e3 * (time.perf_counter() - w) / (t1 - t0)
def host_steps(t0, t1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); w = time.perf_counter(); e0.record(stream)
    for t in range(t0, t1):
        bank.step_host(zc_h[t], fl_h[t], u_h[t], match_mode=1, x_out=xo, flags_out=fo, stats_out=so)
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (t1 - t0), 1e3 * (time.perf_counter() - w) / (t1 - t0)
print("warm", dev_steps(1, 5))
for a in range(5, 55, 10):
    d = dev_steps(a, a + 5); h = host_steps(a + 5, a + 10)
    st = bank.download_stats()
    print("frames %2d-%2d dev %.2f ms (wall %.2f) | frames %2d-%2d host %.2f ms (wall %.2f) | k_li %.1f k_hi %.1f" %
          (a, a + 4, d[0], d[1], a + 5, a + 9, h[0], h[1], 2 * st["n_li"].mean(), 2 * st["n_hi"].mean()))
