"""Tuning driver of the tensor-core convolution: loads a stand-alone build of csrc/conv_tc.cu (nvcc -D switches) and
times K = 1024 / 4096 on 65536 x 8192 waveforms.  usage: tc_conv_tune.py <conv_tc.so>"""
import ctypes as C
import sys

import numpy as np
import torch

lib = C.CDLL(sys.argv[1])
lib.dspb_convolve_tc_workspace.restype = C.c_int64
n, L = 65536, 8192
x = (torch.randn((n, L), device="cuda") * 4 + 1000).contiguous()
rng = np.random.default_rng(0)
for K in (256, 1024, 4096):
    k = torch.from_numpy(rng.standard_normal(K).astype(np.float32)).cuda()
    P = L - K + 1
    out = torch.empty((n, P), device="cuda")
    need = int(lib.dspb_convolve_tc_workspace(C.c_int64(K)))
    ws = torch.empty(need, device="cuda")

    def run():
        rc = lib.dspb_convolve_tc_f32(C.c_void_p(x.data_ptr()), C.c_int64(L), C.c_int64(n), C.c_int64(L), C.c_void_p(k.data_ptr()),
                                      C.c_int64(K), C.c_int32(ord("v")), C.c_void_p(out.data_ptr()), C.c_int64(P), C.c_int64(P),
                                      C.c_void_p(ws.data_ptr()), C.c_int64(need), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3 * 1e-3
    ref = torch.nn.functional.conv1d(x[:64, None, :].double(), k.flip(0)[None, None, :].double())[:, 0, :]
    err = float((out[:64].double() - ref).abs().max() / ref.abs().max())
    print(f"K={K}: {t * 1e3:.2f} ms  {2.0 * K * P * n / t / 1e12:.1f} useful TFLOP/s  err {err:.2e}", flush=True)
