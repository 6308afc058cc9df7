"""Times the FP32 (tcgen05) contraction K1-f32 alone.  usage: python tools/bench_f32.py [N D q]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_data
from pyvb_b200 import _cabi

lib = _cabi.lib()
dev = torch.device("cuda", 0)
cases = [(1000000, 256, 16), (500000, 1024, 32), (200000, 512, 64)]
if len(sys.argv) > 3:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
for N, D, q in cases:
    st = torch.cuda.current_stream(dev).cuda_stream
    X = make_data(torch, N, D, q, 0.2, 1, dev)
    planes = torch.zeros(3, N, D, dtype=torch.bfloat16, device=dev)
    _cabi.check(lib.pyvb_prepare_x_f32(N, D, X.data_ptr(), D, planes.data_ptr(), st), "prep")
    del X
    ncp = int(lib.pyvb_f32_pitch(q)); P = q * (q + 1) // 2
    g = torch.Generator(device=dev); g.manual_seed(2)
    W = torch.randn(D, q, generator=g, device=dev, dtype=torch.float64)
    V = torch.ones(D, q, device=dev, dtype=torch.float64); mu = torch.zeros(D, device=dev, dtype=torch.float64)
    GT = torch.zeros(3, ncp, D, dtype=torch.bfloat16, device=dev); WT = torch.zeros(3, q, D, dtype=torch.bfloat16, device=dev)
    _cabi.check(lib.pyvb_pack_gw_f32(D, q, W.data_ptr(), V.data_ptr(), mu.data_ptr(), GT.data_ptr(), WT.data_ptr(), st), "pack")
    P0 = torch.eye(q, dtype=torch.float64, device=dev); h0 = torch.zeros(q, dtype=torch.float64, device=dev)
    gl = torch.zeros(144, dtype=torch.float64, device=dev); gl[2] = 2.0
    MZ = torch.zeros(N, ncp, dtype=torch.float32, device=dev)
    def run():
        _cabi.check(lib.pyvb_zstep_k1_f32(N, D, q, planes.data_ptr(), GT.data_ptr(), WT.data_ptr(), P0.data_ptr(), h0.data_ptr(),
                                          gl.data_ptr(), MZ.data_ptr(), st), "k1")
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    nct = (ncp + (256 if q == 64 else 192) - 1) // (256 if q == 64 else 192)
    bytes_alg = N * (D * 2.0 * 3 + ncp * 4.0)                      # mask + x_h + x_m once, MZ32 row out
    fl_alg = N * (2.0 * D * P + 2.0 * D * q)                       # SURVEY 8d K1 terms
    fl_mma = N * 2.0 * D * (3.0 * ncp + 5.0 * q)                   # issued bf16 MMA flops
    print("K1-f32 N=%d D=%d q=%d: %.3f ms  %.0f GB/s algorithmic  %.1f TF/s algorithmic  %.0f TF/s issued bf16" % (
        N, D, q, ms, bytes_alg / ms * 1e-6, fl_alg / ms * 1e-9, fl_mma / ms * 1e-9), flush=True)
