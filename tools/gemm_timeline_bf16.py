"""In-kernel timeline (CTA 0) of the bf16 GEMM for a few C2 problems: where a tile's time goes (operand latency, MMA, epilogue)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tacotron_b200  # noqa: F401,E402
from importlib import import_module  # noqa: E402

capi = import_module("multi-speaker-tacotron-tensorflow_b200.capi")
lib = capi.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
BF = torch.bfloat16


def run(name, kw, cold=True):
    d = capi.TacoGemmDesc(); d.alpha = 1.0; d.split_k = 1
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            if k in ("A", "B"):
                v16 = v.to(BF); keep.append(v16); setattr(d, k + "16", v16.data_ptr())
            setattr(d, k, v.data_ptr())
        else:
            setattr(d, k, v)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        capi.check(lib.taco_gemm(C.byref(d), 1, 2, st))
    if cold:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); capi.check(lib.taco_gemm(C.byref(d), 1, 2, st)); e1.record()
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 64)()
    lib.taco_debug_timeline_bf16(out)
    t0 = out[0]
    print("%s: %.1f us by events (%s)" % (name, e0.elapsed_time(e1) * 1e3, "L2 cold" if cold else "warm"))
    print("   setup %.2f us;  epilogue chunk (clk): tmem ld %d, stage %d, store loop %d" % ((out[1] - t0) * 1e-3, out[60], out[61], out[62]))
    for lt in range(15):
        v = [out[2 + 4 * lt + i] for i in range(4)]
        if v[0] <= t0 or v[3] <= t0:
            break
        print("   tile %d: operands landed %.2f  last MMA issued %.2f  accumulator ready %.2f  epilogue done %.2f" %
              (lt, *[(x - t0) * 1e-3 for x in v]))


def main():
    R = 25824
    A = torch.randn(R, 256, device=dev); W = torch.randn(256, 256, device=dev) * 0.05; bias = torch.randn(256, device=dev)
    Cc = torch.zeros(R, 256, device=dev)
    kw = dict(A=A, B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, bias=bias, act=2)
    run("highway NN 25824x256x256", kw)
    run("highway NN 25824x256x256", kw, cold=False)
    gx = torch.zeros(R, 1536, device=dev); Wx = torch.randn(256, 1536, device=dev) * 0.05; bx = torch.randn(1536, device=dev)
    run("gru x-side NN 25824x1536x256", dict(A=A, B=Wx, C=gx, M=R, N=1536, K=256, lda=256, ldb=1536, ldc=1536, bias=bx))
    big = torch.randn(R + 64, 2048, device=dev); W1 = torch.randn(3 * 2048, 256, device=dev) * 0.02
    run("proj_1 conv 25824x256x6144", dict(A=big, B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048, bias=bias, act=1,
                                          mask_period=807, mask_lo=3, mask_hi=803))


if __name__ == "__main__":
    lib.taco_debug_timeline_bf16.argtypes = [C.c_void_p]
    main()
