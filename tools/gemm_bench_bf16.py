"""Times representative taco_gemm problems of the C2 training step (CUDA events, L2-cold): bf16-operand tcgen05 kernel vs the
fp32-operand TF32 kernel, with the L2 -> SM byte model beside each (tile bytes per FLOP x the ~6300 B/clk chip-wide cap)."""
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


def bench(name, kw, flops, precs=(2, 1), reps=7):
    keep = []
    out = []
    st = torch.cuda.current_stream().cuda_stream
    for prec in precs:
        d = capi.TacoGemmDesc()
        d.alpha = 1.0
        d.split_k = 1
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                keep.append(v)
                if prec == 2 and k in ("A", "B"):
                    v16 = v.to(BF) if v.is_contiguous() else None
                    if v16 is None:      # strided view: convert the base storage and re-slice
                        raise RuntimeError("pass contiguous operands")
                    keep.append(v16)
                    setattr(d, k + "16", v16.data_ptr())
                setattr(d, k, v.data_ptr())
            else:
                setattr(d, k, v)
        for _ in range(2):
            capi.check(lib.taco_gemm(C.byref(d), 1, prec, st))
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(lib.taco_gemm(C.byref(d), 1, prec, st))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        out.append("%s %8.1f us %7.1f TFLOP/s" % ({1: "tf32", 2: "bf16"}[prec], us, flops / us / 1e6))
    print("%-46s %s" % (name, " | ".join(out)), flush=True)


def main():
    R = 25824
    A = torch.randn(R, 256, device=dev)
    W = torch.randn(256, 256, device=dev) * 0.05
    bias = torch.randn(256, device=dev)
    Cc = torch.zeros(R, 256, device=dev)
    bench("highway NN 25824x256x256 +bias+sigmoid", dict(A=A, B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, bias=bias, act=2), 2.0 * R * 256 * 256)
    bench("dgrad NT 25824x256x256", dict(A=A, B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, transB=1), 2.0 * R * 256 * 256)
    dW = torch.zeros(256, 256, device=dev)
    bench("wgrad TN 256x256xK=25824 split", dict(A=A, B=Cc, C=dW, M=256, N=256, K=R, lda=256, ldb=256, ldc=256, transA=1, accumulate=1, split_k=64), 2.0 * R * 256 * 256)
    big = torch.randn(R + 64, 2048, device=dev)
    W1 = torch.randn(3 * 2048, 256, device=dev) * 0.02
    stats = torch.zeros(512, dtype=torch.float64, device=dev)
    bench("post proj_1 conv 25824x256x6144 +relu+stats", dict(A=big, B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048, bias=bias, act=1,
                                                             mask_period=807, mask_lo=3, mask_hi=803, colsum=stats, colsumsq=stats[256:]), 2.0 * R * 256 * 6144)
    dp1 = torch.randn(R + 64, 256, device=dev)
    Wd1 = torch.randn(3 * 256, 2048, device=dev) * 0.02
    dpool = torch.zeros(R, 2048, device=dev)
    bench("post proj_1 dgrad 25824x2048x768", dict(A=dp1, B=Wd1, C=dpool, M=R, N=2048, K=768, lda=256, ldb=2048, ldc=2048, ctap=256,
                                                  mask_period=807, mask_lo=3, mask_hi=803), 2.0 * R * 2048 * 768)
    dW1 = torch.zeros(6144, 256, device=dev)
    bench("post proj_1 wgrad 6144x256xK=25824", dict(A=big, B=dp1, C=dW1, M=6144, N=256, K=R, lda=2048, ldb=256, ldc=256, transA=1, ctap=2048, accumulate=1, split_k=8), 2.0 * R * 256 * 6144)
    x80 = torch.randn(R + 64, 80, device=dev)
    W8 = torch.randn(8 * 80, 256, device=dev) * 0.05
    bank = torch.zeros(R, 2048, device=dev)
    bench("post bank k=8 conv 25824x256x640 (C=80)", dict(A=x80, B=W8, C=bank, M=R, N=256, K=640, lda=80, ldb=256, ldc=2048, ctap=80, bias=bias, act=1,
                                                       mask_period=807, mask_lo=3, mask_hi=803), 2.0 * R * 256 * 640)
    post = torch.randn(25600, 512, device=dev)
    Wl = torch.randn(512, 1032, device=dev) * 0.05
    lin = torch.zeros(25600, 1032, device=dev)
    bl = torch.randn(1025, device=dev)
    bench("linear NN 25600x1025x512", dict(A=post, B=Wl, C=lin, M=25600, N=1025, K=512, lda=512, ldb=1032, ldc=1032, bias=bl), 2.0 * 25600 * 1025 * 512)
    drnn = torch.zeros(25600, 512, device=dev)
    bench("linear dgrad NT 25600x512x1025", dict(A=lin, B=Wl, C=drnn, M=25600, N=512, K=1025, lda=1032, ldb=1032, ldc=512, transB=1), 2.0 * 25600 * 1025 * 512)
    dWl = torch.zeros(512, 1025, device=dev)
    bench("linear wgrad 512x1025xK=25600 split", dict(A=post, B=lin, C=dWl, M=512, N=1025, K=25600, lda=512, ldb=1032, ldc=1025, transA=1, accumulate=1, split_k=8), 2.0 * 25600 * 1025 * 512)
    gx = torch.zeros(R, 1536, device=dev)
    Wx = torch.randn(256, 1536, device=dev) * 0.05
    bx = torch.randn(1536, device=dev)
    bench("gru x-side NN 25824x1536x256", dict(A=A, B=Wx, C=gx, M=R, N=1536, K=256, lda=256, ldb=1536, ldc=1536, bias=bx), 2.0 * R * 1536 * 256)
    bench("gru dgrad NT 25824x256x1536", dict(A=gx, B=Wx, C=Cc, M=R, N=256, K=1536, lda=1536, ldb=1536, ldc=256, transB=1), 2.0 * R * 1536 * 256)
    Re = 4576
    xe = torch.randn(Re + 64, 128, device=dev)
    We = torch.randn(16 * 128, 128, device=dev) * 0.05
    be = torch.randn(128, device=dev)
    banke = torch.zeros(Re, 2048, device=dev)
    bench("enc bank k=16 conv 4576x128x2048", dict(A=xe, B=We, C=banke, M=Re, N=128, K=2048, lda=128, ldb=128, ldc=2048, ctap=128, bias=be, act=1,
                                                 mask_period=143, mask_lo=7, mask_hi=135), 2.0 * Re * 128 * 2048)
    pe = torch.randn(Re + 64, 2048, device=dev)
    Wpe = torch.randn(3 * 2048, 128, device=dev) * 0.02
    p1e = torch.zeros(Re, 128, device=dev)
    bench("enc proj_1 conv 4576x128x6144", dict(A=pe, B=Wpe, C=p1e, M=Re, N=128, K=6144, lda=2048, ldb=128, ldc=128, ctap=2048, bias=be, act=1,
                                              mask_period=143, mask_lo=7, mask_hi=135), 2.0 * Re * 128 * 6144)


if __name__ == "__main__":
    main()
