"""Times representative taco_gemm problems of the C2 training step (CUDA events, L2-cold), TF32 tensor-core vs fp32 SIMT."""
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
flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)


def bench(name, kw, flops, precs=(1, 0), reps=5):
    d = capi.TacoGemmDesc()
    d.alpha = 1.0
    d.split_k = 1
    keep = []
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            keep.append(v)
            setattr(d, k, v.data_ptr())
        else:
            setattr(d, k, v)
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for prec in precs:
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
        out.append("%s %8.1f us %7.1f TFLOP/s" % ("tf32" if prec else "fp32", us, flops / us / 1e6))
    print("%-44s %s" % (name, " | ".join(out)), flush=True)


def main():
    R = 25824
    A = torch.randn(R + 64, 256, device=dev)
    W = torch.randn(256, 256, device=dev) * 0.05
    bias = torch.randn(256, device=dev)
    Cc = torch.zeros(R, 256, device=dev)
    bench("highway NN 25824x256x256 +bias+sigmoid", dict(A=A[32:], B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, bias=bias, act=2), 2.0 * R * 256 * 256)
    bench("dgrad NT 25824x256x256", dict(A=A[32:], B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, transB=1), 2.0 * R * 256 * 256)
    dW = torch.zeros(256, 256, device=dev)
    bench("wgrad TN 256x256xK=25824 split 64", dict(A=A[32:], B=Cc, C=dW, M=256, N=256, K=R, lda=256, ldb=256, ldc=256, transA=1, accumulate=1, split_k=64), 2.0 * R * 256 * 256)
    big = torch.randn(R + 64, 2048, device=dev)
    W1 = torch.randn(3 * 2048, 256, device=dev) * 0.02
    stats = torch.zeros(512, dtype=torch.float64, device=dev)
    bench("post proj_1 conv 25824x256x6144 +relu+stats", dict(A=big[31:], B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048, bias=bias, act=1,
                                                             mask_period=807, mask_lo=3, mask_hi=803, colsum=stats, colsumsq=stats[256:]), 2.0 * R * 256 * 6144)
    x80 = torch.randn(R + 64, 80, device=dev)
    W8 = torch.randn(8 * 80, 256, device=dev) * 0.05
    bench("post bank k=8 conv 25824x256x640 (C=80)", dict(A=x80[29:], B=W8, C=big[32:], M=R, N=256, K=640, lda=80, ldb=256, ldc=2048, ctap=80, bias=bias, act=1,
                                                       mask_period=807, mask_lo=3, mask_hi=803), 2.0 * R * 256 * 640)
    post = torch.randn(25600, 512, device=dev)
    Wl = torch.randn(512, 1028, device=dev) * 0.05
    lin = torch.zeros(25600, 1028, device=dev)
    bl = torch.randn(1025, device=dev)
    bench("linear NN 25600x1025x512", dict(A=post, B=Wl, C=lin, M=25600, N=1025, K=512, lda=512, ldb=1028, ldc=1028, bias=bl), 2.0 * 25600 * 1025 * 512)
    dWl = torch.zeros(512, 1025, device=dev)
    bench("linear wgrad 512x1025xK=25600 split 8", dict(A=post, B=lin, C=dWl, M=512, N=1025, K=25600, lda=512, ldb=1028, ldc=1025, transA=1, accumulate=1, split_k=8), 2.0 * 25600 * 1025 * 512)
    Re = 4576
    xe = torch.randn(Re + 64, 128, device=dev)
    We = torch.randn(16 * 128, 128, device=dev) * 0.05
    be = torch.randn(128, device=dev)
    bank = torch.zeros(Re, 2048, device=dev)
    bench("enc bank k=16 conv 4576x128x2048", dict(A=xe[25:], B=We, C=bank, M=Re, N=128, K=2048, lda=128, ldb=128, ldc=2048, ctap=128, bias=be, act=1,
                                                 mask_period=143, mask_lo=7, mask_hi=135), 2.0 * Re * 128 * 2048)
    a_seq = torch.randn(160, 128, device=dev)
    dctx = torch.randn(160, 256, device=dev)
    dmem = torch.zeros(128, 256, device=dev)
    bench("attn dmemory TN 128x256x160", dict(A=a_seq, B=dctx, C=dmem, M=128, N=256, K=160, lda=128, ldb=256, ldc=256, transA=1, accumulate=1), 2.0 * 128 * 256 * 160)


if __name__ == "__main__":
    main()
