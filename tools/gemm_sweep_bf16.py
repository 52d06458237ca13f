"""Env-controlled sweep over a few K-heavy C2 problems of the bf16 GEMM (run once per setting of TACO_BF16_STAGES /
TACO_BF16_TAPG / TACO_BF16_BN): median of 7 L2-cold and 7 L2-warm launches each."""
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


def bench(name, kw, flops, prec=2):
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
        capi.check(lib.taco_gemm(C.byref(d), 1, prec, st))
    res = []
    for cold in (True, False):
        ts = []
        for _ in range(7):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); capi.check(lib.taco_gemm(C.byref(d), 1, prec, st)); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res.append(ts[3])
    print("%-40s cold %7.1f us %6.1f TF/s | warm %7.1f us %6.1f TF/s" % (name, res[0], flops / res[0] / 1e6, res[1], flops / res[1] / 1e6), flush=True)


def main():
    prec = int(os.environ.get("SWEEP_PREC", "2"))
    print("settings: STAGES=%s TAPG=%s BN=%s prec=%d" % (os.environ.get("TACO_BF16_STAGES"), os.environ.get("TACO_BF16_TAPG"), os.environ.get("TACO_BF16_BN"), prec))
    R = 25824
    bias = torch.randn(256, device=dev)
    Cc = torch.zeros(R, 256, device=dev)
    big = torch.randn(R + 64, 2048, device=dev); W1 = torch.randn(3 * 2048, 256, device=dev) * 0.02
    bench("post proj_1 conv 25824x256x6144", dict(A=big, B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048, bias=bias, act=1,
                                                  mask_period=807, mask_lo=3, mask_hi=803), 2.0 * R * 256 * 6144, prec)
    dp1 = torch.randn(R + 64, 256, device=dev)
    dW1 = torch.zeros(6144, 256, device=dev)
    bench("post proj_1 wgrad 6144x256xK=25824", dict(A=big, B=dp1, C=dW1, M=6144, N=256, K=R, lda=2048, ldb=256, ldc=256, transA=1, ctap=2048, accumulate=1, split_k=8), 2.0 * R * 256 * 6144, prec)
    Re = 4576
    be = torch.randn(128, device=dev)
    pe = torch.randn(Re + 64, 2048, device=dev); Wpe = torch.randn(3 * 2048, 128, device=dev) * 0.02
    p1e = torch.zeros(Re, 128, device=dev)
    bench("enc proj_1 conv 4576x128x6144", dict(A=pe, B=Wpe, C=p1e, M=Re, N=128, K=6144, lda=2048, ldb=128, ldc=128, ctap=2048, bias=be, act=1,
                                                mask_period=143, mask_lo=7, mask_hi=135), 2.0 * Re * 128 * 6144, prec)
    A2 = torch.randn(Re, 6144, device=dev)
    bench("plain NN 4576x128x6144 (no taps)", dict(A=A2, B=Wpe, C=p1e, M=Re, N=128, K=6144, lda=6144, ldb=128, ldc=128, bias=be, act=1), 2.0 * Re * 128 * 6144, prec)
    Wt = torch.randn(128, 6144, device=dev) * 0.02
    bench("plain NT 4576x128x6144 (B K-major)", dict(A=A2, B=Wt, C=p1e, M=Re, N=128, K=6144, lda=6144, ldb=6144, ldc=128, transB=1, bias=be, act=1), 2.0 * Re * 128 * 6144, prec)


if __name__ == "__main__":
    main()
