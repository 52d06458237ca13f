"""Operator-level check of taco_gemm (both the fp32 SIMT and the tcgen05/TF32 kernels) against torch, every addressing mode."""
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
torch.backends.cudnn.allow_tf32 = False          # the torch reference must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape):
    return torch.randn(*shape, generator=g).to(dev)


def run(desc_kwargs, prec):
    d = capi.TacoGemmDesc()
    d.alpha = 1.0
    d.split_k = 1
    for k, v in desc_kwargs.items():
        setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    capi.check(lib.taco_gemm(C.byref(d), 1, prec, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def report(name, got, ref, prec):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-5 if prec == 0 else 4e-3) * max(scale, 1.0)
    flag = "ok" if err <= tol and not torch.isnan(got).any() else "FAIL"
    print("%-44s prec=%d max|err| %.3e (ref max %.2e) %s" % (name, prec, err, scale, flag), flush=True)
    return flag == "ok"


def main():
    ok = True
    for prec in (0, 1):
        # 1. plain NN + bias + relu, ragged sizes
        M, N, K = 300, 200, 96
        A, B, bias = rnd(M, K), rnd(K, N), rnd(N)
        Cc = torch.zeros(M, N, device=dev)
        run(dict(A=A, B=B, C=Cc, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, bias=bias, act=1), prec)
        ok &= report("NN bias relu 300x200x96", Cc, torch.relu(A @ B + bias), prec)
        # 2. transB (data gradient form), accumulate=1
        Bt = rnd(N, K)
        C0 = rnd(M, N); Cc = C0.clone()
        run(dict(A=A, B=Bt, C=Cc, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, transB=1, accumulate=1), prec)
        ok &= report("NT accumulate", Cc, C0 + A @ Bt.t(), prec)
        # 3. transA split-K weight gradient (atomic accumulate)
        R, Kin, Nout = 5000, 256, 384
        X, dY = rnd(R, Kin), rnd(R, Nout) * 0.1
        dW = torch.zeros(Kin, Nout, device=dev)
        run(dict(A=X, B=dY, C=dW, M=Kin, N=Nout, K=R, lda=Kin, ldb=Nout, ldc=Nout, transA=1, accumulate=1, split_k=8), prec)
        ok &= report("TN split-K wgrad 256x384x5000", dW, X.t() @ dY, prec)
        # 4. conv as implicit GEMM (ctap == lda == 128), k=5, mask + stats
        Nb, T, Cin, Cout, k = 3, 37, 128, 128, 5
        Kb = 8; PL = (Kb - 1) // 2; Tp = T + Kb - 1; rows = Nb * Tp; slack = Kb
        xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
        x = rnd(Nb, T, Cin)
        xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
        W, bias = rnd(k, Cin, Cout) * 0.1, rnd(Cout)
        l = (k - 1) // 2
        out = torch.full((rows, Cout), 7.0, device=dev)
        stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
        A_ptr = xp_full[slack - l:]
        run(dict(A=A_ptr, B=W, C=out, M=rows, N=Cout, K=k * Cin, lda=Cin, ldb=Cout, ldc=Cout, ctap=Cin, bias=bias, act=1,
                 mask_period=Tp, mask_lo=PL, mask_hi=PL + T, colsum=stats, colsumsq=stats[Cout:]), prec)
        ref = torch.relu(torch.nn.functional.conv1d(torch.nn.functional.pad(x.transpose(1, 2), (l, k - 1 - l)),
                                                    W.permute(2, 1, 0).contiguous(), bias).transpose(1, 2))
        got = out.view(Nb, Tp, Cout)
        ok &= report("conv k=5 C=128 (valid rows)", got[:, PL:PL + T], ref, prec)
        ok &= report("conv pad rows are zero", got[:, :PL].abs().max().view(1), torch.zeros(1, device=dev), prec)
        ok &= report("conv column sums", stats[:Cout].float(), ref.sum((0, 1)), prec)
        ok &= report("conv column sumsq", stats[Cout:].float(), (ref * ref).sum((0, 1)), prec)
        # 5. conv with 80 input channels (contiguous im2col rows, overlapping TMA rows), k=4
        Cin, Cout, k = 80, 256, 4
        xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
        x = rnd(Nb, T, Cin)
        xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
        W, bias = rnd(k, Cin, Cout) * 0.1, rnd(Cout)
        l = (k - 1) // 2
        out = torch.zeros(rows, Cout, device=dev)
        run(dict(A=xp_full[slack - l:], B=W, C=out, M=rows, N=Cout, K=k * Cin, lda=Cin, ldb=Cout, ldc=Cout, ctap=Cin, bias=bias,
                 mask_period=Tp, mask_lo=PL, mask_hi=PL + T), prec)
        ref = torch.nn.functional.conv1d(torch.nn.functional.pad(x.transpose(1, 2), (l, k - 1 - l)), W.permute(2, 1, 0).contiguous(), bias).transpose(1, 2)
        ok &= report("conv k=4 C=80", out.view(Nb, Tp, Cout)[:, PL:PL + T], ref, prec)
        # 6. strided taps: A is a 256-wide column slice of a 1024-wide buffer (bank data-gradient form), k=3
        Cw, Cs, Cn, k = 1024, 256, 128, 3
        buf = torch.zeros(rows + 2 * slack, Cw, device=dev)
        dy = rnd(Nb, T, Cs)
        buf[slack:slack + rows].view(Nb, Tp, Cw)[:, PL:PL + T, 512:512 + Cs] = dy
        Wd = rnd(k * Cs, Cn) * 0.1
        r = k - 1 - (k - 1) // 2
        out = torch.zeros(rows, Cn, device=dev)
        A_ptr = buf[slack - r:, 512:]
        run(dict(A=A_ptr, B=Wd, C=out, M=rows, N=Cn, K=k * Cs, lda=Cw, ldb=Cn, ldc=Cn, ctap=Cs, accumulate=2,
                 mask_period=Tp, mask_lo=PL, mask_hi=PL + T), prec)
        dyp = torch.zeros(Nb, Tp + 2 * slack, Cs, device=dev)
        dyp[:, slack + PL:slack + PL + T] = dy
        ref = torch.zeros(Nb, T, Cn, device=dev)
        for j in range(k):
            ref += dyp[:, slack + PL - r + j: slack + PL - r + j + T] @ Wd[j * Cs:(j + 1) * Cs]
        ok &= report("strided-tap dgrad (atomic)", out.view(Nb, Tp, Cn)[:, PL:PL + T], ref, prec)
        # 7. conv weight gradient: transA + taps + split-K
        Cin, Cout, k = 128, 256, 3
        xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
        x = rnd(Nb, T, Cin)
        xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
        dyb = torch.zeros(rows, Cout, device=dev)
        dy = rnd(Nb, T, Cout)
        dyb.view(Nb, Tp, Cout)[:, PL:PL + T] = dy
        l = (k - 1) // 2
        dW = torch.zeros(k * Cin, Cout, device=dev)
        run(dict(A=xp_full[slack - l:], B=dyb, C=dW, M=k * Cin, N=Cout, K=rows, lda=Cin, ldb=Cout, ldc=Cout, transA=1, ctap=Cin,
                 accumulate=1, split_k=3), prec)
        xpad = torch.nn.functional.pad(x, (0, 0, l, k - 1 - l))
        ref = torch.stack([torch.einsum("ntc,nto->co", xpad[:, j:j + T], dy) for j in range(k)]).reshape(k * Cin, Cout)
        ok &= report("conv wgrad k=3 (TN taps split-K)", dW, ref, prec)
        # 8. row remap (mel projection into the padded layout)
        Nb2, Td, Y, MR, Tp2, PL2 = 2, 9, 256, 400, 45 + 7, 3
        y2, Wm, bm = rnd(Nb2 * Td, Y), rnd(Y, MR) * 0.1, rnd(MR)
        dst = torch.zeros(Nb2 * Tp2 * 80, device=dev)
        run(dict(A=y2, B=Wm, C=dst[PL2 * 80:], M=Nb2 * Td, N=MR, K=Y, lda=Y, ldb=MR, ldc=MR, bias=bm, remap_period=Td,
                 remap_outer=Tp2 * 80, remap_inner=MR), prec)
        ref = (y2 @ Wm + bm).view(Nb2, Td * 5, 80)
        ok &= report("row remap", dst.view(Nb2, Tp2, 80)[:, PL2:PL2 + Td * 5], ref, prec)
        # 9. wide N with odd leading dimension falls back to SIMT even in TF32 mode
        M, N, K = 257, 1025, 512
        A, B, bias = rnd(M, K), rnd(K, N) * 0.1, rnd(N)
        Cc = torch.zeros(M, N, device=dev)
        run(dict(A=A, B=B, C=Cc, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, bias=bias), prec)
        ok &= report("NN N=1025 (ldb odd)", Cc, A @ B + bias, prec)
    print("launches", lib.taco_launch_count())
    print("GEMM_CHECK", "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()
