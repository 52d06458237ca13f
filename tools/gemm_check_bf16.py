"""Operator-level check of the bf16-operand tcgen05 GEMM (taco_gemm, precision 2) against torch on the SAME bf16-rounded
operands (so the only differences are accumulation order and the bf16 rounding of mirror outputs): every addressing mode
the model uses, the bf16 mirror output, bf16-only output, tile widths 16..256, K / M / N remainders, multi-tile persistence."""
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
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device="cpu").manual_seed(0)
BF = torch.bfloat16


def rnd(*shape):
    return torch.randn(*shape, generator=g).to(dev)


def h(x):
    """bf16 mirror and its fp32 value"""
    x16 = x.to(BF).contiguous()
    return x16, x16.float()


def run(kw):
    d = capi.TacoGemmDesc()
    d.alpha = 1.0
    d.split_k = 1
    for k, v in kw.items():
        setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    capi.check(lib.taco_gemm(C.byref(d), 1, 2, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def report(name, got, ref, tol=2e-3):
    err = (got.float() - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1.0)
    flag = "ok" if err <= tol * scale and not torch.isnan(got.float()).any() else "FAIL"
    print("%-52s max|err| %.3e (ref max %.2e) %s" % (name, err, scale, flag), flush=True)
    return flag == "ok"


def main():
    ok = True
    # 1. NN + bias + relu over several tile widths / remainders; fp32 + bf16 mirror outputs
    for (M, N, K) in [(300, 200, 96), (128, 16, 64), (1000, 80, 200), (777, 256, 512), (4500, 1025 + 7, 136), (40000, 256, 256), (520, 400, 256)]:
        A, B, bias = rnd(M, K), rnd(K, N) * 0.2, rnd(N)
        A16, Af = h(A); B16, Bf = h(B)
        Cc = torch.full((M, N), 3.0, device=dev); C16 = torch.zeros(M, N, device=dev, dtype=BF)
        run(dict(A=A, B=B, A16=A16, B16=B16, C=Cc, C16=C16, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, bias=bias, act=1))
        ref = torch.relu(Af @ Bf + bias)
        ok &= report("NN bias relu %dx%dx%d" % (M, N, K), Cc, ref)
        ok &= report("  bf16 mirror", C16, ref, 8e-3)
    # 2. bf16-only output, sigmoid
    M, N, K = 2000, 256, 256
    A, B, bias = rnd(M, K), rnd(K, N) * 0.1, rnd(N)
    A16, Af = h(A); B16, Bf = h(B)
    C16 = torch.zeros(M, N, device=dev, dtype=BF)
    run(dict(A16=A16, B16=B16, C16=C16, M=M, N=N, K=K, lda=K, ldb=N, ldc=N, bias=bias, act=2))
    ok &= report("NN bf16-only out sigmoid", C16, torch.sigmoid(Af @ Bf + bias), 8e-3)
    # 3. transB (data gradient form), accumulate=1
    M, N, K = 3000, 200, 1025 + 7
    A, Bt = rnd(M, K), rnd(N, K) * 0.1
    A16, Af = h(A); B16, Bf = h(Bt)
    C0 = rnd(M, N); Cc = C0.clone()
    run(dict(A=A, B=Bt, A16=A16, B16=B16, C=Cc, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, transB=1, accumulate=1))
    ok &= report("NT accumulate 3000x200x1032", Cc, C0 + Af @ Bf.t())
    # 4. transA split-K weight gradient (atomic accumulate)
    for (R, Kin, Nout, sp) in [(5000, 256, 384, 8), (25824, 6144 // 8, 256, 64), (1000, 80, 256, 4), (4576, 128, 1536, 16)]:
        X, dY = rnd(R, Kin), rnd(R, Nout) * 0.1
        X16, Xf = h(X); dY16, dYf = h(dY)
        dW = torch.zeros(Kin, Nout, device=dev)
        run(dict(A=X, B=dY, A16=X16, B16=dY16, C=dW, M=Kin, N=Nout, K=R, lda=Kin, ldb=Nout, ldc=Nout, transA=1, accumulate=1, split_k=sp))
        ok &= report("TN split-K wgrad %dx%dx%d" % (Kin, Nout, R), dW, Xf.t() @ dYf, 3e-3)
    # 5. conv as implicit GEMM (ctap == lda == 128), k=5, mask + stats
    Nb, T, Cin, Cout, k = 3, 137, 128, 128, 5
    Kb = 8; PL = (Kb - 1) // 2; Tp = T + Kb - 1; rows = Nb * Tp; slack = Kb
    xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
    x = rnd(Nb, T, Cin)
    xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
    W, bias = rnd(k, Cin, Cout) * 0.1, rnd(Cout)
    x16, xf = h(xp_full); W16, Wf = h(W)
    l = (k - 1) // 2
    out = torch.full((rows, Cout), 7.0, device=dev); out16 = torch.zeros(rows, Cout, device=dev, dtype=BF)
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
    run(dict(A16=x16[slack - l:], B16=W16, A=xp_full[slack - l:], B=W, C=out, C16=out16, M=rows, N=Cout, K=k * Cin, lda=Cin, ldb=Cout, ldc=Cout, ctap=Cin, bias=bias, act=1,
             mask_period=Tp, mask_lo=PL, mask_hi=PL + T, colsum=stats, colsumsq=stats[Cout:]))
    xv = xf[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T]
    ref = torch.relu(torch.nn.functional.conv1d(torch.nn.functional.pad(xv.transpose(1, 2), (l, k - 1 - l)),
                                                Wf.permute(2, 1, 0).contiguous(), bias).transpose(1, 2))
    got = out.view(Nb, Tp, Cout)
    ok &= report("conv k=5 C=128 (valid rows)", got[:, PL:PL + T], ref)
    ok &= report("conv pad rows are zero", got[:, :PL].abs().max().view(1), torch.zeros(1, device=dev))
    ok &= report("conv bf16 mirror", out16.view(Nb, Tp, Cout)[:, PL:PL + T], ref, 8e-3)
    ok &= report("conv column sums", stats[:Cout].float(), ref.sum((0, 1)), 2e-3)
    ok &= report("conv column sumsq", stats[Cout:].float(), (ref * ref).sum((0, 1)), 2e-3)
    # 6. conv with 80 input channels (contiguous im2col rows, overlapping TMA rows), k=4 and k=1
    for k in (4, 1, 8):
        Cin, Cout = 80, 256
        xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
        x = rnd(Nb, T, Cin)
        xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
        W, bias = rnd(k, Cin, Cout) * 0.1, rnd(Cout)
        x16, xf = h(xp_full); W16, Wf = h(W)
        l = (k - 1) // 2
        out = torch.zeros(rows, Cout, device=dev)
        run(dict(A16=x16[slack - l:], B16=W16, C=out, M=rows, N=Cout, K=k * Cin, lda=Cin, ldb=Cout, ldc=Cout, ctap=Cin, bias=bias,
                 mask_period=Tp, mask_lo=PL, mask_hi=PL + T))
        xv = xf[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T]
        ref = torch.nn.functional.conv1d(torch.nn.functional.pad(xv.transpose(1, 2), (l, k - 1 - l)), Wf.permute(2, 1, 0).contiguous(), bias).transpose(1, 2)
        ok &= report("conv k=%d C=80" % k, out.view(Nb, Tp, Cout)[:, PL:PL + T], ref)
    # 7. conv k=3 over 2048 channels (proj_1 form: tap-inner order), N=256
    Cin, Cout, k = 2048, 256, 3
    xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
    x = rnd(Nb, T, Cin)
    xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
    W, bias = rnd(k, Cin, Cout) * 0.02, rnd(Cout)
    x16, xf = h(xp_full); W16, Wf = h(W)
    l = 1
    out = torch.zeros(rows, Cout, device=dev)
    run(dict(A16=x16[slack - l:], B16=W16, C=out, M=rows, N=Cout, K=k * Cin, lda=Cin, ldb=Cout, ldc=Cout, ctap=Cin, bias=bias, act=1,
             mask_period=Tp, mask_lo=PL, mask_hi=PL + T))
    xv = xf[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T]
    ref = torch.relu(torch.nn.functional.conv1d(torch.nn.functional.pad(xv.transpose(1, 2), (1, 1)), Wf.permute(2, 1, 0).contiguous(), bias).transpose(1, 2))
    ok &= report("conv k=3 C=2048 (tap-inner)", out.view(Nb, Tp, Cout)[:, PL:PL + T], ref)
    # 8. strided taps: A is a 256-wide column slice of a 1024-wide buffer (dgrad form), k=3
    Cw, Cs, Cn, k = 1024, 256, 128, 3
    buf = torch.zeros(rows + 2 * slack, Cw, device=dev)
    dy = rnd(Nb, T, Cs)
    buf[slack:slack + rows].view(Nb, Tp, Cw)[:, PL:PL + T, 512:512 + Cs] = dy
    Wd = rnd(k * Cs, Cn) * 0.1
    buf16, buff = h(buf); Wd16, Wdf = h(Wd)
    r = k - 1 - (k - 1) // 2
    out = torch.zeros(rows, Cn, device=dev)
    run(dict(A16=buf16[slack - r:, 512:], B16=Wd16, C=out, M=rows, N=Cn, K=k * Cs, lda=Cw, ldb=Cn, ldc=Cn, ctap=Cs,
             mask_period=Tp, mask_lo=PL, mask_hi=PL + T))
    dyp = torch.zeros(Nb, Tp + 2 * slack, Cs, device=dev)
    dyp[:, slack + PL:slack + PL + T] = buff[slack:slack + rows].view(Nb, Tp, Cw)[:, PL:PL + T, 512:512 + Cs]
    ref = torch.zeros(Nb, T, Cn, device=dev)
    for j in range(k):
        ref += dyp[:, slack + PL - r + j: slack + PL - r + j + T] @ Wdf[j * Cs:(j + 1) * Cs]
    ok &= report("strided-tap dgrad", out.view(Nb, Tp, Cn)[:, PL:PL + T], ref)
    # 9. conv weight gradient: transA + taps + split-K
    Cin, Cout, k = 128, 256, 3
    xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
    x = rnd(Nb, T, Cin)
    xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
    dyb = torch.zeros(rows, Cout, device=dev)
    dy = rnd(Nb, T, Cout)
    dyb.view(Nb, Tp, Cout)[:, PL:PL + T] = dy
    x16, xf = h(xp_full); dy16, dyf = h(dyb)
    l = 1
    dW = torch.zeros(k * Cin, Cout, device=dev)
    run(dict(A16=x16[slack - l:], B16=dy16, C=dW, M=k * Cin, N=Cout, K=rows, lda=Cin, ldb=Cout, ldc=Cout, transA=1, ctap=Cin,
             accumulate=1, split_k=3))
    xpad = torch.nn.functional.pad(xf[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T], (0, 0, l, k - 1 - l))
    dyv = dyf.view(Nb, Tp, Cout)[:, PL:PL + T]
    ref = torch.stack([torch.einsum("ntc,nto->co", xpad[:, j:j + T], dyv) for j in range(k)]).reshape(k * Cin, Cout)
    ok &= report("conv wgrad k=3 (TN taps split-K)", dW, ref, 3e-3)
    # 9b. conv weight gradient with 80 channels (overlapping rows, transA)
    Cin, Cout, k = 80, 256, 4
    xp_full = torch.zeros(rows + 2 * slack, Cin, device=dev)
    x = rnd(Nb, T, Cin)
    xp_full[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T] = x
    x16, xf = h(xp_full)
    l = (k - 1) // 2
    dW = torch.zeros(k * Cin, Cout, device=dev)
    run(dict(A16=x16[slack - l:], B16=dy16, C=dW, M=k * Cin, N=Cout, K=rows, lda=Cin, ldb=Cout, ldc=Cout, transA=1, ctap=Cin,
             accumulate=1, split_k=3))
    xpad = torch.nn.functional.pad(xf[slack:slack + rows].view(Nb, Tp, Cin)[:, PL:PL + T], (0, 0, l, k - 1 - l))
    ref = torch.stack([torch.einsum("ntc,nto->co", xpad[:, j:j + T], dyv) for j in range(k)]).reshape(k * Cin, Cout)
    ok &= report("conv wgrad k=4 C=80 (TN overlapping rows)", dW, ref, 3e-3)
    # 10. row remap (mel projection into the padded layout), fp32 + bf16 mirror
    Nb2, Td, Y, MR, Tp2, PL2 = 2, 9, 256, 400, 45 + 7, 3
    y2, Wm, bm = rnd(Nb2 * Td, Y), rnd(Y, MR) * 0.1, rnd(MR)
    y16, yf = h(y2); Wm16, Wmf = h(Wm)
    dst = torch.zeros(Nb2 * Tp2 * 80, device=dev); dst16 = torch.zeros(Nb2 * Tp2 * 80, device=dev, dtype=BF)
    run(dict(A16=y16, B16=Wm16, C=dst[PL2 * 80:], C16=dst16[PL2 * 80:], M=Nb2 * Td, N=MR, K=Y, lda=Y, ldb=MR, ldc=MR, bias=bm, remap_period=Td,
             remap_outer=Tp2 * 80, remap_inner=MR))
    ref = (yf @ Wmf + bm).view(Nb2, Td * 5, 80)
    ok &= report("row remap", dst.view(Nb2, Tp2, 80)[:, PL2:PL2 + Td * 5], ref)
    ok &= report("row remap bf16 mirror", dst16.view(Nb2, Tp2, 80)[:, PL2:PL2 + Td * 5], ref, 8e-3)
    # 11. tap table (conv-bank data gradient form): members k=1..4 of width Cb=128 over a [rows, 4*128] gradient
    Kbk, Cb, Cn = 4, 128, 128
    KC = Kbk * Cb
    dbank = torch.zeros(rows + 2 * slack, KC, device=dev)
    dv = rnd(Nb, T, KC) * 0.5
    dbank[slack:slack + rows].view(Nb, Tp, KC)[:, PL:PL + T] = dv
    d16, df = h(dbank)
    KK = Cb * Kbk * (Kbk + 1) // 2
    Wd = rnd(KK, Cn) * 0.1
    Wd16, Wdf = h(Wd)
    tab = []
    for kk in range(1, Kbk + 1):
        l = (kk - 1) // 2; r = kk - 1 - l
        for j in range(kk):
            for qq in range(Cb // 64):
                tab += [(kk - 1) * Cb + 64 * qq, j - r + slack]
    tabt = torch.tensor(tab, dtype=torch.int32, device=dev)
    out = torch.zeros(rows, Cn, device=dev)
    run(dict(A16=d16, B16=Wd16, C=out, M=rows, N=Cn, K=KK, lda=KC, ldb=Cn, ldc=Cn, tap_table=tabt, tap_rows=rows + 2 * slack,
             mask_period=Tp, mask_lo=PL, mask_hi=PL + T))
    dfull = torch.zeros(Nb, Tp + 2 * slack, KC, device=dev)
    dfull[:, slack:slack + Tp] = df[slack:slack + rows].view(Nb, Tp, KC)
    ref = torch.zeros(Nb, T, Cn, device=dev)
    off = 0
    for kk in range(1, Kbk + 1):
        l = (kk - 1) // 2; r = kk - 1 - l
        for j in range(kk):
            ref += dfull[:, slack + PL - r + j: slack + PL - r + j + T, (kk - 1) * Cb:kk * Cb] @ Wdf[off:off + Cb]
            off += Cb
    ok &= report("tap-table bank dgrad", out.view(Nb, Tp, Cn)[:, PL:PL + T], ref)
    # 12. many launches back to back (scheduler slots are self-cleaning)
    M, N, K = 3000, 256, 128
    A, B = rnd(M, K), rnd(K, N) * 0.1
    A16, Af = h(A); B16, Bf = h(B)
    Cc = torch.zeros(M, N, device=dev)
    for _ in range(600):
        d = dict(A16=A16, B16=B16, C=Cc, M=M, N=N, K=K, lda=K, ldb=N, ldc=N)
        dd = capi.TacoGemmDesc(); dd.alpha = 1.0; dd.split_k = 1
        for kq, v in d.items():
            setattr(dd, kq, v.data_ptr() if isinstance(v, torch.Tensor) else v)
        capi.check(lib.taco_gemm(C.byref(dd), 1, 2, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ok &= report("600 launches back to back", Cc, Af @ Bf)
    print("launches", lib.taco_launch_count())
    print("GEMM_CHECK_BF16", "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()
