import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import tacotron_b200
from importlib import import_module
capi = import_module("multi-speaker-tacotron-tensorflow_b200.capi"); lib = capi.load()
lib.taco_debug_timeline.argtypes = [C.POINTER(C.c_ulonglong)]
dev = torch.device("cuda", 0)
def run(name, kw):
    d = capi.TacoGemmDesc(); d.alpha = 1.0; d.split_k = 1
    for k, v in kw.items(): setattr(d, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    for _ in range(3):
        capi.check(lib.taco_gemm(C.byref(d), 1, 1, torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)(); lib.taco_debug_timeline(buf)
    t = list(buf); print(name, " ".join("%d:%+.1fus" % (i, (t[i] - t[0]) / 1e3) for i in (1, 3, 4, 5, 6, 7)), flush=True)
R = 25824
A = torch.randn(R + 64, 256, device=dev); W = torch.randn(256, 256, device=dev); Cc = torch.zeros(R, 256, device=dev); bias = torch.randn(256, device=dev)
run("NN K=256 bias sigmoid", dict(A=A[32:], B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, bias=bias, act=2))
run("NN K=256 plain       ", dict(A=A[32:], B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256))
run("NT K=256 plain       ", dict(A=A[32:], B=W, C=Cc, M=R, N=256, K=256, lda=256, ldb=256, ldc=256, transB=1))
a_seq = torch.randn(160, 128, device=dev); dctx = torch.randn(160, 256, device=dev); dmem = torch.zeros(128, 256, device=dev)
run("TN tiny 128x256x160  ", dict(A=a_seq, B=dctx, C=dmem, M=128, N=256, K=160, lda=128, ldb=256, ldc=256, transA=1, accumulate=1))
big = torch.randn(R + 64, 2048, device=dev); W1 = torch.randn(3 * 2048, 256, device=dev)
run("conv K=6144          ", dict(A=big[31:], B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048))
