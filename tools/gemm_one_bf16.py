"""One bf16 GEMM problem a few times (for ncu --set full): python tools/gemm_one_bf16.py [proj1|nn36|wgrad]"""
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
BF = torch.bfloat16
what = sys.argv[1] if len(sys.argv) > 1 else "proj1"
R = 25824
if what == "proj1":
    big = torch.randn(R + 64, 2048, device=dev); W1 = torch.randn(3 * 2048, 256, device=dev) * 0.02
    bias = torch.randn(256, device=dev); Cc = torch.zeros(R, 256, device=dev)
    kw = dict(A=big, B=W1, C=Cc, M=R, N=256, K=6144, lda=2048, ldb=256, ldc=256, ctap=2048, bias=bias, act=1, mask_period=807, mask_lo=3, mask_hi=803)
elif what == "nn36":
    A2 = torch.randn(4576, 6144, device=dev); W = torch.randn(6144, 128, device=dev) * 0.02; Cc = torch.zeros(4576, 128, device=dev)
    kw = dict(A=A2, B=W, C=Cc, M=4576, N=128, K=6144, lda=6144, ldb=128, ldc=128)
else:
    big = torch.randn(R + 64, 2048, device=dev); dp1 = torch.randn(R + 64, 256, device=dev); dW1 = torch.zeros(6144, 256, device=dev)
    kw = dict(A=big, B=dp1, C=dW1, M=6144, N=256, K=R, lda=2048, ldb=256, ldc=256, transA=1, ctap=2048, accumulate=1, split_k=8)
d = capi.TacoGemmDesc(); d.alpha = 1.0; d.split_k = 1
keep = []
for k, v in kw.items():
    if isinstance(v, torch.Tensor):
        if k in ("A", "B"):
            v16 = v.to(BF); keep.append(v16); setattr(d, k + "16", v16.data_ptr())
        setattr(d, k, v.data_ptr())
    else:
        setattr(d, k, v)
for _ in range(4):
    capi.check(lib.taco_gemm(C.byref(d), 1, 2, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("done")
