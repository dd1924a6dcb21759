"""Isolated time of the DiT GEMM shapes per pair mode (rap_test_gemm: epi | 0x200 plain, | 0x100 weight-tile multicast, | 0x300 2-CTA UMMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rap_b200 import _lib
import bench
T = 8192
dev = torch.device("cuda")
st = torch.cuda.current_stream().cuda_stream
for name, (epi, K, N) in {"qkv 512->1536": (2, 512, 1536), "out 512->512": (0, 512, 512), "ff1 512->4096": (3, 512, 4096), "ff2 2048->512": (0, 2048, 512)}.items():
    a = torch.randn(T, K, device=dev, dtype=torch.float16); w = torch.randn(N, K, device=dev, dtype=torch.float16) / K ** 0.5
    bias = torch.randn(N, device=dev); gamma = torch.ones(1024, device=dev)
    h = torch.randn(T, N, device=dev) if epi == 0 else None
    o32 = torch.empty(T, N, device=dev) if epi == 0 else None
    o16 = torch.empty(T, N // 2 if epi == 3 else N, device=dev, dtype=torch.float16) if epi != 0 else None
    row = []
    for mode in (0x200, 0x100, 0x300, 0x200, 0x300):
        fn = lambda: _lib.check(_lib.lib.rap_test_gemm(epi | mode, 256, T, N, K, a.data_ptr(), w.data_ptr(), bias.data_ptr() if epi != 2 else None, _lib.ptr(h), _lib.ptr(o32),
                                                       _lib.ptr(o16), gamma.data_ptr() if epi == 2 else None, None, st), "gemm")
        ms = bench.timed(fn, 50, 5)
        row.append("%s %.1f us (%.0f TFLOP/s)" % ({0x200: "plain", 0x100: "multicast", 0x300: "2-CTA"}[mode], ms * 1e3, 2.0 * T * N * K / ms / 1e9))
    if epi == 2:          # QKV epilogue: packed f32x2 RMS-norm against the scalar form (| 0x400), alternating
        for e in (2, 2 | 0x400, 2 | 0x800, 2 | 0x1800, 2, 2 | 0x400, 2 | 0x800, 2 | 0x1800):
            fn = lambda: _lib.check(_lib.lib.rap_test_gemm(e | 0x200, 256, T, N, K, a.data_ptr(), w.data_ptr(), None, None, None, o16.data_ptr(), gamma.data_ptr(), None, st), "gemm")
            ms = bench.timed(fn, 50, 5)
            row.append("%s %.1f us" % ({2: "packed norm", 2 | 0x400: "scalar norm", 2 | 0x800: "packed + bulk stores", 2 | 0x1800: "+ column-major tiles"}[e], ms * 1e3))
    if epi == 3:          # GEGLU epilogue: packed f32x2 erfc = 2^poly form against the scalar Abramowitz-Stegun one (| 0x400), alternating
        for e in (3, 3 | 0x400, 3 | 0x800, 3 | 0x1800, 3, 3 | 0x400, 3 | 0x800, 3 | 0x1800):
            fn = lambda: _lib.check(_lib.lib.rap_test_gemm(e | 0x200, 256, T, N, K, a.data_ptr(), w.data_ptr(), bias.data_ptr(), None, None, o16.data_ptr(), None, None, st), "gemm")
            ms = bench.timed(fn, 50, 5)
            row.append("%s %.1f us" % ({3: "packed gelu", 3 | 0x400: "scalar gelu", 3 | 0x800: "packed + bulk stores", 3 | 0x1800: "+ column-major tiles"}[e], ms * 1e3))
    if epi == 0:          # in-place residual form: register epilogue (addend = out) against the bulk-reduction epilogue (EPI_F32_RED = 5), alternating
        for e in (0x200, 5, 0x200, 5):
            fn = lambda: _lib.check(_lib.lib.rap_test_gemm(e, 256, T, N, K, a.data_ptr(), w.data_ptr(), bias.data_ptr(), h.data_ptr(), h.data_ptr(), None, None, None, st), "gemm")
            ms = bench.timed(fn, 50, 5)
            row.append("%s %.1f us" % ("in-place" if e != 5 else "bulk-reduction", ms * 1e3))
    print(name + ": " + " | ".join(row), flush=True)
