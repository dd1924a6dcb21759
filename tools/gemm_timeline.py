"""clock64() timeline of GEMM CTA 0 per tile (debug aid): producer span, MMA wait/issue, epilogue wait/duration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rap_b200 import _lib
def run(name, epi, bn, M, N, K):
    A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / 22).half()
    bias = torch.randn(N, device="cuda"); add = torch.randn(M, N, device="cuda") if epi == 0 else None
    gamma = torch.ones(2, 8, 64, device="cuda")
    out32 = torch.zeros(M, N, device="cuda") if epi == 0 else None
    out16 = torch.zeros(M, N, device="cuda", dtype=torch.float16)
    dbg = torch.zeros(16, 8, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(4):
        if it == 3: ev[0].record()
        _lib.check(_lib.lib.rap_test_gemm(epi, bn, M, N, K, A.data_ptr(), W.data_ptr(), bias.data_ptr(), _lib.ptr(add), _lib.ptr(out32), out16.data_ptr(),
                                          gamma.data_ptr(), dbg.data_ptr(), _lib.stream_ptr()))
    ev[1].record(); torch.cuda.synchronize()
    d = dbg.cpu().numpy(); t0 = d[0][0]
    print(f"== {name}: M={M} N={N} K={K} BN={bn}: {ev[0].elapsed_time(ev[1])*1e3:.1f} us, {2*M*N*K/ev[0].elapsed_time(ev[1])/1e9:.0f} TFLOP/s")
    for t in range(2 if epi == 0 else 4):
        r = d[t]
        if r[0] == 0: break
        if t == 0 and epi == 0:
            c = d[2:4].reshape(-1)
            print("  epilogue chunks (warp 2): " + " | ".join(f"c{k}: ld {c[4*k+1]-c[4*k]} stage {c[4*k+2]-c[4*k+1]} add+store {c[4*k+3]-c[4*k+2]} (start {c[4*k]-r[6]})" for k in range(4) if c[4*k]))
        print(f"  tile{t}: prod [{r[0]-t0},{r[1]-t0}]  mma wait_acc {r[3]-r[2]} issue_span {r[4]-r[3]} (start {r[2]-t0})  epi start {r[5]-t0} acc_ready {r[6]-t0} done {r[7]-t0} (epi work {r[7]-r[6]})")
run("outproj f32", 0, 256, 8192, 512, 512)
run("ff2 f32", 0, 256, 8192, 512, 2048)
run("outproj f32 bn128", 0, 128, 8192, 512, 512)
run("ff2 f32 bn128", 0, 128, 8192, 512, 2048)
run("qkv rmsnorm", 2, 256, 8192, 1536, 512)
run("geglu", 3, 256, 8192, 4096, 512)
