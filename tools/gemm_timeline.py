"""GEMM timelines at speed: a kernel is launched back to back for SECS seconds (clocks ramped up, every launch stamps the same debug buffer, the
last launch's stamps survive); printed per shape: the back-to-back time per launch (CUDA events over the last 200 launches), the SM clock
(clock64 / globaltimer of the CTAs), CTA 0's clock64 timeline per tile (producer span, MMA wait / issue, epilogue), and over all CTAs the
distribution of entry and exit times (globaltimer, relative to the first entry) -- kernel span = last exit - first entry; what the
back-to-back time has on top of that is the launch-to-launch gap."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rap_b200 import _lib
SECS = float(os.environ.get("SECS", "1.0"))
def run(name, epi, bn, M, N, K):
    A = torch.randn(M, K, device="cuda").half(); W = (torch.randn(N, K, device="cuda") / 22).half()
    bias = torch.randn(N, device="cuda"); add = torch.randn(M, N, device="cuda") if epi in (0, 5) else None
    gamma = torch.ones(2, 8, 64, device="cuda")
    out32 = (add if epi in (0, 5) else None)          # in place, as in the sampler
    out16 = torch.zeros(M, N, device="cuda", dtype=torch.float16) if (epi & 0xff) not in (0, 5) else None
    dbg = torch.zeros(128 + 4 * 160, dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    def launch():
        _lib.check(_lib.lib.rap_test_gemm(epi, bn, M, N, K, A.data_ptr(), W.data_ptr(), bias.data_ptr(), _lib.ptr(add), _lib.ptr(out32), _lib.ptr(out16),
                                          gamma.data_ptr(), dbg.data_ptr(), _lib.stream_ptr()))
    t_end = time.time() + SECS
    while time.time() < t_end:
        for _ in range(200): launch()
        torch.cuda.synchronize()
    ev[0].record()
    for _ in range(200): launch()
    ev[1].record(); torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1e3 / 200
    d = dbg.cpu().numpy(); tl = d[:128].reshape(16, 8); t0 = tl[0][0]
    ctas = d[128:].reshape(-1, 4); ctas = ctas[ctas[:, 0] != 0]
    ghz = np.median((ctas[:, 3] - ctas[:, 1]) / np.maximum(ctas[:, 2] - ctas[:, 0], 1))
    e = (ctas[:, 0] - ctas[:, 0].min()) / 1e3; x = (ctas[:, 2] - ctas[:, 0].min()) / 1e3
    print(f"== {name}: M={M} N={N} K={K} BN={bn}: {us:.2f} us per launch back to back ({2*M*N*K/us/1e6:.0f} TFLOP/s), SM clock {ghz:.2f} GHz, {len(ctas)} CTAs")
    print(f"  CTA entry us min/median/max {e.min():.2f}/{np.median(e):.2f}/{e.max():.2f}; exit {x.min():.2f}/{np.median(x):.2f}/{x.max():.2f}; "
          f"lifetime median {np.median(x-e):.2f} max {np.max(x-e):.2f}; kernel span {x.max():.2f} -> launch-to-launch gap {us - x.max():.2f} us")
    for t in range(2 if (epi & 0xff) in (0, 5) else 4):
        r = tl[t]
        if r[0] == 0: break
        print(f"  CTA 0 tile{t} (cycles from CTA entry {r[0]-ctas[0][1]}): prod [{r[0]-t0},{r[1]-t0}]  mma wait_acc {r[3]-r[2]} issue_span {r[4]-r[3]} (start {r[2]-t0})  epi start {r[5]-t0} acc_ready {r[6]-t0} done {r[7]-t0} (epi work {r[7]-r[6]})")
    if epi == 5: print(f"  all bulk reductions of warp 2 complete at {tl[15][0]-t0} (slot 'done' above = last one issued)")
    print(f"  CTA 0 lifetime {ctas[0][3]-ctas[0][1]} cycles")
run("outproj f32 in-place", 0, 256, 8192, 512, 512)
run("outproj bulk-reduction (EPI_F32_RED)", 5, 256, 8192, 512, 512)
run("ff2 f32 in-place", 0, 256, 8192, 512, 2048)
run("ff2 bulk-reduction (EPI_F32_RED)", 5, 256, 8192, 512, 2048)
if not os.environ.get("F32_ONLY"):
    run("outproj f32 bn128", 0, 128, 8192, 512, 512)
    run("ff2 f32 bn128", 0, 128, 8192, 512, 2048)
    run("qkv rmsnorm", 2, 256, 8192, 1536, 512)
    run("geglu", 3, 256, 8192, 4096, 512)
    run("geglu, scalar gelu epilogue", 3 | 0x400, 256, 8192, 4096, 512)
    run("qkv rmsnorm, bulk tensor stores", 2 | 0x800, 256, 8192, 1536, 512)
    run("geglu, bulk tensor stores", 3 | 0x800, 256, 8192, 4096, 512)
    run("qkv rmsnorm, bulk tensor stores, column-major tiles", 2 | 0x1800, 256, 8192, 1536, 512)
