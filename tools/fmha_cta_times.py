"""Per-CTA durations of ONE persistent attention launch (globaltimer + clock64 recorded by every CTA): is the launch time the common per-block
cost, or the slowest SM of a statically balanced partition?  RAP_FMHA_CTAS=n runs the same work on n CTAs (contention probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from rap_b200 import _lib
T = 8192
qkv = torch.randn(T, 3, 8, 64, device="cuda"); qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8; qkv = qkv.half().contiguous()
out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
for cu in ([0, T], [0, T // 2, T]):
    dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
    for _ in range(3):
        _lib.check(_lib.lib.rap_test_fmha(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, 0, out.data_ptr(), dbg.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    d = dbg.cpu().numpy().reshape(-1, 4)
    d = d[d[:, 1] > 0]
    dur = (d[:, 1] - d[:, 0]) / 1e3
    span = (d[:, 1].max() - d[:, 0].min()) / 1e3
    mhz = d[:, 2] / dur
    blocks = sum(((e - s + 255) // 256) * 8 * ((e - s + 127) // 128) for s, e in zip(cu[:-1], cu[1:]))
    print(f"cu={cu}: {len(d)} CTAs, launch span {span:.1f} us; CTA duration us min/median/max {dur.min():.1f}/{np.median(dur):.1f}/{dur.max():.1f}; "
          f"start skew {(d[:, 0].max() - d[:, 0].min()) / 1e3:.1f} us; SM clock from clock64/globaltimer median {np.median(mhz):.0f} MHz; "
          f"cycles per 256x128 block (median CTA) {np.median(d[:, 2]) / (blocks / len(d)):.0f}")
    order = np.argsort(dur)
    print("   slowest 5 (cta, sm, us):", [(int(i), int(d[i, 3]), round(float(dur[i]), 1)) for i in order[-5:]], " fastest 5:", [(int(i), int(d[i, 3]), round(float(dur[i]), 1)) for i in order[:5]])
