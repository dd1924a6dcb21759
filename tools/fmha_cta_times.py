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
    dbg = torch.zeros(8192 + 512, dtype=torch.int64, device="cuda")
    for _ in range(3):
        _lib.check(_lib.lib.rap_test_fmha(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, 0, out.data_ptr(), dbg.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    raw = dbg.cpu().numpy()
    tl = raw[8192:].reshape(-1, 8)
    d = raw[:8192].reshape(-1, 4)
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
    print("   CTA 5 timeline (cycles): item: blocks, split piece(+1), first block, per-block steady, last block done -> P.V retired, -> partial/combine, -> output, gap to next item")
    tl = tl[tl[:, 0] > 0]
    for i, r in enumerate(tl):
        nb = int(r[6]); nxt = int(tl[i + 1][0] - r[5]) if i + 1 < len(tl) else -1
        print(f"     item {i}: {nb} blocks, piece {int(r[7])}: first {int(r[1]-r[0])}, steady {(r[2]-r[1])/max(nb-1,1):.0f}/block, pv {int(r[3]-r[2])}, partial {int(r[4]-r[3])}, out {int(r[5]-r[4])}, next-start gap {nxt}; item total {int(r[5]-r[0])}")
