"""Dumps the clock64() timeline of one FMHA CTA (debug aid): per K/V block, cycles spent in each softmax / UMMA-issue phase."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from rap_b200 import _lib
T = int(os.environ.get("T", 8192)); nkv = T // 128
qkv = torch.randn(T, 3, 8, 64, device="cuda"); qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8; qkv = qkv.half().contiguous()
out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
dbg = torch.zeros(nkv, 16, dtype=torch.int64, device="cuda")
cu = _lib.i32_array([0, T])
for it in range(3):
    _lib.check(_lib.lib.rap_test_fmha(qkv.data_ptr(), T, 8, cu, 1, int(os.environ.get("FORCE", 3)), out.data_ptr(), dbg.data_ptr(), _lib.stream_ptr()))   # 3 = the one-item-per-CTA kernel (its timeline format)
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.int64)
names = ["s_wait", "ld_S", "max", "exp", "pv_wait", "sts_P", "fence+arrive"]
print("softmax warp (tile 0), cycles per phase, blocks 8..15:")
for j in range(8, 16):
    row = d[j]; ph = [row[i + 1] - row[i] for i in range(7)]
    print(j, dict(zip(names, ph)), "iter period", d[j + 1][0] - d[j][0])
print("UMMA warp: p_full(t0) ready -> after QK issue -> after PV issue | same for t1; relative to softmax t0 arrive")
for j in range(8, 12):
    row = d[j]
    print(j, "t0:", row[8] - row[7], row[9] - row[8], row[10] - row[9], " t1:", row[12] - row[10], row[13] - row[12], row[14] - row[13], " loop period", d[j + 1][8] - row[8])
per = np.diff(d[4:60, 0]); print("mean softmax period", per.mean(), "cycles; ideal MUFU-bound 1024/tile")

t_in, t_out = d[0][15], d[1][15]
starts = d[:nkv, 0]
print(f"CTA 0: entry -> first S wait {starts[0]-t_in} cycles; first 6 block periods {np.diff(starts[:7]).tolist()}; last 6 {np.diff(starts[-7:]).tolist()}")
print(f"       last block start -> softmax done {t_out-starts[-1]} cycles; whole CTA {t_out-t_in} cycles = {nkv} blocks x {(t_out-t_in)/nkv:.0f}")
