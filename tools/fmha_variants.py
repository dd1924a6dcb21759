"""Isolated per-launch time of the attention kernel variants on the headline shapes (rap_bench_fmha_variant: 0 default, 1 / 2 persistent, 3 three-tile)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from rap_b200 import _lib
T = 8192
qkv = torch.randn(T, 3, 8, 64, device="cuda"); qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8; qkv = qkv.half().contiguous()
out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
for v in [int(x) for x in (sys.argv[1:] or ["0", "3", "1", "0", "3"])]:
    r = []
    for cu in ([0, 4096, 8192], [0, 8192]):
        ms = C.c_float()
        _lib.check(_lib.lib.rap_bench_fmha_variant(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, out.data_ptr(), 200, v, C.byref(ms)))
        r.append(ms.value * 1e3)
    print("variant", v, "per-view %.1f us  per-scene %.1f us" % tuple(r), flush=True)
