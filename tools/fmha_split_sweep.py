"""A/B of the key-split last wave of the attention kernel: per-launch time at cfg2's two layouts for several small-piece shares.
Each setting runs in its own process (the schedule knobs are read once from the environment)."""
import os, subprocess, sys
CHILD = r'''
import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, torch.nn.functional as F
from rap_b200 import _lib
T = 8192
qkv = torch.randn(T, 3, 8, 64, device="cuda"); qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8; qkv = qkv.half().contiguous()
out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
res = []
for cu in ([0, 4096, 8192], [0, 8192]):
    ms = C.c_float()
    _lib.check(_lib.lib.rap_bench_fmha(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, out.data_ptr(), 200, C.byref(ms)))
    res.append(ms.value * 1e3)
print("%-28s part-level %.1f us   scene-level %.1f us" % (os.environ.get("TAG"), res[0], res[1]))
'''
settings = [("no split", {"RAP_FMHA_SPLIT": "0"})] + [(f"small piece {f:.3f}", {"RAP_FMHA_SMALL_FRAC": str(f)}) for f in (0.10, 0.125, 0.16, 0.19, 0.22, 0.25)]
for tag, env in settings:
    e = dict(os.environ, TAG=tag, **env)
    subprocess.run([sys.executable, "-c", CHILD], env=e, check=False)
