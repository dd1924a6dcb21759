"""Is the path clock-/power-limited?  Runs (a) the per-scene attention launch, (b) the QKV GEMM and (c) whole pair registrations back to back for
a few seconds each while nvidia-smi samples SM clock and board power every 50 ms, and prints per-launch time, median clock and power.
A B200 under its 1000 W cap cannot hold 1965 MHz with the tensor + MUFU pipes busy: the per-SM per-clock bounds quoted in DESIGN.md (MUFU:
16 ex2/clk/SM) turn into time through the SUSTAINED clock measured here, not the nominal one."""
import ctypes as C, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from rap_b200 import _lib
from rap_b200.engine import Registrar
from rap_b200.synth import DiTConfig, make_batch, make_state_dict


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
                                  stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for ln in self.p.stdout:
            try:
                self.rows.append((time.perf_counter(), *[float(x) for x in ln.split(",")]))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.3 <= x[0] <= t1]
        if not r:
            return "no samples"
        clk = sorted(x[1] for x in r); pw = sorted(x[2] for x in r)
        return f"sm clock median {clk[len(clk)//2]:.0f} MHz (min {clk[0]:.0f}, max {clk[-1]:.0f}), power median {pw[len(pw)//2]:.0f} W (max {pw[-1]:.0f}), {len(r)} samples"


smi = Smi()
T = 8192
qkv = torch.randn(T, 3, 8, 64, device="cuda"); qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8; qkv = qkv.half().contiguous()
out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
time.sleep(1.0)
t0 = time.perf_counter(); time.sleep(1.0); print("idle:", smi.window(t0 - 0.3, time.perf_counter()))
for name, cu in (("attention per-scene [0,8192]", [0, T]), ("attention per-view [0,4096,8192]", [0, 4096, T])):
    for iters in (200, 15000):
        ms = C.c_float()
        t0 = time.perf_counter()
        _lib.check(_lib.lib.rap_bench_fmha(qkv.data_ptr(), T, 8, _lib.i32_array(cu), len(cu) - 1, out.data_ptr(), iters, C.byref(ms)))
        t1 = time.perf_counter()
        print(f"{name}: {iters} launches, {ms.value*1e3:.1f} us/launch;", smi.window(t0, t1) if iters > 1000 else "(short burst)")
    time.sleep(1.0)
a = torch.randn(T, 512, device="cuda", dtype=torch.float16); w = torch.randn(1536, 512, device="cuda", dtype=torch.float16) / 22.6
o16 = torch.empty(T, 1536, device="cuda", dtype=torch.float16); gamma = torch.ones(1024, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for iters in (200, 150000):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(iters):
        _lib.lib.rap_test_gemm(2, 256, T, 1536, 512, a.data_ptr(), w.data_ptr(), None, None, None, o16.data_ptr(), gamma.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"QKV GEMM: {iters} launches, {e0.elapsed_time(e1)/iters*1e3:.1f} us/launch;", smi.window(t0, t1) if iters > 1000 else "(short burst)")
time.sleep(1.0)
big = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(4000):
    torch.matmul(big, big)
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"cuBLAS bf16 8192^3: {2*8192**3/(e0.elapsed_time(e1)/4000*1e-3)/1e12:.0f} TFLOP/s;", smi.window(t0, t1))
del big
time.sleep(1.0)
reg = Registrar(make_state_dict(DiTConfig(num_layers=12)), num_layers=12, device="cuda", num_steps=20)
d = make_batch(1, 2, 4096).to("cuda")
plan = reg.plan_for(d.cu_seqlens_batch.cpu(), d.points_per_part.cpu())
for n in (3, 40):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        reg.register_device(plan, d.cond, d.features, d.scales, d.anchor_indices, d.x_1)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"pair registration: {n} pairs, {e0.elapsed_time(e1)/n:.2f} ms/pair;", smi.window(t0, t1) if n > 10 else "(short)")
smi.p.terminate()
