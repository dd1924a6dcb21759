"""Turns gpurun_out/{launches.csv, prof_fmha.ncu-rep, prof_gemm.ncu-rep} into profiles/<tag>_ncu_summary.md (+ the raw launch list)."""
import collections, csv, os, re, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
note = sys.argv[2] if len(sys.argv) > 2 else ""
G = "gpurun_out"
def launch_summary(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = row["Kernel Name"]; m = re.search(r"gemm_kernel<(\d+), *(\d+)(?:, *(\w+))?(?:, *\w+)*>", name); f = re.search(r"fmha_kernel<(\d+)>", name)
        key = (f"gemm_kernel<{m.group(1)},{m.group(2)}" + (",mc" if m.group(3) in ("1", "true") else "") + ">") if m else f"fmha_kernel<{f.group(1)}>" if f else re.sub(r"\(.*", "", name).split("::")[-1]
        agg[key][0] += 1; agg[key][1] += v
    tot = sum(t for _, t in agg.values())
    return [(k, c, t, t / c, t / tot) for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])], tot
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size"]
out = [f"# {tag} -- ncu evidence, cfg2 (2x4096 pts, 20 flow steps, rap_12), 1x B200", "", note, "",
       "Commands: `bash tools/gpu_profile.sh` = `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` over ONE registered pair",
       "(tools/profile_pair.py brackets it with cudaProfilerStart/Stop) and `ncu --set full --clock-control none --import-source on` on fmha_kernel / gemm_kernel.",
       "Times under ncu are cold-cache and serialised: compare SHARES with bench.py's kernel_breakdown_ms, not absolutes.", "",
       "## Per-launch durations of one registered pair", "", "| kernel | launches | total ms | avg us | share |", "|---|---|---|---|---|"]
rows, tot = launch_summary(f"{G}/launches.csv")
for k, c, t, a, sh in rows: out.append(f"| {k} | {c} | {t/1e3:.2f} | {a:.1f} | {sh:.1%} |")
out += ["", f"total {tot/1e3:.1f} ms over {sum(r[1] for r in rows)} launches", ""]
for rep, title in [("prof_fmha", "fmha_kernel<2>, the shipped hardware-scheduled kernel (launch 1: per-view segments 2x4096, launch 2: per-scene 1x8192)"), ("prof_gemm", "gemm_kernel<BN,EPI,MC> (EPI 0=f32 bias+residual, 1=f16 silu, 2=qkv+rmsnorm, 3=geglu, 5=in-place residual through bulk reductions; mc = 2-CTA clusters sharing the weight tile by multicast)")]:
    raw = subprocess.run(["ncu", "-i", f"{G}/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines())); hdr, units = r[0], r[1]; idx = {h: i for i, h in enumerate(hdr)}
    out += [f"## ncu --set full: {title}", "", "| kernel | " + " | ".join(w.replace(".avg.pct_of_peak_sustained_active", " %act").replace(".avg.pct_of_peak_sustained_elapsed", " %el") for w in want) + " |", "|" + "---|" * (len(want) + 1)]
    for d in r[2:]:
        name = d[idx["Kernel Name"]]; m = re.search(r"gemm_kernel<(\d+), *(\d+)(?:, *(\w+))?(?:, *\w+)*>", name)
        nm = (f"gemm<{m.group(1)},{m.group(2)}" + (",mc" if m.group(3) in ("1", "true") else "") + ">") if m else re.sub(r"\(.*", "", name).split("::")[-1]
        out.append(f"| {nm} | " + " | ".join(f"{d[idx[w]]} {units[idx[w]]}" for w in want) + " |")
    out.append("")
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
shutil.copy(f"{G}/launches.csv", f"profiles/{tag}_ncu_launches_cfg2.csv")
print("\n".join(out[:40]))
