"""gpurun_out/prof_hbm.ncu-rep -> profiles/<tag>_ncu_hbm_kernels.md: achieved DRAM / L2 traffic and GB/s of the HBM-bound kernels."""
import csv, json, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6574.8
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_hbm.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines())); hdr, units = r[0], r[1]; idx = {h: i for i, h in enumerate(hdr)}
def val(d, k):
    try: return float(d[idx[k]].replace(",", ""))
    except (KeyError, ValueError): return float("nan")
def to_bytes(v, u): return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_us(v, u): return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
out = [f"# {tag} -- ncu --set full, HBM-/latency-bound kernels (tools/gpu_profile_hbm.sh), 1x B200", "",
       f"GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; peak = measured HBM copy bandwidth {peak} GB/s (MEASURED_PEAKS.json).",
       "L2 GB/s = lts__t_sectors.sum x 32 B / duration (data that is L2-resident between kernels never reaches DRAM: see the lts column before judging the dram one).",
       "Under ncu every launch is cold-cache and serialised.", "",
       "| kernel | grid | us | dram read MB | dram write MB | dram GB/s | frac of HBM peak | L2 traffic MB | L2 GB/s | lts %peak | sm busy % | regs |", "|" + "---|" * 12]
for d in r[2:]:
    name = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).split("::")[-1]
    us = to_us(val(d, "gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
    rd = to_bytes(val(d, "dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]]); wr = to_bytes(val(d, "dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
    l2 = val(d, "lts__t_sectors.sum") * 32.0
    gbs = (rd + wr) / us / 1e3
    out.append(f"| {name} | {d[idx['launch__grid_size']]} | {us:.1f} | {rd/1e6:.2f} | {wr/1e6:.2f} | {gbs:.0f} | {gbs/peak:.3f} | {l2/1e6:.1f} | {l2/us/1e3:.0f} | "
               f"{val(d, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {val(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {d[idx['launch__registers_per_thread']]} |")
open(f"profiles/{tag}_ncu_hbm_kernels.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
