"""Summaries of an `ncu --set full` report for profiles/: selected raw metrics as CSV, the executed
SASS opcode mix per tile, and the DRAM traffic per voice-sample (traffic.json).
usage: python tools/ncu_summary.py report.ncu-rep <kernel> <voices> <samples per voice in the launch>
       <samples per tile> <out prefix> "<command line that was profiled>" """
import csv, json, subprocess, sys
from collections import Counter

rep, kernel, voices, samples, tile, prefix, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6], sys.argv[7]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units, vals = rows[0], rows[1], rows[2]
d = {n: (u, v) for n, u, v in zip(names, units, vals)}
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct"]
keep += sorted(k for k in d if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k][1] or 0) >= 0.04)
vs = voices * samples
with open(prefix + ".csv", "w") as f:
    f.write(f"# ncu --set full --clock-control none, kernel {kernel}\n# `{cmd}`\n")
    f.write(f"# {voices} voices x {samples} samples = {vs:.4g} voice-samples in the captured launch\n")
    f.write("metric,unit,value\n")
    for k in keep:
        if k in d:
            f.write(f"{k},{d[k][0]},{d[k][1]}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
tiles = voices / 32 * (samples / tile)
c, smp, tot = Counter(), Counter(), 0
for r in rows[2:]:
    n = int(r[ix])
    t = r[1].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    parts = op.split(".")
    key = parts[0] + ("." + parts[1] if len(parts) > 1 and parts[0] in ("F2F", "I2F", "F2I", "MUFU", "LDS", "STS", "STG") else "")
    c[key] += n
    smp[key] += int(r[isamp])
    tot += n
with open(prefix + "_sass_mix.txt", "w") as f:
    f.write(f"executed SASS of {kernel}, per warp and tile of {tile} samples (32 voices x {tile} samples)\n")
    f.write(f"total {tot / tiles:.1f} warp-instructions per tile = {tot / tiles / tile:.2f} thread-instructions per voice-sample\n")
    f.write("opcode         per tile   share   stall-sample share\n")
    ts = sum(smp.values())
    for op, n in c.most_common(40):
        if n / tiles >= 0.5:
            f.write(f"{op:14s} {n / tiles:8.1f}  {100 * n / tot:5.1f}%  {100 * smp[op] / ts:5.1f}%\n")
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
rd = float(d["dram__bytes_read.sum"][1]) * unit[d["dram__bytes_read.sum"][0]]
wr = float(d["dram__bytes_write.sum"][1]) * unit[d["dram__bytes_write.sum"][0]]
print(json.dumps({"kernel": kernel, "dram_bytes_per_voice_sample": (rd + wr) / vs, "algorithmic_bytes_per_voice_sample": 4.0,
                  "source": f"{prefix.split('/')[-1]}.csv: (dram__bytes_read.sum + dram__bytes_write.sum) / ({voices} voices x {samples} samples) "
                            "of one ncu --set full launch of the same kernel and per-voice work, scaled to this launch's voice-samples"}, indent=1))
