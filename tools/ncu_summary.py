"""Developer tool: condense an `ncu --set full --import-source on` report of the PDIPM kernel into the
small CSV summaries kept under profiles/ (raw metrics + per-source-line instruction / stall /
shared-memory wavefront counts).  Usage: python tools/ncu_summary.py REPORT.ncu-rep OUT_PREFIX [envs iters]"""
import collections, csv, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
N = (int(sys.argv[3]) if len(sys.argv) > 3 else 4096) * (int(sys.argv[4]) if len(sys.argv) > 4 else 5)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio")
with open(out + "_summary.csv", "w") as f:
    f.write(f"# {' '.join(sys.argv)}\nmetric,value,unit\n")
    for i, h in enumerate(hdr):
        if h in keep or ("issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h):
            f.write(f"{h},{vals[i]},{units[i]}\n")
    try:
        i_w = hdr.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"); i_c = hdr.index("sm__cycles_elapsed.avg")
        f.write(f"derived_smem_wavefronts_per_sm_cycle,{float(vals[i_w]) / 148 / float(vals[i_c]):.4f},of 1.0\n")
    except ValueError:
        pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, h2, per = None, None, collections.OrderedDict()
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        h2 = r; ix = h2.index("Instructions Executed"); isamp = h2.index("# Samples"); iw = h2.index("L1 Wavefronts Shared"); continue
    if h2 is None or len(r) <= ix or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    per[(fname, ln)] = (int(r[ix] or 0), int(r[isamp] or 0), int(r[iw] or 0), r[1].strip())
ts = sum(v[1] for v in per.values()) or 1
with open(out + "_lines.csv", "w") as f:
    f.write("file,line,warp_instructions_per_env_iteration,stall_sample_share_pct,smem_wavefronts_per_env_iteration,source\n")
    for (fn, ln), v in sorted(per.items(), key=lambda kv: -kv[1][0]):
        if v[0] * 200 < N:
            continue
        f.write(f"{fn},{ln},{v[0] / N:.1f},{100 * v[1] / ts:.2f},{v[2] / N:.1f},\"{v[3][:100].replace(chr(34), chr(39))}\"\n")
print("wrote", out + "_summary.csv", out + "_lines.csv")
