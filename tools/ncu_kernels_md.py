"""One markdown table per distinct kernel of an .ncu-rep (first launch of each): time, pipes, memory system, occupancy,
top stall reasons.  usage: ncu_kernels_md.py rep title out.md"""
import csv, io, subprocess, sys
rep, title, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers / thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
        ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor-core unit (tcgen05, operand fetch included) cycles active"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor math pipe cycles active"),
        ("sm__ops_path_tensor_src_tf32_dst_fp32.avg.pct_of_peak_sustained_elapsed", "TF32 tensor throughput"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts read by the tensor cores"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (LTS) throughput"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"), ("lts__t_sectors.sum", "L2 sectors (32 B)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("smsp__inst_executed.sum", "warp instructions")]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
# one table per kernel name: its LONGEST launch (an adaptive run launches both repulsion kernels every iteration and the
# one whose form was not chosen returns at once)
best = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].replace("unnamed>::", "").replace("void ", "")
    short = name.split("(")[0]
    try:
        dur = float(d["gpu__time_duration.sum"].replace(",", ""))
    except (KeyError, ValueError):
        dur = 0.0
    if short not in best or dur > best[short][0]:
        best[short] = (dur, d)
text = ["# " + title, ""]
for short, (_dur, d) in sorted(best.items(), key=lambda kv: -kv[1][0]):
    text += ["## `%s`" % short, "", "| metric | value |", "|---|---|"]
    for k, label in KEYS:
        if k in d and d[k] != "":
            v = d[k]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            text.append("| %s (`%s`) | %s %s |" % (label, k, v, units[hdr.index(k)]))
    st = sorted(((float(d[h] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)
    text += ["| warps stalled per issued instruction, by reason | " + ", ".join("%s %.2f" % (n, v) for v, n in st[:6]) + " |", ""]
open(out, "w").write("\n".join(text) + "\n")
print("\n".join(text))
