"""ncu report -> the two text summaries kept under profiles/ (<out>_metrics.txt: selected raw metrics per launch,
<out>_details.txt: the details page, one line per item).
usage: summarize_ncu.py <report.ncu-rep> <out prefix> [kernel-name regex]"""
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size launch__block_size
smsp__inst_executed.sum sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tensor.sum
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_uniform.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum smsp__cycles_active.avg
smsp__issue_active.avg.pct_of_peak_sustained_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
lts__t_sector_hit_rate.pct l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio""".split()

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out + "_metrics.txt", "w") as f:
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        f.write(f"Kernel Name = {name}\n")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"{w} = {r[i]} {units[i]}\n")
        f.write("\n")
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(det)))
h = rows[0]
ix = {k: h.index(k) for k in ("Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value")}
with open(out + "_details.txt", "w") as f:
    for r in rows[1:]:
        name = r[ix["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        short = name.split("(")[0].split("::")[-1]
        f.write(f"{short} | {r[ix['Section Name']]} | {r[ix['Metric Name']]} | {r[ix['Metric Value']]} {r[ix['Metric Unit']]}\n")
print("wrote", out + "_metrics.txt", out + "_details.txt")
