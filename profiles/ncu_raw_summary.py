"""Key metrics of an `ncu --page raw --csv` dump.  Usage: python profiles/ncu_raw_summary.py prof_raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = """gpu__time_duration.sum launch__registers_per_thread launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
launch__shared_mem_per_block_dynamic launch__grid_size launch__block_size sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio dram__bytes_read.sum dram__bytes_write.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
sass__inst_executed_local_loads sass__inst_executed_local_stores sass__inst_executed_shared_loads sass__inst_executed_shared_stores
smsp__warps_eligible.avg.per_cycle_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct
sm__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed""".split()
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k][0]:>18s} {d[k][1]}")
print("stall reasons (warp-cycles per issued instruction):")
for h in sorted(d):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
        try:
            x = float(d[h][0])
        except ValueError:
            continue
        if x > 0.05:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {x:6.2f}")
