"""Print selected metrics of an .ncu-rep (raw page) per kernel launch.
usage: ncu -i X.ncu-rep --page raw --csv | python ncu_raw.py"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    for w in WANT:
        if w in hdr:
            print(f"   {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
