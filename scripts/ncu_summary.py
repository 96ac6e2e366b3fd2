"""ncu report -> compact per-launch CSV of the metrics the roofline / stall discussion uses.
usage: ncu_summary.py <report.ncu-rep> <out.csv>"""
import csv, subprocess, sys
rep, out = sys.argv[1:3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
h = rows[0]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_tma_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
idx = [(w, h.index(w)) for w in want if w in h]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([k for k, _ in idx] + ["units: see the ncu raw page (second header row)"])
    w.writerow([rows[1][i] for _, i in idx])
    for r in rows[2:]:
        if len(r) > 10:
            w.writerow([r[i] for _, i in idx])
print(open(out).read()[:1500])
