"""Key counters per launch from an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep [out.csv]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = [hdr.index("ID"), hdr.index("Kernel Name")] + [hdr.index(k) for k in KEYS if k in hdr]
out = io.StringIO()
w = csv.writer(out)
w.writerow([hdr[c] for c in cols]); w.writerow([units[c] for c in cols])
for r in data:
    row = [r[c] for c in cols]
    row[1] = row[1].split("(")[0].replace("<unnamed>::", "")
    w.writerow(row)
txt = out.getvalue()
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
