"""Key metrics per kernel from an `ncu -i X.ncu-rep --page raw --csv` dump.  usage: python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid CTAs"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"), ("launch__waves_per_multiprocessor", "waves")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].replace("void <unnamed>::", "").replace("<unnamed>::", "")
    print(f"== {name.split('(')[0]}   grid {r[hdr.index('Grid Size')]}  block {r[hdr.index('Block Size')]}")
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            print(f"   {label:32s} {r[i]:>14s} {units[i]}")
