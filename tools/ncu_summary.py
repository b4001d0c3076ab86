"""Summarise `ncu --page raw --csv` exports (one file per kernel) into the text kept under profiles/.
    python tools/ncu_summary.py out.txt file.csv [file.csv ...]"""
import csv, sys
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("launch__registers_per_thread", "registers / thread"), ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM allowed by smem"), ("launch__occupancy_limit_registers", "CTAs/SM allowed by registers"),
]
out = open(sys.argv[1], "w")
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    for d in rows[hi + 2:]:
        out.write("== %s   (%s)\n" % (d[hdr.index("Kernel Name")][:110], path.split("/")[-1]))
        for key, label in WANT:
            if key in hdr:
                out.write("   %-34s %s %s\n" % (label, d[hdr.index(key)], units[hdr.index(key)]))
        st = [(float(d[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for i, h in enumerate(hdr)
              if "issue_stalled" in h and "per_issue_active" in h and d[i] not in ("", "n/a")]
        out.write("   stalls per issued instruction:     " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:6]) + "\n\n")
