#!/usr/bin/env python
"""Tables for profiles/r1_ktrace_depth0_summary.md from two `ncu --page raw --csv` exports (production / no compaction) and
the dominant-kernel traffic JSON bench.py reads:
  python tools/profile_tables.py on.csv off.csv bench.json > table.md      (also rewrites profiles/dominant_kernel_traffic.json)"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(p):
    rows = list(csv.reader(open(p))); h = rows[0]
    units = {h[i]: rows[1][i] for i in range(len(h))}
    return [{h[i]: r[i] for i in range(len(h))} for r in rows[2:]], units


def g(r, k):
    return float(r[k].replace(",", ""))


F, units = load(sys.argv[1]); N, _ = load(sys.argv[2])
bench = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])


def to_bytes(r, k):
    return g(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[k]]


conn = F[1]
rd, wr = to_bytes(conn, "dram__bytes_read.sum"), to_bytes(conn, "dram__bytes_write.sum")
n_rays = 1920 * 1080 * 4 * 4            # 4 spp batch x 4 light samples (every camera path of the bench scene hits something)
json.dump({"traffic_bytes_per_launch": int(rd + wr),
           "launch": "k_trace<true> (connect) at depth 0 of one 4-spp wavefront batch at 1080p (8.3 M camera paths): 33.2 M shadow rays, %.3f ms under ncu" % g(conn, "gpu__time_duration.sum"),
           "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
           "algorithmic_bytes_in_launch": int(n_rays * bench["roofline"]["bytes_per_segment"]),
           "source": "" + os.path.relpath(sys.argv[1], ROOT) + " (ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 python profiles/profile_run.py 4)",
           "note": "DRAM traffic = streaming the 48-byte shadow-queue records once (33.2 M x 48 B = 1.59 GB) + framebuffer atomics; node / primitive fetches are served by L1 and L2"},
          open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"), "w"), indent=1)
K = [("duration (ms)", "gpu__time_duration.sum", "%.3f"), ("warp instructions (M)", "smsp__inst_executed.sum", None),
     ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", "%.1f"),
     ("active lanes / instruction", "smsp__thread_inst_executed_per_inst_executed.ratio", "%.1f"),
     ("branch targets uniform %", "smsp__sass_average_branch_targets_threads_uniform.pct", "%.1f"),
     ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "%.1f"),
     ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "%.1f"),
     ("L1 hit %", "l1tex__t_sector_hit_rate.pct", "%.1f"), ("L2 hit %", "lts__t_sector_hit_rate.pct", "%.1f"),
     ("L1 throughput % of peak", "l1tex__throughput.avg.pct_of_peak_sustained_active", "%.1f"),
     ("L2 throughput % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "%.1f"),
     ("registers / thread", "launch__registers_per_thread", "%.0f"),
     ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active", "%.1f")]
print("| metric | extend, compaction ON | extend, OFF | connect, compaction ON | connect, OFF |")
print("|---|---|---|---|---|")
for name, k, f in K:
    cell = lambda r: ("%.0f" % (g(r, k) / 1e6)) if f is None else f % g(r, k)
    print("| %s | %s | %s | %s | %s |" % (name, cell(F[0]), cell(N[0]), cell(F[1]), cell(N[1])))
print("| DRAM read + write (MB) | " + " | ".join("%.0f + %.0f" % (to_bytes(r, "dram__bytes_read.sum") / 1e6, to_bytes(r, "dram__bytes_write.sum") / 1e6) for r in (F[0], N[0], F[1], N[1])) + " |")
