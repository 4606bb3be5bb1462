#!/usr/bin/env python
"""Markdown table from the JSON lines of tools/run_configs.py:  python tools/configs_table.py a.jsonl [b.jsonl ...]"""
import json, sys
rows = [json.loads(l) for f in sys.argv[1:] for l in open(f) if l.strip().startswith("{")]
print("| case | GPUs | prims | wide nodes | accel MB | segments | s/frame (device) | wall s incl. reduce + read-back | Mrays/s | nodes/seg | prims/seg | B/seg | algorithmic GB/s | SAH build s | flatten+upload s |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    wall = r.get("wall_s_incl_reduce_and_readback")
    print("| %s | %d | %d | %d | %.1f | %.1f M | %.4f | %s | %.0f | %.2f | %.2f | %.0f | %.0f | %.3f | %.3f |" % (
        r["case"], r.get("n_gpus", 1), r["prims"], r["wide_nodes"], r["accel_MB"], r["segments"] / 1e6, r["s_per_frame"],
        ("%.4f" % wall) if wall is not None else "-", r["Mrays_s"], r["nodes_per_seg"], r["prims_per_seg"], r["bytes_per_seg"],
        r["algorithmic_GB_s"], r["sah_build_s"], r["flatten_upload_s"]))
