#!/usr/bin/env python
"""profiles/r2_configs.md from the bench.py lines of tools/measure_r2.sh (one per workload / builder) and the multi-GPU files of
tools/measure_r2_8gpu.sh:  python tools/configs_table_r2.py profiles > profiles/r2_configs.md"""
import json, os, sys
P = sys.argv[1]


def lines(name):
    p = os.path.join(P, name)
    if not os.path.exists(p):
        return []
    out = []
    for l in open(p):
        l = l.strip()
        if l.startswith("{"):
            out.append(json.loads(l))
    return out


print("# BASELINE.json configurations on B200 (round 2, final build)\n")
print("Every 1-GPU row is one `python bench.py --workload W [--device-build] --steps 2 --warmup 3` line (`r2_configs.jsonl`; c2 = the headline,\n"
      "`r2_bench_1gpu.json`): device-timed frames with inputs resident, 256 MiB L2 flush between steps; `e2e` = through the C ABI with host\n"
      "buffers (scene re-upload + render + frame read-back per step).  Roofline: algorithmic bytes of the dominant kernel / its CUDA-event\n"
      "time, against the measured L2 read bandwidth while wide BVH + primitive records fit in the 126 MB L2, against the HBM copy peak above.\n")
print("| workload | builder | triangles | Mrays/s | ms/frame | e2e Mrays/s | dominant kernel | nodes / seg | prims / seg | B / seg | roofline bound | achieved GB/s | frac | prepare s (SAH + collapse/flatten/upload) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
rows = lines("r2_bench_1gpu.json") + lines("r2_configs.jsonl")
for d in rows:
    c, r, e = d["config"], d["roofline"], d["e2e"]
    print(f"| {c['name']} ({c['width']}x{c['height']}, {c['spp']} spp, l{c['light_samples']} m{c['max_depth']}) | {'device LBVH' if 'device' in c.get('builder', '') else 'host SAH'} | {c['triangles']} | "
          f"{d['value']:.0f} | {d['ms_per_step']:.1f} | {e['value']:.0f} | {r['kernel'].split(' ')[0]} | {r['nodes_per_segment']:.1f} | {r['prims_per_segment']:.2f} | {r['bytes_per_segment']:.0f} | "
          f"{r['bound']} ({r['peak']:.0f} GB/s) | {r['achieved']:.0f} | {r['frac']:.3f} | {e['sah_build_seconds_once']:.2f} + {e['scene_prepare_seconds_once']:.2f} |")
print("\n## Multi-GPU (one process per GPU under torchrun + one NCCL reduce; `r2_bench_{2,4,8}gpu.json`)\n")
print("| workload | GPUs | split | Mrays/s | ms/frame | x 1 GPU (same box) | e2e Mrays/s | render ms per rank (min / max) | tail ms | reduce ms |")
print("|---|---|---|---|---|---|---|---|---|---|")
one = lines("r2_bench_1gpu_samebox.json")
base = one[0]["value"] if one else None
for name in ("r2_bench_1gpu_samebox.json", "r2_bench_2gpu.json", "r2_bench_4gpu.json", "r2_bench_8gpu.json", "r2_bench_8gpu_tiles.json", "r2_bench_8gpu_c4.json"):
    for d in lines(name):
        m = d["multi_gpu"]; c = d["config"]
        rel = f"{d['value'] / base:.2f}" if base and c["name"] == "c2" else "-"
        print(f"| {c['name']} | {d['n_gpus']} | {'tiles' if 'tile' in c['parallelism'] else 'samples'} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {rel} | {d['e2e']['value']:.0f} | "
              f"{min(m['render_ms_per_rank']):.2f} / {max(m['render_ms_per_rank']):.2f} | {m['tail_ms']:.2f} | {m['reduce_ms']:.3f} |")
print("\n## One context over 8 GPUs (`dsrt_create_multi`: sample split, scene host -> GPU 0 -> peers, fused peer reduce + resolve; `tools/run_configs.py --gpus 8`)\n")
print("| case | builder | Mrays/s | s/frame (device) | wall s incl. reduce + read-back | nodes / seg | host SAH s | build_accel s (collapse or device build, records, upload to 8 GPUs) |")
print("|---|---|---|---|---|---|---|---|")
for d in lines("r2_configs_8gpu.jsonl"):
    print(f"| {d['case']} | {d['builder']} | {d['Mrays_s']:.0f} | {d['s_per_frame']:.4f} | {d['wall_s_incl_reduce_and_readback']:.4f} | {d['nodes_per_seg']:.1f} | {d['sah_build_s']:.2f} | {d['flatten_upload_s']:.2f} |")
