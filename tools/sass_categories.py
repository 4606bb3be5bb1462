"""Executed warp instructions / active lanes / PC samples of one k_trace instantiation, grouped by what the code does.
Joins an ncu source page with nvdisasm line info like tools/sass_by_line.py; the groups are line ranges found through marker
strings in the sources the measured library was built from (so the table survives edits that move lines).
  ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
  cuobjdump -xelf all libdsrt.so; nvdisasm -g -c dsrt_api.sm_100a.cubin > dis.txt
  python tools/sass_categories.py src.csv dis.txt k_traceILb1ELb0 2 [csrc directory of the measured build]"""
import csv, os, re, sys
from collections import defaultdict

src, dis, sect, kidx = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
csrc = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dsgpuraytracing_b200", "csrc")


def marks(fname, table):
    """[(first line, group)] sorted by line: a group runs from its marker to the next marker"""
    text = open(os.path.join(csrc, fname)).read().split("\n")
    out = []
    for needle, group in table:
        hits = [i + 1 for i, l in enumerate(text) if needle in l]
        if not hits:
            raise SystemExit(f"marker not found in {fname}: {needle}")
        out.append((hits[0], group))
    return sorted(out)


RANGES = {
    "kernels.cuh": marks("kernels.cuh", [
        ("__device__ __forceinline__ void prefetch_l2", "shared-memory accesses through the helper lines (ray blocks, pair table, stack)"),
        ("template <bool ANY, bool COUNT>", "kernel prologue"),
        ("---- write the results of the rays", "result write-back"),
        ("---- refill idle lanes", "refill + per-ray set-up stores"),
        ("---- traverse until the warp is due", "node-step control (stack pop / push, node fetch call)"),
        ("(2) primitive step, warp-wide decision", "primitive-round decision (votes, trigger)"),
        ("Cooperative test: the pending", "cooperative test: reservation + scatter by the owners"),
        ("for (int base = 0; base < P; base += 32)", "cooperative test: pair decode, ray-block / record loads, flags"),
        ("if (!coop) {", "per-lane primitive loop (closest hit; small rounds)"),
        ("(3) retire finished rays", "retire checks, refill decision, loop control"),
    ]),
    "traverse.cuh": marks("traverse.cuh", [
        ("DSRT_HD int source_code", "source drop / record index (drop_source, prim_slot)"),
        ("DSRT_HD WatertightRay make_watertight", "per-ray set-up (make_frame, make_watertight, any_hit_scale)"),
        ("DSRT_HD bool hit_triangle(", "primitive test arithmetic"),
        ("DSRT_HD NodeRegs load_node", "node fetch (load_node, load_node_prims)"),
        ("DSRT_HD NodeFrame make_frame", "per-ray set-up (make_frame, make_watertight, any_hit_scale)"),
        ("DSRT_HD float byte_unit", "node test (test_children: dequantise, slabs, hit nibbles)"),
        ("DSRT_HD uint32_t order_children", "child ordering / selection (order_children, next_child, split_hits)"),
        ("DSRT_HD float any_hit_scale", "per-ray set-up (make_frame, make_watertight, any_hit_scale)"),
        ("DSRT_HD uint32_t next_child", "child ordering / selection (order_children, next_child, split_hits)"),
        ("DSRT_HD void trace_ray", "other"),
    ]),
}
HD = {"hd_fma_sat": "node test (test_children: dequantise, slabs, hit nibbles)", "hd_fma(": "primitive test arithmetic", "hd_mul": "primitive test arithmetic",
      "hd_sub": "primitive test arithmetic", "hd_rcp": "per-ray set-up (make_frame, make_watertight, any_hit_scale)", "hd_rsqrt": "per-ray set-up (make_frame, make_watertight, any_hit_scale)"}
hd_lines = {}
cur = None
for i, l in enumerate(open(os.path.join(csrc, "hd.h")).read().split("\n")):
    m = re.match(r"DSRT_HD \w+ (\w+\(?)", l)
    if m:
        cur = next((g for k, g in HD.items() if l.split("DSRT_HD")[1].strip().split(" ", 1)[1].startswith(k.rstrip("("))), "helpers (bit casts, clz / popc)")
        if "hd_fma(" in l: cur = HD["hd_fma("]
        if "hd_fma_sat" in l: cur = HD["hd_fma_sat"]
    hd_lines[i + 1] = cur or "helpers (bit casts, clz / popc)"


def group(f, ln):
    if f in RANGES:
        g = "other"
        for first, name in RANGES[f]:
            if ln >= first:
                g = name
        return g
    if f == "hd.h":
        return hd_lines.get(ln, "helpers (bit casts, clz / popc)")
    return "warp votes, shuffles, popc / clz intrinsics, atomics"


rows = list(csv.reader(open(src)))
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
k = ks[kidx]; h = k["hdr"]
iI, iT, iM = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
lines = []; inside = False; curline = ("?", 0)
for l in open(dis):
    if l.startswith("\t.section"):
        inside = (".text." in l) and (sect in l); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): lines.append(curline)
agg = defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for i in range(min(len(lines), len(k["rows"]))):
    r = k["rows"][i]; a = agg[group(*lines[i])]
    for j, c in enumerate((iI, iT, iM)):
        a[j] += int(r[c]); tot[j] += int(r[c])
print("kernel %s: %d SASS instructions (ncu) / %d (nvdisasm)" % (k["name"][:48], len(k["rows"]), len(lines)))
print("total: %d warp instructions, %.1f lanes per instruction, %d PC samples\n" % (tot[0], tot[1] / max(tot[0], 1), tot[2]))
print("| part | % of warp instructions | lanes | % of samples |\n|---|---|---|---|")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| %s | %.1f | %.1f | %.1f |" % (key, 100 * a[0] / tot[0], a[1] / max(a[0], 1), 100 * a[2] / max(tot[2], 1)))
