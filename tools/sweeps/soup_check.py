"""Large-soup check: a 32 Mi-triangle soup rendered several times under different knobs (HBM regime)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sc, cam = S.triangle_soup(n << 20)
core = D.Core(0); core.set_params(64, 1, 8, 0); core.load(sc, camera=cam); core.set_option("stage_timing", 1)
for o in [{}, {}, {"max_ctas_per_sm": 6}, {"max_ctas_per_sm": 5}, {"max_ctas_per_sm": 4}, {"pool_batches": 16}, {"refill_busy_lanes": 26}, {"refill_busy_lanes": 12}]:
    d = {"max_ctas_per_sm": 0, "pool_batches": 8, "refill_busy_lanes": 20}; d.update(o)
    for k, v in d.items():
        core.set_option(k, v)
    rgb, st = core.render()
    print(n, o, "Mrays/s %.1f gpu_s %.4f extend %.4f connect %.4f shade %.4f" % (st.segments / st.gpu_seconds / 1e6, st.gpu_seconds, st.extend_seconds, st.connect_seconds, st.shade_seconds), flush=True)
