import sys, time; sys.path.insert(0,'.')
import numpy as np
import dsgpuraytracing_b200 as D
from tests.scenes import CONFIGS
for name, W, H, spp in [("CBspheres_lambertian",480,360,16), ("CBbunny",1920,1080,16), ("CBgems",1920,1080,16), ("bunny", 1920,1080,16)]:
    z = np.load(f'tests/golden/{name}.npz'); g = {k: z[k] for k in z.files}
    cfg = CONFIGS[name]
    cam = g["camera"].copy(); cam[14] *= H / cam[13]; cam[12], cam[13] = W, H
    core = D.Core(0)
    core.set_params(spp, cfg["nl"], cfg["depth"], 0)
    t=time.time(); core.load(g, camera=cam); print(name, 'load', time.time()-t, core.accel_info())
    core.set_option("stage_timing", 1)
    for it in range(3):
        rgb, st = core.render()
        print(name, it, 'gpu_s %.4f'%st.gpu_seconds, 'Mseg/s %.1f'%(st.segments/st.gpu_seconds/1e6), 'extend %d shadow %d'%(st.extend_rays, st.shadow_rays),
              'ext_s %.4f con_s %.4f shade_s %.4f'%(st.extend_seconds, st.connect_seconds, st.shade_seconds), 'launches', st.kernel_launches, 'mean', rgb.mean(axis=(0,1)))
    core.set_option("stage_timing", 0)
    rgb, st = core.render(); print(name, 'no stage timing: gpu_s %.4f Mseg/s %.1f'%(st.gpu_seconds, st.segments/st.gpu_seconds/1e6))
    core.set_option("count_traversal", 1)
    rgb, st = core.render(); print(name, 'counting: nodes/seg %.2f prims/seg %.2f'%(st.nodes_visited/st.segments, st.prims_tested/st.segments))
    core.close()
