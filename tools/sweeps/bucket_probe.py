"""Would a finer SAH sweep give a better tree?  The host builder with 32 buckets (the reference's count, the default) against a build
compiled with -DDSRT_SAH_BUCKETS=256 (close to a full sweep), collapsed to the wide layout and walked on the CPU (product code):
  python tools/sweeps/bucket_probe.py c2|<Ki triangles of the soup>          # DSRT_LIB=.../libdsrt_b256.so for the other build
(make -C dsgpuraytracing_b200/csrc ../libdsrt_b256.so OUT=../libdsrt_b256.so EXTRA=-DDSRT_SAH_BUCKETS=256)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
from tests.cpuwalk import Walk
which = sys.argv[1]
if which == 'c2':
    sc, cam = S.cbdragon_standin(192, 108); nl = 4
else:
    sc, cam = S.triangle_soup(int(which) << 10, W=192, H=108); nl = 1
bvh = D.build_bvh2(sc)
w = Walk(sc, bvh, nl, camera=cam)
rgb, c = w.render(2, 8, seed=3)
seg = float(c[1] + c[2])
print(os.path.basename(D.lib_path()), which, 'bvh2 nodes', len(bvh['node_start']), 'wide', w.info(), 'nodes/seg %.3f prims/seg %.3f' % (c[3] / seg, c[4] / seg))
