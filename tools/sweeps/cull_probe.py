"""How much could distance culling of stale stack entries save on the soups?  CPU walk (product code, host build) of camera rays through the
wide BVH of an N Mi-triangle soup, first as the kernel does it, then with tmax preset to each ray's hit distance (= perfect culling: no node
beyond the hit is opened).  python tools/sweeps/cull_probe.py [N Mi]"""
import numpy as np, sys, time
sys.path.insert(0,'/root/repo')
import dsgpuraytracing_b200 as D
from dsgpuraytracing_b200 import scenes as S
from tests.cpuwalk import Walk
from oracle import oracle as O
n_mi = float(sys.argv[1]) if len(sys.argv)>1 else 1
sc,cam=S.triangle_soup(int(n_mi*(1<<20)), W=384, H=216)
t=time.time(); bvh=D.build_bvh2(sc); print('bvh2',time.time()-t)
t=time.time(); w=Walk(sc,bvh,1,camera=cam); print('wide',time.time()-t, w.info())
W,H=384,216
# camera rays through pixel centres (float)
ys,xs=np.mgrid[0:H:4,0:W:4]
o=[];d=[]
for y,x in zip(ys.ravel(),xs.ravel()):
    ro,rd=O.generate_ray(cam,(x+0.5)/W,(y+0.5)/H); o.append(ro); d.append(rd)
o=np.array(o,np.float32); d=np.array(d,np.float32)
ids,ts,cnt=w.trace(o,d)
hit=ids>=0
print('rays',len(o),'hit',hit.mean(),'nodes/ray',cnt[0]/len(o),'prims/ray',cnt[1]/len(o))
tm=np.where(hit, ts*(1+1e-5), np.float32(1e30)).astype(np.float32)
ids2,ts2,cnt2=w.trace(o,d,tmax=tm)
print('with tmax = hit distance: nodes/ray',cnt2[0]/len(o),'prims/ray',cnt2[1]/len(o), 'same ids', (ids2==ids).mean())
# only rays that hit
ids3,ts3,cnt3=w.trace(o[hit],d[hit]); ids4,ts4,cnt4=w.trace(o[hit],d[hit],tmax=tm[hit])
print('hit rays only: nodes/ray',cnt3[0]/hit.sum(),'->',cnt4[0]/hit.sum(),' prims/ray',cnt3[1]/hit.sum(),'->',cnt4[1]/hit.sum())
