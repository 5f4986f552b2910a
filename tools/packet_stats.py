"""CPU estimate behind the frustum front end: for random pixel blocks of several sizes, how many\nBVH nodes and leaves the UNION of the block's rays enters, against what one ray enters alone.  numpy only."""
import sys, numpy as np, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opencl_raytracer_b200 import scene as scn, scenes
v,f = scenes.sibenik_standin(); sc = scn.scene_from_mesh(v,f)
nodes = sc.nodes; lo = sc.aabbs[0::2,:3].astype(np.float32); hi = sc.aabbs[1::2,:3].astype(np.float32)
def rays(W,H,x0,y0,pw,ph):
    a = np.float32(max(W,H))
    xs = np.arange(x0,x0+pw,dtype=np.float32); ys = np.arange(y0,y0+ph,dtype=np.float32)
    X,Y = np.meshgrid(xs,ys)
    dx = (X+0.5)/a - W/(2*a); dy = -((Y+0.5)/a - H/(2*a)); dz = -np.ones_like(dx)
    d = np.stack([dx,dy,dz],-1).reshape(-1,3); d /= np.linalg.norm(d,axis=1,keepdims=True)
    return d.astype(np.float32)
def slab(i, d, idr):
    o = np.array([0,0,2],np.float32)
    t0 = (lo[i]-o)*idr; t1 = (hi[i]-o)*idr
    tmin = np.minimum(t0,t1).max(1); tmax = np.maximum(t0,t1).min(1)
    return (tmin<=tmax)&(tmax>0)&(tmin<1e5)
def packet(W,H,x0,y0,pw,ph):
    d = rays(W,H,x0,y0,pw,ph); idr = 1/d
    stack=[0]; nvis=0; leaves=0; cand=0
    while stack:
        i = stack.pop(); nvis+=1
        h = slab(i,d,idr)
        if not h.any(): continue
        if nodes[i]==1: leaves+=1; cand+=int(h.sum()); continue
        l=i+1; r=l+nodes[l]; stack.append(r); stack.append(l)
    return nvis, leaves, cand/ (pw*ph)
rng=np.random.default_rng(0)
for (W,H,name) in ((3840,2160,'C2'),(15360,8640,'C3')):
  for (pw,ph) in ((8,4),(16,8),(32,16),(32,32)):
    res=[]
    for _ in range(150):
        x0 = int(rng.integers(0,W//pw))*pw; y0=int(rng.integers(0,H//ph))*ph
        res.append(packet(W,H,x0,y0,pw,ph))
    r=np.array(res)
    print(name,(pw,ph),'nodes/packet %.0f (max %d)  leaves/packet %.1f (max %d)  per-ray leafbox passes %.2f'%(r[:,0].mean(),r[:,0].max(),r[:,1].mean(),r[:,1].max(),r[:,2].mean()))
