"""CPU emulation of the hashed-grid ball query (csrc/pn2_grouping.cu, ball_query_grid_kernel): cell function,
hash buckets with shuffled (atomics-like) order, padded cell ranges, exact-cell filter, hit cap + brute-force
fall-back, rank selection -- against the oracle.  Needs the built library for pn2_ball_threshold.
Run: python scripts/emulation/sim_grid.py"""
import os
import numpy as np, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
f32=np.float32
lib=ctypes.CDLL(os.path.join(ROOT, 'open3d-pointnet2-semantic3d_b200', 'lib', 'libpn2_b200.so'))
lib.pn2_ball_threshold.restype=ctypes.c_float; lib.pn2_ball_threshold.argtypes=[ctypes.c_float]
CL=f32(2.0**30)
def cellf(x,inv):
    v=np.floor((x.astype(f32)*inv).astype(f32))
    return np.clip(v,-CL,CL).astype(np.int64)
def hsh(c,T):
    u=lambda a: a.astype(np.int64)&0xFFFFFFFF
    h=(u(c[...,0])*73856093)^(u(c[...,1])*19349663)^(u(c[...,2])*83492791)
    return (h&0xFFFFFFFF)&(T-1)
def sqd(dx,dy,dz):
    dx=dx.astype(np.float64);dy=dy.astype(np.float64);dz=dz.astype(np.float64)
    t=f32(dy*dy).astype(np.float64)          # fmul rn
    t=f32(dx*dx+t).astype(np.float64)        # fma (exact product + add, single rounding): double has 53 bits; product of two f32 is exact in double; sum rounding double then float = double rounding risk (tiny) -> acceptable for emulation
    return f32(dz*dz+t)
def grid_ball(radius,ns,x1,x2,cap=512):
    n=len(x1); m=len(x2)
    thr=f32(lib.pn2_ball_threshold(ctypes.c_float(radius)))
    r=f32(radius); inv=f32(1.0)/r
    rpad=np.nextafter(f32(r*f32(1.0001)),f32(np.inf))
    T=1
    while T<2*n: T*=2
    cells=np.stack([cellf(x1[:,a],inv) for a in range(3)],-1)
    bucket=hsh(cells,T)
    order=np.argsort(bucket,kind='stable')
    # emulate atomics: shuffle within bucket
    rs=np.random.RandomState(0)
    starts=np.searchsorted(bucket[order],np.arange(T+1))
    for bkt in np.unique(bucket):
        s,e=starts[bkt],starts[bkt+1]
        seg=order[s:e].copy(); rs.shuffle(seg); order[s:e]=seg
    idx=np.zeros((m,ns),np.int32); cnt=np.zeros(m,np.int32); fallback=0
    for j in range(m):
        q=x2[j]
        lo=[cellf(np.array([np.nextafter(f32(q[a]-rpad),f32(-np.inf))]),inv)[0] for a in range(3)]
        hi=[cellf(np.array([np.nextafter(f32(q[a]+rpad),f32(np.inf))]),inv)[0] for a in range(3)]
        hits=[]
        if max(hi[a]-lo[a]+1 for a in range(3))>4:
            fallback+=1; hits=None
        else:
            for cx in range(lo[0],hi[0]+1):
                for cy in range(lo[1],hi[1]+1):
                    for cz in range(lo[2],hi[2]+1):
                        bk=hsh(np.array([cx,cy,cz]),T)
                        cand=order[starts[bk]:starts[bk+1]]
                        if len(cand)==0: continue
                        same=(cells[cand]==np.array([cx,cy,cz])).all(1)
                        cand=cand[same]
                        p=x1[cand]
                        d=sqd(q[0]-p[:,0],q[1]-p[:,1],q[2]-p[:,2])
                        hits+=cand[~(d>=thr)].tolist()
            if len(hits)>cap: fallback+=1; hits=None
        if hits is None:
            d=sqd(q[0]-x1[:,0],q[1]-x1[:,1],q[2]-x1[:,2])
            hits=np.where(~(d>=thr))[0].tolist()
        hits=sorted(hits)[:ns]
        cnt[j]=len(hits)
        if hits:
            idx[j,:len(hits)]=hits; idx[j,len(hits):]=hits[0]
    return idx,cnt,fallback
rs=np.random.RandomState(1)
for name,(n,m,radius,ns,scale,shift) in {
  'cfg2':(4096,300,0.5,32,(10,10,5),(-5,-5,0)),
  'unit':(3000,300,0.2,32,(1,1,1),(0,0,0)),
  'dense':(3000,200,0.35,16,(1,1,1),(0,0,0)),
  'big_r':(1000,100,4.0,32,(10,10,5),(-5,-5,0)),
  'offset':(3000,200,0.2,32,(1,1,1),(1000,-2000,50)),
  'tiny_r':(2000,200,1e-3,8,(1,1,1),(0,0,0)),
  'lattice':(3000,200,1.0,32,None,None),
}.items():
    if scale is None:
        x1=rs.randint(0,8,(n,3)).astype(f32)
    else:
        x1=(rs.random_sample((n,3))*scale+shift).astype(f32)
    x2=x1[rs.choice(n,m,replace=False)].copy()
    x2[:5]+=f32(0.013)
    gi,gc,fb=grid_ball(radius,ns,x1,x2)
    ei,ec=orc.query_ball_point(radius,ns,x1[None],x2[None])
    ok=(gi==ei[0]).all() and (gc==ec[0]).all()
    print(name,ok,'fallbacks',fb,'mean cnt %.1f'%gc.mean())
    assert ok
