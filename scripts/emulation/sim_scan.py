"""CPU emulation of prob_cdf_kernel's data flow (csrc/pn2_sampling.cu): quad pairs per thread, warp-shuffle
tree levels, one warp over the 32 warp totals, the down-sweep ladder, the compensated carry -- in numpy float32,
lane for lane, against the oracle's prefix sum (which is itself pinned against a simulation of the reference's
shared-memory scan, tests/test_prob_sample_cpu.py).  Run: python scripts/emulation/sim_scan.py"""
import os
import numpy as np, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
f32=np.float32
def shfl_up(a, d):
    # a: (32 warps, 32 lanes)
    o=a.copy(); o[:,d:]=a[:,:-d]; return o
def kernel_sim(row):
    n=len(row); out=np.zeros(n,f32); run=f32(0); comp=f32(0)
    T=1024
    lane=np.arange(T)%32; w=np.arange(T)//32
    L=lane.reshape(32,32); W=w.reshape(32,32)
    for j in range(0,n,8192):
        ln=min(n-j,8192); nq=(ln+3)>>2
        e=np.zeros((T,2,4),f32); qt=np.zeros((T,2),f32)
        for t in range(T):
            for h in range(2):
                q=2*t+h; k=4*q
                if k+3<ln:
                    v=row[j+k:j+k+4].astype(f32)
                    ba=f32(v[1]+v[0]); dc=f32(v[3]+v[2])
                    e[t,h]=(v[0],ba,f32(v[2]+ba),f32(dc+ba)); qt[t,h]=e[t,h,3]
                elif k<ln:
                    acc=f32(0)
                    for i in range(4):
                        if k+i<ln: acc=f32(acc+row[j+k+i])
                        e[t,h,i]=acc
                    qt[t,h]=acc
        a=(qt[:,1]+qt[:,0]).astype(f32).reshape(32,32)
        d=1
        while d<32:
            o=shfl_up(a,d); m=((L+1)&(2*d-1))==0
            a=np.where(m,(a+o).astype(f32),a); d<<=1
        wt=a[:,31].copy()
        # warp 0
        x=wt.reshape(1,32).copy(); l0=np.arange(32).reshape(1,32)
        d=1
        while d<32:
            o=shfl_up(x,d); m=((l0+1)&(2*d-1))==0
            x=np.where(m,(x+o).astype(f32),x); d<<=1
        d=8
        while d>=1:
            o=shfl_up(x,d); m=(((l0+1)&(2*d-1))==d)&(l0+1>d)
            x=np.where(m,(x+o).astype(f32),x); d>>=1
        wt=x.reshape(32)
        Wp=np.zeros((32,32),f32); Wp[1:,:]=wt[:-1,None]
        a[:,31]=wt
        d=16
        while d>=1:
            o=shfl_up(a,d); m=((L+1)&(2*d-1))==d
            inside=m&(L+1>d); edge=m&(L+1==d)&(W>0)
            a=np.where(inside,(a+o).astype(f32),np.where(edge,(a+Wp).astype(f32),a)); d>>=1
        pprev=shfl_up(a,1); pprev[:,0]=Wp[:,0]
        a=a.reshape(T); pprev=pprev.reshape(T)
        total=None
        for t in range(T):
            has_prev=t>0
            pre0 = f32(qt[t,0]+pprev[t]) if has_prev else qt[t,0]
            for h in range(2):
                q=2*t+h
                for i in range(4):
                    k=4*q+i
                    if k<ln:
                        v=e[t,h,i]
                        if h==0:
                            if has_prev: v=f32(v+pprev[t])
                        else: v=f32(v+pre0)
                        out[j+k]=f32(v+run)
            if t==(nq-1)>>1:
                total = a[t] if ((nq-1)&1) else pre0
        tt=f32(total+comp); r2=f32(run+tt); comp=f32(tt-f32(r2-run)); run=r2
    return out
for n in [1,2,3,4,5,7,8,9,63,64,65,100,255,256,257,1023,1024,1025,4099,8191,8192,8193,8200,12345,16389]:
    rs=np.random.RandomState(n)
    x=(rs.random_sample((1,n))*rs.choice([1e-3,1.0,37.0],size=(1,n))).astype(f32)
    a=orc.cumsum(x)[0]; b=kernel_sim(x[0])
    ok=(a.view(np.uint32)==b.view(np.uint32)).all()
    print(n, ok)
    assert ok
