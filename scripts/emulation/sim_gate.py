"""CPU emulation of three_nn_filtered_kernel's fp32 gate (csrc/pn2_interpolate.cu): the fp32 distance, the
round-up gate ru(ru(b3)*(1+2^-20))+1e-37 and the exact fp64 path, on adversarial inputs; asserts that the gate never
rejects a pair that would have entered the top 3 and that the result equals the oracle bit for bit.
Run: python scripts/emulation/sim_gate.py"""
import os
import numpy as np, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
f32=np.float32; f64=np.float64
def ru32(x64):
    r=x64.astype(f32)
    lo=r.astype(f64)<x64
    r=np.where(lo,np.nextafter(r,f32(np.inf)),r)
    return r.astype(f32)
def gate_of(b3):
    g=ru32(b3)
    g=ru32(g.astype(f64)*(1+2.0**-20))
    g=ru32(g.astype(f64)+f64(f32(1e-37)))
    return g
def sim(x1,x2):
    n=len(x1); m=len(x2)
    q=x1.astype(f32); q64=q.astype(f64)
    b=np.full((n,3),np.inf); bi=np.zeros((n,3),np.int32)
    gate=np.full(n,np.inf,f32)
    skipped=0
    for k in range(m):
        p=x2[k].astype(f32)
        dxf=(q[:,0]-p[0]).astype(f32); dyf=(q[:,1]-p[1]).astype(f32); dzf=(q[:,2]-p[2]).astype(f32)
        df=(dzf*dzf+((dyf*dyf).astype(f32)+(dxf*dxf).astype(f32)).astype(f32)).astype(f32)
        passed=df<=gate
        skipped+=(~passed).sum()
        d=q64-p.astype(f64)
        d=(d[:,0]*d[:,0]+d[:,1]*d[:,1])+d[:,2]*d[:,2]
        # check: any skipped pair that would have entered?
        bad=(~passed)&(d<b[:,2])
        assert not bad.any(), ("gate rejected a winner", k, df[bad], gate[bad], d[bad], b[bad,2])
        ins=passed&(d<b[:,2])
        for j in np.where(ins)[0]:
            dj=d[j]
            if dj<b[j,0]: b[j]=[dj,b[j,0],b[j,1]]; bi[j]=[k,bi[j,0],bi[j,1]]
            elif dj<b[j,1]: b[j]=[b[j,0],dj,b[j,1]]; bi[j]=[bi[j,0],k,bi[j,1]]
            else: b[j,2]=dj; bi[j,2]=k
            gate[j]=gate_of(np.array([b[j,2]]))[0]
    return b.astype(f32),bi,skipped/(n*m)
rs=np.random.RandomState(0)
cases={
 'uniform':(rs.random_sample((300,3)),rs.random_sample((600,3))),
 'offset1e3':(1000+1e-3*rs.random_sample((300,3)),1000+1e-3*rs.random_sample((600,3))),
 'offset1e5':(1e5+rs.random_sample((300,3)),1e5+rs.random_sample((600,3))),
 'lattice':(rs.randint(0,3,(200,3)).astype(float),rs.randint(0,3,(400,3)).astype(float)),
 'tiny':(1e-22*rs.random_sample((200,3)),1e-22*rs.random_sample((400,3))),
 'tiny2':(1e-19*rs.random_sample((200,3)),1e-19*rs.random_sample((400,3))),
 'dups':(np.repeat(rs.random_sample((50,3)),4,0),np.repeat(rs.random_sample((100,3)),5,0)),
 'huge':(1e18*rs.random_sample((100,3)),1e18*rs.random_sample((300,3))),
 'mixed':(rs.random_sample((200,3))*[1e-3,1,1e3],rs.random_sample((500,3))*[1e-3,1,1e3]),
}
for name,(a,b_) in cases.items():
    x1=a.astype(f32); x2=b_.astype(f32)
    d,i,sk=sim(x1,x2)
    ed,ei=orc.three_nn(x1[None],x2[None])
    ok=(i==ei[0]).all() and (d.view(np.uint32)==ed[0].view(np.uint32)).all()
    print(name, ok, 'skipped %.3f'%sk)
    assert ok
