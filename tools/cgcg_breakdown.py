"""numpy reproduction of the single-reduction (Chronopoulos-Gear) PCG breakdown on the 8-rank check bar: classic
PCG converges in 1039 iterations, the single-reduction recurrences hit a non-positive step denominator at 220."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, scipy.sparse as sp
from multirank_worker import small_bar
from oracle.oracle import PortOracle
for world in (4,8):
    m=small_bar(world)
    rng=np.random.default_rng(3); x0=m.nodes+0.004*rng.standard_normal(m.nodes.shape)
    o=PortOracle(m); o.set_nodes(x0); o.apply_increment(1.0); o.update_state(); o.assemble_stiffness(); o.assemble_residual(); o.apply_bc(0.0)
    rp,ci,v=o.get_csr(); A=sp.csr_matrix((v,ci,rp)); b=o.get_forces(); N=A.shape[0]; dinv=1/A.diagonal()
    def pcg(tol=1e-13):
        x=np.zeros(N); r=b.copy(); z=dinv*r; p=z.copy(); rz=r@z
        for k in range(100000):
            q=A@p; al=rz/(p@q); x+=al*p; r-=al*q
            if np.linalg.norm(r)<=tol*np.linalg.norm(b): return x,k+1
            z=dinv*r; rz2=r@z; p=z+(rz2/rz)*p; rz=rz2
    def cgcg(tol=1e-13):
        x=np.zeros(N); r=b.copy(); z=dinv*r; w=A@z
        gam=r@z; dl=w@z; rr=r@r; p=np.zeros(N); s=np.zeros(N); gam_old=al_old=None; hist=[]
        for k in range(100000):
            hist.append(np.sqrt(rr/(b@b)))
            if rr<=tol**2*(b@b): return x,k,hist
            if k==0: be=0.0; den=dl
            else: be=gam/gam_old; den=dl-be*gam/al_old
            if not den>0: return x,-k,hist
            al=gam/den
            p=z+be*p; s=w+be*s; x=x+al*p; r=r-al*s; z=dinv*r
            gam_old,al_old=gam,al
            gam=r@z; rr=r@r; w=A@z; dl=w@z
    x1,k1=pcg(); x2,k2,h=cgcg()
    print(world,N,k1,k2,h[min(220,len(h)-1)], np.abs(x1-x2).max()/np.abs(x1).max())
