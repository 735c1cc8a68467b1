"""CPU check (not product code): the formulation the draft multilevel kernels use in the block-Jacobi-SCALED system,
z~ = r~ + L^T P Mc P^T L r~ on K~ = S K S^T, produces the iterates of the unscaled PCG with M = D^-1 + P Mc P^T (x = S^T y)."""
import os, sys, warnings
import numpy as np, scipy.sparse as sp
warnings.filterwarnings("ignore", category=DeprecationWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import meshbeamfea_b200 as mb
from meshbeamfea_b200 import synthetic as syn
from oracle import oracle as orc
from helpers import to_oracle
job=syn.cubic_lattice(8); O=to_oracle(job); V=O.nodes.shape[0]
fm=orc.free_map(V,O.fixed); rp,col=orc.bsr_pattern(O,fm); vals=orc.bsr_assemble(O,fm,rp,col).reshape(-1,6,6); nb=rp.size-1
K=sp.bsr_matrix((vals,col,rp),shape=(6*nb,6*nb)).tocsr()
f=orc.load_vector(V,O.forces).reshape(V,6)[fm>=0].ravel()
H=mb.host_hierarchy(job.nodes,job.elements,job.fixedNodes,cell0=0.2,ratio=2.0,max_levels=2,coarsest_rows=1)
Lv=H[0]; off=Lv["off"]; agg=Lv["agg"]
B=np.tile(np.eye(6),(nb,1,1)); x,y,z=off[:,0],off[:,1],off[:,2]
B[:,0,4],B[:,0,5],B[:,1,3],B[:,1,5],B[:,2,3],B[:,2,4]=z,-y,-z,x,y,-x
P=sp.bsr_matrix((B,agg,np.arange(nb+1)),shape=(6*nb,6*Lv["n_coarse"])).tocsr()
Ac=(P.T@K@P).toarray(); Aci=np.linalg.inv(Ac)
D=np.array([vals[rp[r]+np.searchsorted(col[rp[r]:rp[r+1]],r)] for r in range(nb)])
Lc=np.linalg.cholesky(D)                       # D = L L^T
Lbd=sp.block_diag([Lc[r] for r in range(nb)]).tocsr(); Sbd=sp.block_diag([np.linalg.inv(Lc[r]) for r in range(nb)]).tocsr()
Kt=(Sbd@K@Sbd.T).tocsr(); ft=Sbd@f
def pcg(A,b,prec,its):
    x=np.zeros_like(b); r=b.copy(); z=prec(r); p=z.copy(); rz=r@z; out=[]
    for _ in range(its):
        Ap=A@p; a=rz/(p@Ap); x=x+a*p; r=r-a*Ap; z=prec(r); rz2=r@z; p=z+(rz2/rz)*p; rz=rz2; out.append(x.copy())
    return out
Dinv=sp.block_diag([np.linalg.inv(D[r]) for r in range(nb)]).tocsr()
xs=pcg(K,f,lambda r: Dinv@r + P@(Aci@(P.T@r)),25)
ys=pcg(Kt,ft,lambda r: r + Lbd.T@(P@(Aci@(P.T@(Lbd@r)))),25)     # the draft kernels' formulation
print("max |x_k - S^T y_k| / |x_k| over 25 iterations:", max(np.linalg.norm(xk-Sbd.T@yk)/np.linalg.norm(xk) for xk,yk in zip(xs,ys)))
