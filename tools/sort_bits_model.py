"""CPU model: how many key bits does the per-step sort need?  Orders the c3 workload by the top B bits of the 30-bit
Hilbert key (worst case: random order inside a cell) and reports candidate leaves, surviving targets and leaf-box size
per query leaf.  DESIGN.md section 9 quotes the numbers.  Usage: python tools/sort_bits_model.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from test_gpu_stages import hilbert30_numpy
from scipy.spatial import cKDTree
w = bench.make_workload("c3")
x = w["pos"].astype(np.float32); r=float(w["cutoff"]); n=len(x); r2=r*r
key = hilbert30_numpy(x).astype(np.uint64)
rng=np.random.default_rng(1)
def model(order, label, nsamp=400):
    xs=x[order].astype(np.float64); nL=n//32
    tree=cKDTree(xs)
    lo=xs[:nL*32].reshape(nL,32,3).min(1); hi=xs[:nL*32].reshape(nL,32,3).max(1)
    ltree=None
    T=[];C=[];H=[]
    for A in rng.choice(nL-1,nsamp,replace=False):
        a0=A*32; q=xs[a0:a0+32]
        c=(lo[A]+hi[A])/2; R=np.linalg.norm((hi[A]-lo[A])/2)+r
        idx=np.array(tree.query_ball_point(c,R)); idx=idx[idx>=a0+32]
        p=xs[idx]
        g=np.maximum(0,np.maximum(lo[A]-p,p-hi[A])); near=(g*g).sum(1)<=r2
        T.append(near.sum())
        B=np.unique(idx//32); B=B[B<nL]
        gb=np.maximum(0,np.maximum(lo[A]-hi[B],lo[B]-hi[A])); C.append(((gb*gb).sum(1)<=r2).sum())
        d2=((q[:,None,:]-p[None,near,:])**2).sum(-1); H.append((d2<r2).sum())
    print(f"{label:40s} cand leaves {np.mean(C):6.1f}  targets {np.mean(T):7.1f}  hits {np.mean(H):6.1f}  box diag {np.mean(np.linalg.norm(hi-lo,axis=1))/r:5.2f} r")
for bits in (30,24,20,16,12):
    k=key>>np.uint64(30-bits)
    # worst case: random order inside a cell
    perm=rng.permutation(n); o=perm[np.argsort(k[perm],kind="stable")]
    model(o,f"top {bits} bits, random inside a cell")
k16=key>>np.uint64(14)
o_fine=np.argsort(key,kind="stable")
model(o_fine[np.argsort(k16[o_fine],kind="stable")],"top 16 bits, fine order kept inside")
