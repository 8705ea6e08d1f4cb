"""CPU model of the traversal's distance-test volume on the c3 workload (no GPU needed).

For a sample of query leaves (32 Hilbert-consecutive atoms) it counts the candidate leaves, the targets that survive
the leaf-box filter, the true hits, and how many pair tests a filter per 16/8/4-atom sub-run of the leaf would leave.
DESIGN.md section 9 quotes the numbers.  Usage: python tools/leaf_filter_model.py"""
import sys

import numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from test_gpu_stages import hilbert30_numpy
from scipy.spatial import cKDTree
w = bench.make_workload("c3")
x = w["pos"].astype(np.float32); r = np.float32(w["cutoff"]); n=len(x)
key = hilbert30_numpy(x)
order = np.argsort(key, kind="stable")
xs = x[order].astype(np.float64)
nL = (n+31)//32
rng = np.random.default_rng(0)
sample = rng.choice(nL-1, 600, replace=False)
tree = cKDTree(xs)
def gap2(lo,hi,p):
    g = np.maximum(0, np.maximum(lo-p, p-hi)); return (g*g).sum(-1)
res = {k:[] for k in ("hits","T32","T16","T8","T4","cand_leaves","T8_jc4")}
r2=float(r)*float(r)
for A in sample:
    a0=A*32; q=xs[a0:a0+32]
    lo,hi=q.min(0),q.max(0)
    c=(lo+hi)/2; R=np.linalg.norm((hi-lo)/2)+float(r)
    idx=np.array(tree.query_ball_point(c,R)); idx=idx[idx>=a0+32]   # later leaves only (half list)
    p=xs[idx]
    near=gap2(lo,hi,p)<=r2
    # candidate leaves: leaf boxes near
    leaves=np.unique(idx//32)
    cl=0
    for B in leaves:
        pb=xs[B*32:B*32+32]; blo,bhi=pb.min(0),pb.max(0)
        g=np.maximum(0,np.maximum(lo-bhi,blo-hi)); cl+= (g*g).sum()<=r2
    res["cand_leaves"].append(cl)
    T32=near.sum()+0
    res["T32"].append(T32*32 + 32*31//2*0)   # pair tests (excluding self tile)
    d2=((q[:,None,:]-p[None,near,:])**2).sum(-1)
    res["hits"].append((d2<r2).sum())
    for g,name in ((16,"T16"),(8,"T8"),(4,"T4")):
        tot=0
        for s in range(0,32,g):
            qq=q[s:s+g]; l2,h2=qq.min(0),qq.max(0)
            tot+= (gap2(l2,h2,p[near])<=r2).sum()*g
        res[name].append(tot)
    # j clusters of 4 consecutive atoms (aligned), i clusters of 8: box-box test
    pj_idx=np.unique(idx[near]//4)
    tot=0
    for s in range(0,32,8):
        qq=q[s:s+8]; l2,h2=qq.min(0),qq.max(0)
        jb=xs[(pj_idx[:,None]*4+np.arange(4)[None,:]).clip(max=n-1)]
        jlo,jhi=jb.min(1),jb.max(1)
        g=np.maximum(0,np.maximum(l2-jhi,jlo-h2)); tot+=((g*g).sum(-1)<=r2).sum()*32
    res["T8_jc4"].append(tot)
for k,v in res.items(): print(k, np.mean(v))
h=np.mean(res["hits"])
for k in ("T32","T16","T8","T4","T8_jc4"): print(k,"tests/hit",np.mean(res[k])/h)
