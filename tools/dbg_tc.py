import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from superpoints_registration_b200 import ops
dev="cuda:0"
rng=np.random.default_rng(0)
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (ns,nq,H,c) in [(500,64,8,32),(500,64,16,32),(500,200,40,32),(700,333,23,64)]:
    s=rng.uniform(0,1,size=(ns,3)).astype(np.float32)
    q=s[:nq].copy()
    idx=rng.integers(0,ns,size=(nq,H))
    x=rng.normal(size=(ns,c)).astype(np.float32)
    w=(rng.normal(size=(15,c,c))/np.sqrt(15*c)).astype(np.float32)
    kp=(rng.normal(size=(15,3))*0.15).astype(np.float32)
    o1=ops.kpconv_forward(t(q),t(s),t(idx),t(x),t(w),t(kp),0.3,mode=1).cpu().numpy()
    o0=ops.kpconv_forward(t(q),t(s),t(idx),t(x),t(w),t(kp),0.3,mode=0).cpu().numpy()
    e=np.abs(o1-o0)
    print(ns,nq,H,c,"max err",e.max(),"scale",np.abs(o0).max(), "bad rows", (e.max(1)>1e-4).sum(), "bad cols", (e.max(0)>1e-4).sum())
    if e.max()>1e-4:
        print(" row err", np.round(e.max(1)[:20],3)); print(" col err", np.round(e.max(0),3))
# per-kernel-point probe: W nonzero only for one k, x = onehot channel
ns,nq,H,c=300,64,8,32
s=rng.uniform(0,1,size=(ns,3)).astype(np.float32); q=s[:nq].copy()
idx=rng.integers(0,ns,size=(nq,H)); x=rng.normal(size=(ns,c)).astype(np.float32)
kp=(rng.normal(size=(15,3))*0.15).astype(np.float32)
for k in range(15):
    w=np.zeros((15,c,c),np.float32); w[k]=np.eye(c)
    o1=ops.kpconv_forward(t(q),t(s),t(idx),t(x),t(w),t(kp),0.3,mode=1).cpu().numpy()
    o0=ops.kpconv_forward(t(q),t(s),t(idx),t(x),t(w),t(kp),0.3,mode=0).cpu().numpy()
    print("k",k,"err",np.abs(o1-o0).max(), "scale", np.abs(o0).max())
