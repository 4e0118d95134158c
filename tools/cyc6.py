import sys; sys.argv=['x']
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import numpy as np, pickle, os
import asproto as A
probs=pickle.load(open('/tmp/probs_l3.pkl','rb'))
def dscore(f, acts, r, mu_):
    im=1.0/mu_; rf=r[3*f:3*f+3]; act=acts[f]
    a0,a1,a2,a3=act&1,(act>>1)&1,(act>>2)&1,(act>>3)&1
    if a0+a1==2 or a2+a3==2 or a0+a1+a2+a3>=3: return rf[2]-(abs(rf[0])+abs(rf[1]))*im
    lx = rf[0]*im if act & 1 else (-rf[0]*im if act & 2 else 0.0)
    ly = rf[1]*im if act & 4 else (-rf[1]*im if act & 8 else 0.0)
    lc = (lx+ly-rf[2]) if act & 16 else 0.0
    return min(lx if act&3 else 0, ly if act&12 else 0, lc if act&16 else 0)
def ascore(f, acts, x, ub, mu_):
    fx,fy,fz=x[3*f:3*f+3]
    c=[mu_*fx+fz,-mu_*fx+fz,mu_*fy+fz,-mu_*fy+fz,ub[f]-fz]
    return min(c)
def restrict(mode, acts, new, x, r, ub, mu_):
    out=list(new)
    drops=[f for f in range(len(acts)) if new[f]!=acts[f] and (new[f]|acts[f])!=new[f]]
    adds=[f for f in range(len(acts)) if new[f]!=acts[f] and (new[f]|acts[f])==new[f]]
    bd = min(drops,key=lambda f:dscore(f,acts,r,mu_)) if drops else None
    ba = min(adds,key=lambda f:ascore(f,acts,x,ub,mu_)) if adds else None
    if mode=='onedrop':
        keep=set(adds)|({bd} if bd is not None else set())
    elif mode=='one_dropfirst':
        keep={bd} if bd is not None else {ba}
    elif mode=='one_addfirst':
        keep={ba} if ba is not None else {bd}
    elif mode=='alladds_else_onedrop':
        keep=set(adds) if adds else {bd}
    elif mode=='oneadd_onedrop':
        keep={ba,bd}
    for f in range(len(acts)):
        if f not in keep: out[f]=acts[f]
    return out
def solve_cd(H, g, ub, mu_, links, mode, max_rounds=40):
    nf = len(ub); acts = [0]*nf; hist=[None]*16; cyc = False; rdet = None
    for rd in range(max_rounds):
        x, r, nred, vert = A.round_solve(H, g, ub, mu_, acts)
        new = A.update_base(acts, x, r, ub, mu_, vert)
        if new == acts: return rd + 1, rdet
        if cyc: new = restrict(mode, acts, new, x, r, ub, mu_)
        else:
            t = tuple(new)
            if any(hist[j] == t for j in range(16) if j != (rd & 15)):
                cyc = True; rdet = rd
            hist[rd & 15] = t
        acts = new
    return max_rounds, rdet
import glob
sets=[('/tmp/probs_l3.pkl',[29,338,366,520,720,802,927])]
for mode in ('onedrop','one_dropfirst','one_addfirst','alladds_else_onedrop','oneadd_onedrop'):
    out=[]
    for i in sets[0][1]:
        rd,rdet=solve_cd(*probs[i],mode)
        out.append((i,rdet,rd))
    print(mode,out)
