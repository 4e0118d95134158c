import sys; sys.argv=['x']
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import numpy as np, pickle, os
import asproto as A

def pdas(H,g,ub,mu_,acts0,max_rounds=40):
    acts=list(acts0); sizes=[]
    for rd in range(max_rounds):
        x,r,nred,vert=A.round_solve(H,g,ub,mu_,acts)
        sizes.append(nred)
        new=A.update_base(acts,x,r,ub,mu_,vert)
        if new==acts: return rd+1,sizes,acts,x
        acts=new
    return max_rounds,sizes,acts,x

def blocks_of(links, nf, bs):
    # chains: follow links from heads
    has_prev=[False]*nf
    for f in range(nf):
        if links[f]>=0: has_prev[links[f]]=True
    groups=[]
    for f in range(nf):
        if not has_prev[f]:
            chain=[f]; n=links[f]
            while n>=0: chain.append(n); n=links[n]
            for i in range(0,len(chain),bs): groups.append(chain[i:i+bs])
    return groups

def run(probs,bs):
    tot_r=[]; tot_cost=[]; base_cost=[]; base_r=[]; cr=[]
    for (H,g,ub,mu_,links) in probs:
        nf=len(ub)
        rb,sb,_,xb=pdas(H,g,ub,mu_,[0]*nf)
        base_r.append(rb); base_cost.append(sum(A.tiles((s+2)//3) for s in sb))
        groups=blocks_of(links,nf,bs)
        ng=len(groups)
        T=np.zeros((3*nf,3*ng))
        for j,grp in enumerate(groups):
            for f in grp:
                T[3*f:3*f+3,3*j:3*j+3]=np.eye(3)
        Hc=T.T@H@T; gc=T.T@g; ubc=np.array([min(ub[f] for f in grp) for grp in groups])
        rc,sc,actc,xc=pdas(Hc,gc,ubc,mu_,[0]*ng)
        acts0=[0]*nf
        for j,grp in enumerate(groups):
            for f in grp: acts0[f]=actc[j]
        rf,sf,_,xf=pdas(H,g,ub,mu_,acts0)
        if rf<40 and rb<40 and not np.allclose(xf,xb,atol=1e-5): print('MISMATCH',np.abs(xf-xb).max())
        cr.append(rc)
        tot_r.append(rf); tot_cost.append(sum(A.tiles((s+2)//3) for s in sc)+sum(A.tiles((s+2)//3) for s in sf))
    print('bs',bs,'base rounds',np.mean(base_r),'tiles',np.mean(base_cost),'| coarse rounds',np.mean(cr),'fine rounds',np.mean(tot_r),'max',np.max(tot_r),'tiles total',np.mean(tot_cost))
probs=pickle.load(open('/tmp/probs_l3b.pkl','rb'))[:150]
for bs in (2,3,5): run(probs,bs)
