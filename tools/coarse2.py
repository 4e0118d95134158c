import sys; sys.argv=['x']
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import numpy as np, pickle, os
import asproto as A
from coarse import blocks_of
def pdas_lim(H,g,ub,mu_,acts0,max_rounds,stop_changes=-1):
    acts=list(acts0); sizes=[]
    for rd in range(max_rounds):
        x,r,nred,vert=A.round_solve(H,g,ub,mu_,acts)
        sizes.append(nred)
        new=A.update_base(acts,x,r,ub,mu_,vert)
        nch=sum(a!=b for a,b in zip(new,acts))
        if new==acts: return rd+1,sizes,acts,True
        acts=new
        if nch<=stop_changes: return rd+1,sizes,acts,False
    return max_rounds,sizes,acts,False
def run(probs,cap,stop):
    steps=[]; fr=[]; cr=[]
    for (H,g,ub,mu_,links) in probs:
        nf=len(ub)
        groups=blocks_of(links,nf,2); ng=len(groups)
        T=np.zeros((3*nf,3*ng))
        for j,grp in enumerate(groups):
            for f in grp: T[3*f:3*f+3,3*j:3*j+3]=np.eye(3)
        Hc=T.T@H@T; gc=T.T@g; ubc=np.array([min(ub[f] for f in grp) for grp in groups])
        rc,sc,actc,_=pdas_lim(Hc,gc,ubc,mu_,[0]*ng,cap,stop)
        acts0=[0]*nf
        for j,grp in enumerate(groups):
            for f in grp: acts0[f]=actc[j]
        rf,sf,_,_=pdas_lim(H,g,ub,mu_,acts0,40)
        steps.append(sum((s+2)//3 for s in sc)+sum((s+2)//3 for s in sf)); fr.append(rf); cr.append(rc)
    print('cap',cap,'stop',stop,'coarse rounds',np.mean(cr),'fine rounds',np.mean(fr),'block steps',np.mean(steps),flush=True)
for name in ('/tmp/probs_l3b.pkl',):
    probs=pickle.load(open(name,'rb'))[:250]
    for cap,stop in ((3,-1),(4,-1),(5,-1),(16,-1),(16,1),(16,2),(16,3),(8,2),(6,2)):
        run(probs,cap,stop)
