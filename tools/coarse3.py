import sys; sys.argv=['x']
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import numpy as np, pickle, os
import asproto as A
from coarse import blocks_of
from coarse2 import pdas_lim
def build(H,g,ub,groups,nf):
    ng=len(groups); T=np.zeros((3*nf,3*ng))
    for j,grp in enumerate(groups):
        for f in grp: T[3*f:3*f+3,3*j:3*j+3]=np.eye(3)
    return T.T@H@T, T.T@g, np.array([min(ub[f] for f in grp) for grp in groups])
def run(probs,caps):
    steps=[]; fr=[]
    for (H,g,ub,mu_,links) in probs:
        nf=len(ub)
        g2=blocks_of(links,nf,2); g4=blocks_of(links,nf,4)
        H4,gg4,u4=build(H,g,ub,g4,nf); H2,gg2,u2=build(H,g,ub,g2,nf)
        tot=0
        if caps[0]>0:
            r4,s4,a4,_=pdas_lim(H4,gg4,u4,mu_,[0]*len(g4),caps[0]); tot+=sum((s+2)//3 for s in s4)
            m4={}
            for j,grp in enumerate(g4):
                for f in grp: m4[f]=a4[j]
            a20=[m4[grp[0]] for grp in g2]
        else: a20=[0]*len(g2)
        r2,s2,a2,_=pdas_lim(H2,gg2,u2,mu_,a20,caps[1]); tot+=sum((s+2)//3 for s in s2)
        acts0=[0]*nf
        for j,grp in enumerate(g2):
            for f in grp: acts0[f]=a2[j]
        rf,sf,_,_=pdas_lim(H,g,ub,mu_,acts0,40); tot+=sum((s+2)//3 for s in sf)
        steps.append(tot); fr.append(rf)
    print('caps',caps,'fine rounds',np.mean(fr),'block steps',np.mean(steps),flush=True)
probs=pickle.load(open('/tmp/probs_l3b.pkl','rb'))[:250]
for caps in ((0,4),(3,3),(4,2),(4,3),(6,3),(3,2),(16,2)):
    run(probs,caps)
