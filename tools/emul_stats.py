import sys, numpy as np
_R = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, _R + '/tests')
import _pkg; pkg=_pkg.load()
from quadruped_robot_b200 import capi
import emul_binding as EB
em=EB.load()
for robot,h,B,seed,gait in (('lite3',10,1500,3,'trot'),('a1',10,1500,5,'trot'),('aliengo',10,1000,13,'mixed'),('lite3',5,1500,4,'trot')):
    batch=pkg.synth.make_mpc_batch(robot,h,0.03,B,seed=seed,gait=gait)
    P=capi.params_of(batch["robot"],h,0.03)
    r=em.solve(P,batch)
    it=r['iters']
    print(robot,h,gait,'status!=0',int((r['status']!=0).sum()),'ipm instances',int((it[:,0]>0).sum()),'ipm iters mean(those)',float(it[it[:,0]>0,0].mean()) if (it[:,0]>0).any() else 0,'rounds mean',it[:,1].mean(),'max',it[:,1].max(), 'hist>=14', int((it[:,1]>=14).sum()))
