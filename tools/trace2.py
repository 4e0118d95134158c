import sys, ctypes as C, numpy as np
_R = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, _R + '/tests')
import _pkg
pkg=_pkg.load()
from quadruped_robot_b200 import capi
import emul_binding as EB
lib=C.CDLL(_R + '/scratch/libtrace.so')   # g++ -O2 -std=c++17 -shared -fPIC -o scratch/libtrace.so tools/trace.cpp
em=EB.Emul(lib)
def tr(robot,h,B,seed,gait,idxs):
    batch=pkg.synth.make_mpc_batch(robot,h,0.03,B,seed=seed,gait=gait)
    P=capi.params_of(batch["robot"],h,0.03)
    nred=np.zeros(64,np.int32); acts=np.zeros(64*64,np.int32); rounds=C.c_int()
    for i in idxs:
        one={k:(v[i:i+1] if isinstance(v,np.ndarray) and v.shape[:1]==(B,) else v) for k,v in batch.items()}
        r=em.solve(P,one)
        lib.trace_get(nred.ctypes.data_as(C.POINTER(C.c_int)),acts.ctypes.data_as(C.POINTER(C.c_int)),C.byref(rounds))
        nr=rounds.value; A=acts.reshape(64,64)
        nf=int((batch["gait"][i]>0).sum())
        print(robot,i,'iters',r['iters'].tolist(),'traced rounds',nr)
        for rr in range(nr): print('   ',''.join('%x'%A[rr,f] if A[rr,f]<16 else chr(ord('A')+A[rr,f]-16) for f in range(nf)), nred[rr])
tr('lite3',10,4096,3,'trot',[24])
tr('a1',10,16384,0,'trot',[822])
