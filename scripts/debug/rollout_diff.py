import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, conftest as cf
import asif_b200 as ab
from oracle import pyref
O=pyref.OracleLib()
n,steps,dt=3000,60,1e-3
x0,ud=cf.c2_inputs(n,seed=91); x0*=0.4
opts=cf.C2_TB_OPTS
eng=ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(opts))
x=x0.copy(); np.set_printoptions(precision=17, linewidth=200)
shown=0
for t in range(steps):
    u0,r0,rc0,d0=O.filter_batch(2,x,ud,opts,True)
    u,r,rc,d=eng.filter_batch(x,ud,want_diag=True)
    diff=np.abs(u-u0)[:,0]
    bad=np.where(diff>1e-7)[0]
    for k in bad[:3]:
        if shown<6:
            shown+=1
            print('t',t,'k',k,'gpu',u[k,0],r[k,0],rc[k],'oracle',u0[k,0],r0[k,0],rc0[k],'ud',ud[k,0])
            print(' rows A0',d0[k,8:26]); print(' rows A1',d0[k,26:44]); print(' b',d0[k,44:62]); print(' rows equal',np.array_equal(d[k],d0[k]))
    x=x+dt*np.stack([x[:,1],u0[:,0]],1)
print('done')
