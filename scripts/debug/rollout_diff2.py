import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, conftest as cf
import asif_b200 as ab
from oracle import pyref
O=pyref.OracleLib()
n,dt=3000,1e-3
x0,ud=cf.c2_inputs(n,seed=91); x0*=0.4
opts=cf.C2_TB_OPTS
eng=ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(opts))
np.set_printoptions(precision=17, linewidth=200)
prev=None
for steps in (1,2,3,5,8,12,20,30,45,60):
    x,u,rc,h=eng.rollout(x0,ud,steps,dt)
    xo,uo,rco,ho=O.rollout(2,x0,ud,steps,dt,opts)
    dx=np.abs(x-xo).max(1); du=np.abs(u-uo)[:,0]
    k=int(dx.argmax())
    print('steps',steps,'max dx',dx.max(),'k',k,'max du',du.max(),'k',int(du.argmax()),'n dx>1e-9',int((dx>1e-9).sum()))
# single-state trace for the worst state
x,u,rc,h=eng.rollout(x0,ud,60,dt); xo,uo,rco,ho=O.rollout(2,x0,ud,60,dt,opts)
k=int(np.abs(x-xo).max(1).argmax())
xs=x0[k:k+1].copy(); 
for t in range(60):
    u0,r0,rc0,d0=O.filter_batch(2,xs,ud[k:k+1],opts,True)
    # gpu non-diag on a batch of copies
    xb=np.repeat(xs,64,0); ub=np.repeat(ud[k:k+1],64,0)
    ug,rg,rcg=eng.filter_batch(xb,ub)[:3]
    ug2,rg2,rcg2,dg=eng.filter_batch(xb,ub,want_diag=True)
    if abs(ug[0,0]-u0[0,0])>1e-9 or abs(ug2[0,0]-u0[0,0])>1e-9:
        print('t',t,'x',xs,'oracle',u0[0,0],r0[0,0],rc0[0],'gpu',ug[0,0],rg[0,0],rcg[0],'gpu diag',ug2[0,0],rg2[0,0])
        print(' A0',d0[0,8:26]); print(' A1',d0[0,26:44]); print(' b',d0[0,44:62]); print('ud',ud[k,0], 'rows equal', np.array_equal(dg[0],d0[0]))
        break
    xs=xs+dt*np.stack([xs[:,1],u0[:,0]],1)
print('done')
