import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, conftest as cf
import asif_b200 as ab
from oracle import pyref
O=pyref.OracleLib()
x,ud=cf.c3a_inputs(400, seed=cf.SEED+31)
eng = ab.Engine(ab.FILTER_IMPLICIT, ab.MODEL_INVERTED_PENDULUM, **cf.implicit_engine_kwargs(cf.C3A_OPTS))
u, relax, rc, diag = eng.filter_batch(x, ud, want_diag=True)
u0,relax0,rc0,diag0=O.filter_batch(3,x,ud,cf.C3A_OPTS,True)
np.set_printoptions(linewidth=220, precision=10)
i=18
print('x',x[i],'ud',ud[i]); print('gpu',rc[i],u[i],relax[i]); print('orc',rc0[i],u0[i],relax0[i])
print('crit gpu',diag[i,2:12]); print('crit orc',diag0[i,2:12]); print('hS,hBE',diag[i,:2],diag0[i,:2])
A=diag[i,12:12+123].reshape(3,41).T; b=diag[i,135:176]; A0=diag0[i,12:12+123].reshape(3,41).T; b0=diag0[i,135:176]
print('max dA',np.abs(A-A0).max(),'max db',np.abs(b-b0).max())
print(np.hstack([A,b[:,None]])[-6:]); print(np.hstack([A0,b0[:,None]])[-6:])
dd=np.abs(diag-diag0).max(axis=1); print('worst diag diffs', np.argsort(-dd)[:5], np.sort(dd)[-5:])
v=np.array([u[i,0],relax[i,0],relax[i,1]]); v0=np.array([u0[i,0],relax0[i,0],relax0[i,1]])
print('slack gpu sol min', (A0@v-b0).min(), 'orc', (A0@v0-b0).min(), 'obj gpu', (v[0]-ud[i,0])**2+50*(v[1]-10)**2+50*(v[2]-5)**2,'obj orc',(v0[0]-ud[i,0])**2+50*(v0[1]-10)**2+50*(v0[2]-5)**2)
