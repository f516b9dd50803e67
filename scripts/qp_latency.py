import sys, time; sys.path.insert(0,'.')
import numpy as np, asif_b200 as ab
H=np.diag([1.0,50.0]); c=np.array([[-0.4,-1000.0]]); A=np.random.default_rng(0).normal(size=(1,18,2)); b=-np.ones((1,18))*3
lb=np.array([-1.0,10.0]); ub=np.array([1.0,1e20])
for _ in range(20): ab.qp_solve_batch(H,c,A,b,lb,ub)
t=time.perf_counter(); N=2000
for _ in range(N): sol,st=ab.qp_solve_batch(H,c,A,b,lb,ub)
print('single-problem QPWrapper path: %.1f us per solve (incl. ctypes/numpy overhead)'%((time.perf_counter()-t)/N*1e6), sol, st)
