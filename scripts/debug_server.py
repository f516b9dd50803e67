import faulthandler, sys, os, time
faulthandler.dump_traceback_later(25, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest as cf
import asif_b200 as ab
which = sys.argv[1]
def P(*a): print(*a, flush=True)
if which == "explicit":
    eng = ab.Engine(ab.FILTER_EXPLICIT, ab.MODEL_DOUBLE_INTEGRATOR, relaxLb=cf.C1_OPTS[0], relaxCost=cf.C1_OPTS[1])
    x, ud = cf.c1_inputs(4000, seed=321)
else:
    eng = ab.Engine(ab.FILTER_IMPLICIT_TB, ab.MODEL_DOUBLE_INTEGRATOR_TB, **cf.tb_engine_kwargs(cf.C2_TB_OPTS))
    x, ud = cf.c2_inputs(4000, seed=322)
want = eng.filter_batch(x, ud); P("base ok")
eng.latency_server(True); P("server on")
got = eng.filter_batch(x[:1], ud[:1]); P("1 state", got[0][0], want[0][0])
got = eng.filter_batch(x[2:34], ud[2:34]); P("32 states", np.array_equal(got[0], want[0][2:34]))
got = eng.filter_batch(x[66:97], ud[66:97]); P("31 states", np.array_equal(got[0], want[0][66:97]))
t=time.time()
for k in range(200, 700): got = eng.filter_batch(x[k:k+1], ud[k:k+1])
P("500 single calls", (time.time()-t)/500*1e6, "us each")
big = eng.filter_batch(x, ud); P("big batch ok", np.array_equal(big[0], want[0]))
d1 = eng.filter_batch(x[:8], ud[:8], want_diag=True); P("diag small ok")
eng.latency_server(False); P("server off")
eng.close(); P("closed")
