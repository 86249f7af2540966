import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import bench, clann_b200 as cb
from clann_b200 import _lib as cl
class A: small=False; workload='glove100'
w = bench.workload(A)
data, q, src = bench.make_data(w, 'planted')
for rep in range(4):
    t=time.time()
    ix = cb.init_with_config(data, cb.Config(w["L"], w["factor"], w["k"], w["delta"], "b"))
    ix.set_option("seed", 1234)
    t1=time.time()
    ix.build()
    print("rep", rep, "init %.3f build wall %.3f"%(t1-t, time.time()-t1), "ms [gmm, hash, sort, total]:", ix.export(cl.X_BUILD_MS, 0, np.float64))
    ix.close()
