import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oxmpl_b200 as oxg
rng=np.random.default_rng(0)
n,d=1000000,7
st=oxg.StateStore(oxg.Space.real_vector(d),capacity=n); st.append(rng.random((n,d)))
for Q in (100,1000,10000,50000):
    q=rng.random((Q,d))
    for k in (1,10):
        ts=[]
        for _ in range(3):
            t0=time.perf_counter(); st.knn(q,k); ts.append(time.perf_counter()-t0)
        print("Q",Q,"k",k,"dev ms",round(st.last_scan_ms(),3), flush=True)
