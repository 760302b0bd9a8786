import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from hystrath_b200 import capi
from tests import helpers as H
eng = capi.Engine(0)
case, spd, fnum, vol = H.heatbath_setup(eng, scale=1.0)
eng.mesh_fill([0, 1], [case["numberDensities"]["N2"], case["numberDensities"]["O2"]], 30000.0, 30000.0, 30000.0)
eng.evolve(20)
eng.kernel_times(reset=True)
t=time.time(); eng.evolve(100); dt=time.time()-t
print("ms/step", dt*10, "parcels", eng.num_parcels(), eng.counters().collisions)
kt = eng.kernel_times()
print({k: round(v[0]/max(v[1],1),3) for k,v in kt.items()} if isinstance(kt, dict) else kt)
