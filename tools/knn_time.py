"""near.obs kernel time vs station count (profile class 5) and wall time of the host call."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import rfsi_b200 as rb
from rfsi_b200 import synth, _lib
lib = _lib.load()
rng = np.random.default_rng(0)
loc = synth.lattice(500, 500)
def kern_ms(fn, cls=5):
    fn()
    lib.rfsi_profile_enable(1); fn(); lib.rfsi_profile_enable(0)
    t, n = C.c_double(0), C.c_uint64(0)
    lib.rfsi_profile_read(cls, C.byref(t), C.byref(n))
    return t.value, n.value
for S, n in ((100, 25), (500, 40), (1000, 25), (1023, 25), (1024, 25), (5000, 25), (20000, 25)):
    obs = loc[rng.choice(250000, S, replace=False)]
    odf = np.column_stack([obs, rng.normal(size=S)])
    t0 = time.perf_counter(); rb.near_obs(loc, odf, zcol=2, n_obs=n); wall = time.perf_counter() - t0
    ms, k = kern_ms(lambda: rb.near_obs(loc, odf, zcol=2, n_obs=n))
    print(f"S={S:6d} n={n:3d}: kernel {ms:8.3f} ms ({k} launches)  evals/s={250000*S/ms/1e6:8.1f} G  wall {wall*1e3:7.1f} ms")
