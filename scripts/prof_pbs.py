"""One keyswitch + bootstrap launch of a named parameter set (ncu target; not a benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, params as P

name, B = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
P5 = E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                    lwe_std=2.0 ** -20, glwe_std=2.0 ** -50)          # the set the backend selects for cfg5
p = P5 if name == "p5" else P.REFERENCE_SETS[int(name[1:])]
ks = E.KeySet(p, dev, seed=42)
ev = E.EvalKeys.of(ks)
rng = np.random.default_rng(0)
m = rng.integers(0, 1 << p.p, B).astype(np.uint64)
luts = E.from_np_u64(E.build_lut_accumulators(p, np.arange(1 << p.p)[None, :]), dev)
ct = E.encrypt(ks, m)
for _ in range(reps):
    small = E.keyswitch(ev, ct)
    out = E.bootstrap(ev, small, luts, None)
torch.cuda.synchronize()
assert np.array_equal(E.decrypt(ks, out), m)
print("ok", name, B)
