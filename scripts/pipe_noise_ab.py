"""Development probe: output noise and error count of the N = 32768 bootstrap at the golden cfg2 parameter set,
pipelined kernel vs the plan's generic kernel (run twice: FHE_PBS_PIPE=1 / 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, params as P

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 960
for p in (P.select_parameters(8, 1.0, 1e-8), P.REFERENCE_SETS[8]):
    ks = E.KeySet(p, dev, seed=3)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(1)
    m = rng.integers(0, 1 << p.p, B).astype(np.uint64)
    pt = m << np.uint64(63 - p.p)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    luts = E.from_np_u64(E.build_lut_accumulators(p, np.arange(1 << p.p)[None, :]), dev)
    out = E.bootstrap(ev, small, luts, None)
    msg, ph = E.decrypt(ks, out, return_phase=True)
    err = (ph - pt).view(np.int64).astype(np.float64) / 2.0 ** 64
    bad = np.nonzero(msg != m)[0]
    print(f"PIPE={os.environ.get('FHE_PBS_PIPE', '1')} n={p.n} b={p.pbs_base_log}: out noise std 2^{np.log2(err.std()):.2f} max |err| 2^{np.log2(np.abs(err).max()):.2f} "
          f"errors {len(bad)}/{B} first bad {bad[:10].tolist()} model p_err {P.estimate_p_error(p, 1.0):.1e}", flush=True)
    del ks, ev
    torch.cuda.empty_cache()
