"""Development aid: repeat the p=8 small-n bootstrap case and print which ciphertexts decrypt wrong."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E
dev = torch.device("cuda:0")
for (p_bits, n, k, N, level, base_log) in [(8, 12, 1, 32768, 2, 15), (7, 12, 1, 16384, 2, 15), (8, 40, 1, 32768, 2, 15)]:
    p = E.CryptoParams(p=p_bits, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=8,
                       ks_base_log=3, lwe_std=2.0 ** -40, glwe_std=2.0 ** -55)
    ks = E.KeySet(p, dev, seed=5)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(p_bits)
    for B in (20, 20, 7, 64):
        m = rng.integers(0, 1 << p_bits, B).astype(np.uint64)
        tables = rng.integers(0, 1 << p_bits, (3, 1 << p_bits))
        lut_idx = rng.integers(0, 3, B).astype(np.int32)
        luts = E.from_np_u64(E.build_lut_accumulators(p, tables), dev)
        ct = E.encrypt(ks, m)
        small = E.keyswitch(ev, ct)
        exp = tables[lut_idx, m.astype(np.int64)].astype(np.uint64)
        for rep in range(3):
            out = E.bootstrap(ev, small, luts, torch.from_numpy(lut_idx).to(dev))
            got, ph = E.decrypt(ks, out, return_phase=True)
            bad = np.nonzero(got != exp)[0]
            err = (ph - (exp << np.uint64(63 - p_bits))).view(np.int64) / 2.0 ** 64
            print(N, n, "B", B, "rep", rep, "bad", bad.tolist(), "m", m[bad].tolist(), "maxerr(good)",
                  float(np.abs(np.delete(err, bad)).max()) if len(bad) < B else None, "err(bad)", err[bad].tolist(), flush=True)
