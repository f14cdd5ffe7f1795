"""Development aid: empirical lookup error rate at parameters selected for a LARGE p_error (so that errors are
observable), next to the model's estimate.  usage: error_rate_probe.py [bits] [p_error] [lookups]"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, params as P

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
target = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
total = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
dev = torch.device("cuda:0")
p = P.select_parameters(bits, 1.0, target)
model = P.estimate_p_error(p, 1.0)
print("params", p, "model p_error", model, flush=True)
ks = E.KeySet(p, dev, seed=17)
ev = E.EvalKeys.of(ks)
rng = np.random.default_rng(0)
table = rng.integers(0, 1 << bits, (1, 1 << bits))
luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
errs, done, sq = 0, 0, 0.0
B = 1 << 15
while done < total:
    m = rng.integers(0, 1 << bits, B).astype(np.uint64)
    out = E.table_lookup(ev, E.encrypt(ks, m), luts)
    exp = table[0, m.astype(np.int64)].astype(np.uint64)
    msg, ph = E.decrypt(ks, out, return_phase=True)
    errs += int((msg != exp).sum())
    # error of a SECOND lookup on these outputs is what p_error models (input noise = a bootstrap output); measure it too
    out2 = E.table_lookup(ev, out, E.from_np_u64(E.build_lut_accumulators(p, np.arange(1 << bits)[None, :]), dev))
    msg2 = E.decrypt(ks, out2)
    good = msg == exp
    errs2 = int((msg2[good] != exp[good]).sum())
    sq += errs2
    done += B
print(f"lookups {done}: first-level errors {errs} (rate {errs/done:.3g}), second-level errors {int(sq)} (rate {sq/done:.3g}); "
      f"target {target}, model {model:.3g}, binomial 4-sigma bound at target {done*target + 4*math.sqrt(done*target):.0f}")
