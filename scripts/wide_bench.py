"""Development aid: throughput and output noise of the wide-integer (CRT) lookup at the selected parameters."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, params as P, wide

dev = torch.device("cuda:0")
for p_bits, n_ct in [(int(a.split(":")[0]), int(a.split(":")[1])) for a in (sys.argv[1:] or ["16:256", "12:256", "9:256"])]:
    shapes = tuple(tuple(int(v) for v in t.split("x")) for t in os.environ["WIDE_SHAPES"].split(",")) if os.environ.get("WIDE_SHAPES") else P.WIDE_SHAPES
    prm = P.select_wide_parameters(p_bits, 1.0, 1e-6, shapes)
    t0 = time.time()
    ks = E.KeySet(prm, dev, seed=5)
    ev = E.EvalKeys.of(ks)
    torch.cuda.synchronize()
    t_key = time.time() - t0
    M = int(np.prod(prm.moduli))
    rng = np.random.default_rng(1)
    x = rng.integers(0, 1 << p_bits, n_ct)
    table = rng.integers(0, 1 << p_bits, 1 << p_bits)
    luts = wide.build_crt_luts(table, prm.moduli)
    pool = wide.lut_pool(prm, luts, dev)
    polys = max(luts.shape[2] // prm.N, 1)
    cts = [E.encrypt_plaintexts(ks, wide.encode_residue(x, m), first_index=i * n_ct) for i, m in enumerate(prm.moduli)]
    outs = wide.wide_table_lookup(ev, cts, pool, polys)          # warm-up (key re-layouts, caches)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    outs = wide.wide_table_lookup(ev, cts, pool, polys)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    wide.PROFILE = {}
    wide.wide_table_lookup(ev, cts, pool, polys)
    prof, wide.PROFILE = wide.PROFILE, None
    print("   phases (ms): " + ", ".join(f"{k} {v:.1f}" for k, v in prof.items()))
    res, errs = [], []
    for o, m in zip(outs, prm.moduli):
        _, ph = E.decrypt(ks, o, return_phase=True)
        r = wide.decode_residue(ph, m)
        res.append(r)
        d = (ph - wide.encode_residue(r, m)).view(np.int64).astype(np.float64) / 2.0 ** 64
        errs.append(d)
    y = wide.crt_combine(res, prm.moduli)
    wrong = int(np.sum(y != table[x] % M))
    std = float(np.sqrt(np.mean(np.concatenate(errs) ** 2)))
    print(f"p={p_bits} moduli={prm.moduli} n={prm.n} k={prm.k} N={prm.N} pbs=({prm.pbs_level},{prm.pbs_base_log}) "
          f"cbs=({prm.cbs_level},{prm.cbs_base_log}) pf=({prm.pf_level},{prm.pf_base_log}) ks=({prm.ks_level},{prm.ks_base_log})")
    print(f"   keygen {t_key:.2f} s; {n_ct} lookups in {ms:.1f} ms = {n_ct / ms * 1e3:.1f} lookups/s; wrong {wrong}/{n_ct}; "
          f"output noise std 2^{math.log2(max(std, 1e-30)):.2f} (model 2^{0.5 * math.log2(P.wide_output_variance(prm)):.2f})", flush=True)
    del ks, ev, cts, outs, pool
    torch.cuda.empty_cache()
