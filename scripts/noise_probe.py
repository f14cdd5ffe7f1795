"""Development probe: measured noise after keyswitch / bootstrap vs the model in params.py, and a
detailed GPU-vs-oracle diff on a small case."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concrete_numpy_b200 import engine as E, params as P
from oracle.oracle import Oracle

dev = torch.device("cuda:0")
o = Oracle()

def diff_case(N, n, level, base_log, p_bits=3, k=1):
    p = E.CryptoParams(p=p_bits, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=4, ks_base_log=4,
                       lwe_std=2.0 ** -30, glwe_std=2.0 ** -50)
    ks = E.KeySet(p, dev, seed=11, keep_std_bsk=True); ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(0)
    m = rng.integers(0, 1 << p_bits, 32).astype(np.uint64)
    table = rng.integers(0, 1 << p_bits, (1, 1 << p_bits))
    ct = E.encrypt(ks, m); small = E.keyswitch(ev, ct)
    luts = E.build_lut_accumulators(p, table)
    out = E.to_np_u64(E.bootstrap(ev, small, E.from_np_u64(luts, dev), None))
    ref = o.pbs(E.to_np_u64(small), o.bsk_to_fourier(E.to_np_u64(ks.bsk_std)), luts, None, base_log)
    d = (out - ref).view(np.int64)
    bad = np.argwhere(np.abs(d.astype(np.float64)) > 2.0 ** 40)
    print(f"diff_case N={N} n={n} l={level} b={base_log}: max|d|=2^{np.log2(np.abs(d.astype(np.float64)).max()+1):.1f} nbad={len(bad)}")
    for e, j in bad[:8]:
        print("   ct", e, "word", j, "gpu", hex(out[e, j]), "ref", hex(ref[e, j]))

def noise_case(name, p, B=256, oracle_n=0):
    t0 = time.time()
    ks = E.KeySet(p, dev, seed=3, keep_std_bsk=oracle_n > 0); ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(1)
    m = rng.integers(0, 1 << p.p, B).astype(np.uint64)
    pt = m << np.uint64(63 - p.p)
    ct = E.encrypt(ks, m)
    _, ph = E.decrypt(ks, ct, return_phase=True)
    f = lambda e: np.log2(np.std((e).view(np.int64).astype(np.float64)) / 2.0 ** 64 + 1e-300)
    small = E.keyswitch(ev, ct)
    ph_s = E.decrypt_small(ks, small)
    table = np.arange(1 << p.p)[None, :]
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), dev)
    out = E.bootstrap(ev, small, luts, None)
    msg, ph_o = E.decrypt(ks, out, return_phase=True)
    model_ks = 0.5 * np.log2(p.glwe_std ** 2 + P.variance_keyswitch(p.big_dim, p.ks_level, p.ks_base_log, p.lwe_std ** 2))
    model_out = 0.5 * np.log2(P.variance_pbs_out(p.n, p.k, p.N, p.pbs_level, p.pbs_base_log, p.glwe_std ** 2))
    print(f"{name}: fresh 2^{f(ph - pt):.2f} | ks meas 2^{f(ph_s - pt):.2f} model 2^{model_ks:.2f} | pbs out meas 2^{f(ph_o - pt):.2f} "
          f"model 2^{model_out:.2f} | errors {(msg != m).sum()}/{B} | model p_err {P.estimate_p_error(p, 1.0):.2e} | {time.time()-t0:.1f}s", flush=True)
    if oracle_n:
        sm = E.to_np_u64(small)[:oracle_n]
        ref = o.pbs(sm, o.bsk_to_fourier(E.to_np_u64(ks.bsk_std)), E.to_np_u64(luts), None, p.pbs_base_log)
        sk_b = E.to_np_u64(ks.sk_big)
        ph_r = o.decrypt_phase(ref, sk_b)
        print(f"    oracle pbs out 2^{f(ph_r - pt[:oracle_n]):.2f} errors {(o.decode(ph_r, p.p) != m[:oracle_n]).sum()}/{oracle_n}; gpu-vs-oracle max 2^{np.log2(np.abs((E.to_np_u64(out)[:oracle_n]-ref).view(np.int64).astype(np.float64)).max()+1):.1f}")
    del ks, ev; torch.cuda.empty_cache()

diff_case(1024, 64, 2, 12)
diff_case(1024, 16, 1, 20)
diff_case(2048, 16, 2, 12)
for pe in (1e-6,):
    noise_case("sel p8 1e-6", P.select_parameters(8, 1.0, pe), oracle_n=4)
noise_case("ref p8", P.REFERENCE_SETS[8])
noise_case("ref p4", P.REFERENCE_SETS[4], B=1024, oracle_n=64)
noise_case("ref p6", P.REFERENCE_SETS[6])
for p_bits in (1, 2, 3, 4, 5, 6, 7):
    noise_case(f"sel p{p_bits} 1e-6", P.select_parameters(p_bits, 1.0, 1e-6), B=512)
q = P.select_parameters(8, 1.0, 1e-6)
for b in (15, 20, 23):
    lv = 2 if b == 15 else 1
    noise_case(f"p8 N32768 l={lv} b={b}", E.CryptoParams(p=8, n=q.n, k=1, N=32768, pbs_level=lv, pbs_base_log=b, ks_level=q.ks_level, ks_base_log=q.ks_base_log, lwe_std=q.lwe_std, glwe_std=q.glwe_std), B=128)
