"""GPU parity tests of the wide-integer (CRT / without-padding) lookup against the CPU oracle.

Integer work (packing-key generation, packing keyswitch) is bit-exact; the fp64 stages (sign bootstraps, CMUX
tree, blind rotation) must make the same decisions as the oracle and decrypt to the exact table value, with
the output phase within a stated torus distance of the oracle's."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from concrete_numpy_b200 import engine as E
from concrete_numpy_b200 import wide
from oracle import wop as W

MODULI = (7, 8, 9, 11)          # 14 extracted bits: 6 CMUX-tree levels over 64 polynomials at N = 256


def toy_params(**kw):
    base = dict(p=12, n=32, k=1, N=256, pbs_level=4, pbs_base_log=8, ks_level=8, ks_base_log=4,
                lwe_std=2.0 ** -40, glwe_std=2.0 ** -50, moduli=MODULI, cbs_level=3, cbs_base_log=6,
                pf_level=6, pf_base_log=7)
    base.update(kw)
    return E.CryptoParams(**base)


def oracle_params(p):
    return W.WopParams(n=p.n, k=p.k, N=p.N, lwe_std=p.lwe_std, glwe_std=p.glwe_std, ks_level=p.ks_level,
                       ks_base_log=p.ks_base_log, pbs_level=p.pbs_level, pbs_base_log=p.pbs_base_log,
                       cbs_level=p.cbs_level, cbs_base_log=p.cbs_base_log, pf_level=p.pf_level,
                       pf_base_log=p.pf_base_log)


@pytest.fixture(scope="module")
def setup(oracle, cuda_device):
    p = toy_params()
    ks = E.KeySet(p, cuda_device, seed=77)
    ev = E.EvalKeys.of(ks)
    keys = W.WopKeys(oracle, oracle_params(p), 77)
    return p, ks, ev, keys


def encrypt_residues(oracle, keys, p, x, seed0=500):
    return [oracle.encrypt(W.encode_residue(x % m, m), keys.sk_big, p.glwe_std, seed0 + i) for i, m in enumerate(p.moduli)]


def torus_dist(a, b):
    d = (a - b).view(np.int64).astype(np.float64)
    return np.abs(d) / 2.0 ** 64


def sign_decision(ct, sk_small, N, theta):
    """What a sign bootstrap decides for small-key ciphertexts: the half of Z_2N the modulus-switched phase falls in
    (a residue sitting exactly on a cell boundary may go either way, but both neighbouring cells decode to that
    residue).  The multi-output bootstrap switches to multiples of 2^theta."""
    lg = N.bit_length() - theta          # log2(2N) - theta
    ms = ((((ct >> np.uint64(64 - lg - 1)) + np.uint64(1)) >> np.uint64(1)) & np.uint64((1 << lg) - 1)) << np.uint64(theta)
    ph = (ms[:, -1].astype(np.int64) - (ms[:, :-1].astype(np.int64) * sk_small.astype(np.int64)[None, :]).sum(axis=1)) % (2 * N)
    return (ph >= N).astype(np.int64)


def test_packing_key_bit_exact(setup):
    p, ks, ev, keys = setup
    assert np.array_equal(E.to_np_u64(ks.pfksk), keys.pfksk)


def test_packing_keyswitch_bit_exact(oracle, setup, cuda_device):
    """Both keyswitch kernels (tensor-core int8 limbs and CUDA cores) on the (k+1)^2 N-column packing key."""
    p, ks, ev, keys = setup
    rng = np.random.default_rng(3)
    ext = rng.integers(0, 1 << 63, (20, p.big_dim + wide.PF_PAD + 1), dtype=np.int64).astype(np.uint64) * np.uint64(2)
    ext[:, p.big_dim + 1:] = 0
    ref = oracle.keyswitch(ext, keys.pfksk, p.pf_level, p.pf_base_log)
    assert ev.packing.tc_ok
    got = ev.packing.apply(E.from_np_u64(ext, cuda_device))
    assert np.array_equal(E.to_np_u64(got), ref)
    ev.packing.tc_ok = False
    try:
        got = ev.packing.apply(E.from_np_u64(ext, cuda_device))
    finally:
        ev.packing.tc_ok = True
    assert np.array_equal(E.to_np_u64(got), ref)


def test_extract_bits_decisions(oracle, setup, cuda_device):
    p, ks, ev, keys = setup
    x = np.arange(0, 5544, 37)
    cts = encrypt_residues(oracle, keys, p, x)
    for ct, m in zip(cts, p.moduli):
        b, off, dec, _ = wide.block_layout(m)
        bits, levels = wide.extract_and_bootstrap(ev, [E.from_np_u64(ct, cuda_device)], [m])
        bits, levels = bits[0], levels[0]
        ref_bits, ref_levels = W.extract_and_bootstrap(oracle, keys, ct, m)
        assert len(bits) == b and len(levels) == b
        theta = wide.bootstrap_theta(p)
        assert theta == W.bootstrap_theta(keys.p)
        cell = np.zeros(len(x), dtype=np.int64)
        for t, (g, r) in enumerate(zip(bits, ref_bits)):
            d, d_ref = sign_decision(E.to_np_u64(g), keys.sk_small, p.N, theta), sign_decision(r, keys.sk_small, p.N, theta)
            assert np.array_equal(d, d_ref)
            cell |= d << t
            # the circuit-bootstrap inputs of this bit: bit * 2^(64-(lv+1)*b) under the big key, from the same rotation
            lv_gpu = E.to_np_u64(levels[t].contiguous())
            for lv in range(p.cbs_level):
                want = d.astype(np.uint64) << np.uint64(64 - (lv + 1) * p.cbs_base_log)
                ph = oracle.decrypt_phase(lv_gpu[:, lv], keys.sk_big)
                ph_ref = oracle.decrypt_phase(ref_levels[t][:, lv], keys.sk_big)
                assert torus_dist(ph, want).max() < 2.0 ** -24 and torus_dist(ph_ref, want).max() < 2.0 ** -24
        assert np.array_equal(np.asarray(dec)[cell], x % m)


def test_cmux_batch_matches_oracle(oracle, setup, cuda_device):
    p, ks, ev, keys = setup
    rng = np.random.default_rng(5)
    G = 6
    bitvals = rng.integers(0, 2, G)
    small = oracle.encrypt((bitvals.astype(np.uint64) << np.uint64(63)) + np.uint64(1 << 62), keys.sk_small, p.lwe_std, 901)
    zero_idx = np.zeros(G, dtype=np.int32)
    ggsw_ref = W.circuit_bootstrap(oracle, keys, W.sign_bootstrap_multi(oracle, keys, small, [1 << 60], zero_idx)[:, 1:])
    ggsw = wide.circuit_bootstrap(ev, wide.sign_bootstrap_multi(ev, E.from_np_u64(small, cuda_device), [1 << 60],
                                                                torch.from_numpy(zero_idx).to(cuda_device))[:, 1:].contiguous())
    glwe = (p.k + 1) * p.N
    pool = np.zeros((8, glwe), dtype=np.uint64)
    pool[:, p.k * p.N:] = rng.integers(0, 1 << 63, (8, p.N), dtype=np.int64).astype(np.uint64) << np.uint64(1)
    i0 = np.array([0, 2, 4, 6, 1, 3], dtype=np.int32)
    i1 = np.array([1, 3, 5, 7, 0, 2], dtype=np.int32)
    gi = np.arange(G, dtype=np.int32)
    ref = oracle.cmux_batch(pool, i0, i1, ggsw_ref, gi, p.cbs_base_log)
    out = torch.empty((G, glwe), dtype=torch.int64, device=cuda_device)
    from concrete_numpy_b200 import _lib
    d_pool, d_i0, d_i1, d_gi = (E.from_np_u64(pool, cuda_device), torch.from_numpy(i0).to(cuda_device),
                                torch.from_numpy(i1).to(cuda_device), torch.from_numpy(gi).to(cuda_device))
    _lib.call("fhe_cmux_batch", out.data_ptr(), d_pool.data_ptr(), d_i0.data_ptr(), d_i1.data_ptr(),
              ggsw.data_ptr(), d_gi.data_ptr(), G, p.k, p.N, p.cbs_level, p.cbs_base_log, 0,
              torch.cuda.current_stream().cuda_stream)
    got = E.to_np_u64(out)
    # the plaintext-leaf shortcut (masks known to be zero) gives the same ciphertexts up to fp64 rounding
    out_plain = torch.empty_like(out)
    _lib.call("fhe_cmux_batch", out_plain.data_ptr(), d_pool.data_ptr(), d_i0.data_ptr(), d_i1.data_ptr(),
              ggsw.data_ptr(), d_gi.data_ptr(), G, p.k, p.N, p.cbs_level, p.cbs_base_log, 1,
              torch.cuda.current_stream().cuda_stream)
    assert torus_dist(E.to_np_u64(out_plain), got).max() < 2.0 ** -40
    # a second level on encrypted inputs (the outputs of the first): still selects the right polynomial
    d_j0 = torch.tensor([0, 2], dtype=torch.int32, device=cuda_device)
    d_j1 = torch.tensor([1, 3], dtype=torch.int32, device=cuda_device)
    d_gj = torch.tensor([4, 5], dtype=torch.int32, device=cuda_device)
    out2 = torch.empty((2, glwe), dtype=torch.int64, device=cuda_device)
    _lib.call("fhe_cmux_batch", out2.data_ptr(), out.data_ptr(), d_j0.data_ptr(), d_j1.data_ptr(),
              ggsw.data_ptr(), d_gj.data_ptr(), 2, p.k, p.N, p.cbs_level, p.cbs_base_log, 0,
              torch.cuda.current_stream().cuda_stream)
    got2 = E.to_np_u64(out2)

    def glwe_phase(ct):      # body - sum_j A_j * S_j (negacyclic), k = 1
        A, Bp = ct[:p.N], ct[p.N:]
        s = keys.sk_big[:p.N]
        acc = Bp.copy()
        for j in np.nonzero(s)[0]:
            acc[j:] -= A[:p.N - j]
            acc[:j] += A[p.N - j:]
        return acc

    first = [int(i1[b] if bitvals[b] else i0[b]) for b in range(G)]
    for q in range(2):
        sel = first[2 * q + 1] if bitvals[4 + q] else first[2 * q]
        # encrypted inputs: the 18-bit rounding of the (uniform) mask is multiplied by the key, std ~ 2^-16.3
        assert torus_dist(glwe_phase(got2[q]), pool[sel, p.k * p.N:]).max() < 2.0 ** -13
    for b in range(G):
        want = pool[i1[b] if bitvals[b] else i0[b], p.k * p.N:]
        ph, ph_ref = glwe_phase(got[b]), glwe_phase(ref[b])
        # 18-bit decomposition of a random polynomial: rounding error up to 2^-19 per coefficient, plus the noise
        assert torus_dist(ph, want).max() < 2.0 ** -17
        assert torus_dist(ph, ph_ref).max() < 2.0 ** -17        # GPU and oracle differ by their (independent) noises


@pytest.mark.parametrize("mapped", [False, True])
def test_wide_lookup_decrypts_to_table(oracle, setup, cuda_device, mapped):
    p, ks, ev, keys = setup
    M = int(np.prod(p.moduli))
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.integers(0, 1 << p.p, 21), [0, (1 << p.p) - 1]])
    tables = np.stack([(np.arange(1 << p.p) * 3 + 1) % 4001, (np.arange(1 << p.p) ** 2) % 3001 - 1500])
    luts = wide.build_crt_luts(tables, p.moduli)
    pool = wide.lut_pool(p, luts, cuda_device)
    polys = max(luts.shape[2] // p.N, 1)
    cts = encrypt_residues(oracle, keys, p, x)
    sel = (np.arange(len(x)) % 2) if mapped else np.zeros(len(x), dtype=np.int64)
    outs = wide.wide_table_lookup(ev, [E.from_np_u64(c, cuda_device) for c in cts], pool, polys,
                                  torch.from_numpy(sel).to(cuda_device) if mapped else None)
    res = []
    for o, m in zip(outs, p.moduli):
        _, ph = E.decrypt(ks, o, return_phase=True)
        res.append(wide.decode_residue(ph, m))
    y = wide.crt_combine(res, p.moduli)
    want = tables[sel, x] % M
    assert np.array_equal(y, want)
    if not mapped:      # the oracle's restatement gives the same answers on the same ciphertexts
        outs_ref = W.wide_lookup(oracle, keys, cts, list(p.moduli), tables[0])
        res_ref = [W.decode_residue(oracle.decrypt_phase(o, keys.sk_big), m) for o, m in zip(outs_ref, p.moduli)]
        assert np.array_equal(W.crt_combine(res_ref, p.moduli), want)


def test_evaluation_keys_roundtrip_with_packing_key(setup, cuda_device):
    """EvaluationKeys.serialize / unserialize (reference tests/compilation/test_circuit.py:174-177) carries the packing
    keyswitch key of the wide path; the restored keys run a lookup."""
    from concrete_numpy_b200 import compiler as cc

    p, ks, ev, keys = setup
    blob = cc.EvaluationKeys(p, ks.bsk_f, ks.ksk, ks.pfksk).serialize()
    back = cc.EvaluationKeys.unserialize(blob, device=cuda_device)
    assert back.crypto == p and back.crypto.moduli == MODULI
    assert torch.equal(back.pfksk, ks.pfksk) and torch.equal(back.ksk, ks.ksk) and torch.equal(back.bsk_f, ks.bsk_f)
    ev2 = back.on(cuda_device)
    x = np.array([0, 5, 4095, 1234])
    table = (np.arange(1 << p.p) * 7 + 1) % 4093
    luts = wide.build_crt_luts(table, p.moduli)
    pool = wide.lut_pool(p, luts, cuda_device)
    cts = [E.encrypt_plaintexts(ks, wide.encode_residue(x, m), first_index=100 * i) for i, m in enumerate(p.moduli)]
    outs = wide.wide_table_lookup(ev2, cts, pool, max(luts.shape[2] // p.N, 1))
    res = [wide.decode_residue(E.decrypt(ks, o, return_phase=True)[1], m) for o, m in zip(outs, p.moduli)]
    assert np.array_equal(wide.crt_combine(res, p.moduli), table[x])
