"""GPU parity tests of the kernels behind the C ABI against the CPU oracle (same seeded inputs).

Integer / byte work (keygen, encryption, keyswitch, linear ops) must be BIT-EXACT.
The programmable bootstrap runs an fp64 FFT: its outputs must decrypt to the exact table value and
stay within a stated torus distance of the oracle's output on the same inputs (tolerance below).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from concrete_numpy_b200 import engine as E
from concrete_numpy_b200 import _lib


def variance_ratio_slack(samples: int) -> float:
    """3-sigma sampling error of the ratio of two variance estimates from `samples` Gaussian samples each."""
    return 1.0 + 3.0 * (4.0 / max(samples - 1, 1)) ** 0.5


def small_params(**kw):
    base = dict(p=3, n=48, k=1, N=512, pbs_level=2, pbs_base_log=10, ks_level=4, ks_base_log=4,
                lwe_std=2.0 ** -30, glwe_std=2.0 ** -45)
    base.update(kw)
    return E.CryptoParams(**base)


def oracle_keys(oracle, p, seed):
    sk_s = oracle.keygen_lwe_secret(p.n, seed, 1)
    sk_b = oracle.keygen_lwe_secret(p.k * p.N, seed, 2)
    return sk_s, sk_b


@pytest.mark.parametrize("k,N", [(1, 512), (2, 256), (1, 1024)])
def test_keygen_bit_exact(oracle, cuda_device, k, N):
    p = small_params(k=k, N=N, n=40)
    ks = E.KeySet(p, cuda_device, seed=1234, keep_std_bsk=True)
    sk_s, sk_b = oracle_keys(oracle, p, 1234)
    assert np.array_equal(E.to_np_u64(ks.sk_small), sk_s)
    assert np.array_equal(E.to_np_u64(ks.sk_big), sk_b)
    ksk = oracle.keygen_ksk(sk_b, sk_s, p.ks_level, p.ks_base_log, p.lwe_std, 1234)
    assert np.array_equal(E.to_np_u64(ks.ksk), ksk)
    bsk = oracle.keygen_bsk(sk_s, sk_b, p.k, p.N, p.pbs_level, p.pbs_base_log, p.glwe_std, 1234)
    assert np.array_equal(E.to_np_u64(ks.bsk_std), bsk)


def test_encrypt_decrypt_bit_exact(oracle, cuda_device):
    p = small_params()
    ks = E.KeySet(p, cuda_device, seed=7)
    _, sk_b = oracle_keys(oracle, p, 7)
    rng = np.random.default_rng(0)
    m = rng.integers(0, 1 << p.p, 300).astype(np.uint64)
    ct = E.encrypt(ks, m, first_index=11, seed=99)
    ref = oracle.encrypt(m << np.uint64(63 - p.p), sk_b, p.glwe_std, 99, first_index=11)
    assert np.array_equal(E.to_np_u64(ct), ref)
    msg, phase = E.decrypt(ks, ct, return_phase=True)
    assert np.array_equal(phase, oracle.decrypt_phase(ref, sk_b))
    assert np.array_equal(msg, m)
    # empty batch
    assert E.encrypt(ks, np.zeros(0, dtype=np.uint64)).shape[0] == 0
    # trivial ciphertexts
    tv = E.trivial(p, m[:5], cuda_device)
    assert np.array_equal(E.to_np_u64(tv), oracle.trivial(m[:5] << np.uint64(63 - p.p), p.big_dim))


def test_key_material_entropy_and_fresh_encryption_randomness(oracle, cuda_device):
    """Production key sets take 64 bytes of OS entropy (bit-exact against the oracle given the same bytes), two key
    sets never share keys, the public masks are not a stream of the secret key, and two encryptions of the same
    message under one key set never share mask or noise (ciphertext differences must not reveal plaintext differences)."""
    p = small_params()
    raw = os.urandom(64)
    ks = E.KeySet(p, cuda_device, seed=raw, keep_std_bsk=True)
    assert not ks.deterministic
    sk_s = oracle.keygen_lwe_secret(p.n, raw, 1)
    sk_b = oracle.keygen_lwe_secret(p.big_dim, raw, 2)
    assert np.array_equal(E.to_np_u64(ks.sk_small), sk_s) and np.array_equal(E.to_np_u64(ks.sk_big), sk_b)
    assert np.array_equal(E.to_np_u64(ks.ksk), oracle.keygen_ksk(sk_b, sk_s, p.ks_level, p.ks_base_log, p.lwe_std, raw))
    # same secret half, different public half: same secret keys, different masks
    other = E.KeySet(p, cuda_device, seed=raw[:32] + os.urandom(32))
    assert np.array_equal(E.to_np_u64(other.sk_big), sk_b)
    assert not np.array_equal(E.to_np_u64(other.ksk)[:, :, :-1], E.to_np_u64(ks.ksk)[:, :, :-1])
    a, b = E.KeySet(p, cuda_device), E.KeySet(p, cuda_device)       # default: fresh entropy
    assert not np.array_equal(E.to_np_u64(a.sk_big), E.to_np_u64(b.sk_big))
    m = np.arange(16, dtype=np.uint64) % (1 << p.p)
    c1, c2 = E.to_np_u64(E.encrypt(ks, m)), E.to_np_u64(E.encrypt(ks, m))
    assert not np.array_equal(c1[:, :-1], c2[:, :-1])                # fresh masks on every call
    assert len(set((c1[:, 0] - c2[:, 0]).tolist())) == 16
    assert np.array_equal(E.decrypt(ks, E.from_np_u64(c1, cuda_device)), m)
    assert np.array_equal(E.decrypt(ks, E.from_np_u64(c2, cuda_device)), m)


@pytest.mark.parametrize("B", [1, 5, 33, 600])
@pytest.mark.parametrize("kN,n,level,base_log", [(512, 48, 4, 4), (1024, 97, 5, 3), (600, 130, 2, 9)])
def test_keyswitch_bit_exact(oracle, cuda_device, B, kN, n, level, base_log):
    rng = np.random.default_rng(B * 131 + kN)
    ksk = rng.integers(0, 1 << 64, (kN, level, n + 1), dtype=np.uint64)
    ct = rng.integers(0, 1 << 64, (B, kN + 1), dtype=np.uint64)
    out = torch.empty((B, n + 1), dtype=torch.int64, device=cuda_device)
    d_ksk, d_ct = E.from_np_u64(ksk, cuda_device), E.from_np_u64(ct, cuda_device)
    _lib.call("fhe_keyswitch_batch", out.data_ptr(), d_ct.data_ptr(), B, d_ksk.data_ptr(), kN, n, level,
              base_log, torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(E.to_np_u64(out), oracle.keyswitch(ct, ksk, level, base_log))


def _negacyclic_sparse(a_pos, a_val, b):
    N = b.size
    out = np.zeros(N, dtype=np.uint64)
    for pos, val in zip(a_pos, a_val):
        rolled = np.concatenate([np.uint64(0) - b[N - pos:], b[:N - pos]]) if pos else b.copy()
        out += rolled * np.array(val, dtype=np.int64).view(np.uint64)
    return out


@pytest.mark.parametrize("N,k,level", [(256, 1, 1), (512, 2, 2), (2048, 1, 1), (4096, 1, 2), (8192, 1, 2),
                                       (16384, 1, 2), (32768, 1, 2)])
def test_fft_polymul(cuda_device, N, k, level):
    """fp64 negacyclic FFT product == exact integer product (sparse small multiplier), to fp64
    rounding: |err| <= 2^(64-53) * |a|_1 * 2^6."""
    lib = _lib.load()
    C = lib.fhe_pbs_cluster_size(N, k, level)
    assert C >= 1
    rng = np.random.default_rng(N)
    npoly = 3
    a = np.zeros((npoly, N), dtype=np.int64)
    b = rng.integers(0, 1 << 64, (npoly, N), dtype=np.uint64)
    exp = np.zeros((npoly, N), dtype=np.uint64)
    for q in range(npoly):
        pos = rng.choice(N, 6, replace=False)
        pos[0] = 0 if q == 0 else pos[0]
        val = rng.integers(-512, 512, 6)
        a[q, pos] = val
        exp[q] = _negacyclic_sparse(pos, val, b[q])
    d_a = torch.from_numpy(a).to(cuda_device)
    d_b = E.from_np_u64(b, cuda_device)
    out = torch.empty((npoly, N), dtype=torch.int64, device=cuda_device)
    _lib.call("fhe_debug_polymul", out.data_ptr(), d_a.data_ptr(), d_b.data_ptr(), npoly, N, k, level,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = (E.to_np_u64(out) - exp).view(np.int64)
    bound = 2.0 ** 11 * 6 * 512 * 64
    assert np.abs(err).max() <= bound, (C, np.abs(err).max(), bound)


PBS_CASES = [
    # p, n, k, N, level, base_log   (small n keeps the oracle fast; noise is negligible)
    (2, 24, 1, 512, 2, 10),
    (2, 16, 2, 256, 1, 20),
    (3, 24, 3, 256, 2, 12),
    (4, 40, 1, 2048, 1, 23),
    (5, 24, 1, 4096, 2, 15),
    (6, 20, 1, 8192, 2, 15),
    (7, 12, 1, 16384, 2, 15),
    (8, 12, 1, 32768, 2, 15),
    (4, 16, 1, 8192, 3, 10),
    (4, 8, 1, 32768, 3, 10),      # 16-CTA cluster plan (two clusters per SM)
    (3, 12, 1, 4096, 3, 10),      # 2-CTA cluster plan with 1024-point local transforms
    (4, 16, 1, 2048, 3, 12),      # 36 kept bits: 64-bit decomposition state
    (3, 12, 1, 1024, 4, 9),       # 4 levels
    (3, 16, 2, 1024, 1, 18),      # k = 2 (radix-4 fused pass)
    (3, 16, 3, 512, 1, 18),       # k = 3
]


@pytest.mark.parametrize("B,kN,n,level,base_log", [(1, 256, 31, 2, 4), (37, 512, 48, 4, 4), (130, 1024, 97, 5, 3),
                                                   (300, 2048, 130, 2, 7), (129, 4096, 500, 8, 2)])
def test_keyswitch_tensor_core_bit_exact(oracle, cuda_device, B, kN, n, level, base_log):
    """The tcgen05 int8-limb GEMM keyswitch (csrc/keyswitch_tc.cu) equals the oracle bit for bit."""
    lib = _lib.load()
    assert lib.fhe_ks_tc_supported(kN, n, level, base_log) == 1
    assert lib.fhe_ks_tc_supported(kN, n, level, 9) == 0            # digits no longer fit int8
    rng = np.random.default_rng(B * 7 + kN)
    ksk = rng.integers(0, 1 << 64, (kN, level, n + 1), dtype=np.uint64)
    ct = rng.integers(0, 1 << 64, (B, kN + 1), dtype=np.uint64)
    d_ksk, d_ct = E.from_np_u64(ksk, cuda_device), E.from_np_u64(ct, cuda_device)
    st = torch.cuda.current_stream().cuda_stream
    ksk_tc = torch.empty(lib.fhe_ksk_tc_bytes(kN, n, level), dtype=torch.uint8, device=cuda_device)
    _lib.call("fhe_ksk_to_tc", ksk_tc.data_ptr(), d_ksk.data_ptr(), kN, n, level, st)
    scratch = torch.empty(lib.fhe_keyswitch_tc_scratch_bytes(B, kN, level), dtype=torch.uint8, device=cuda_device)
    out = torch.empty((B, n + 1), dtype=torch.int64, device=cuda_device)
    _lib.call("fhe_keyswitch_batch_tc", out.data_ptr(), d_ct.data_ptr(), B, ksk_tc.data_ptr(), scratch.data_ptr(),
              kN, n, level, base_log, st)
    assert np.array_equal(E.to_np_u64(out), oracle.keyswitch(ct, ksk, level, base_log))


@pytest.mark.parametrize("p_bits,n,k,N,level,base_log", PBS_CASES)
def test_pbs_matches_oracle(oracle, cuda_device, p_bits, n, k, N, level, base_log):
    p = E.CryptoParams(p=p_bits, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=8,
                       ks_base_log=3, lwe_std=2.0 ** -40, glwe_std=2.0 ** -55)
    ks = E.KeySet(p, cuda_device, seed=5, keep_std_bsk=True)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(p_bits)
    B = 32
    m = rng.integers(0, 1 << p_bits, B).astype(np.uint64)
    tables = rng.integers(0, 1 << p_bits, (3, 1 << p_bits))
    lut_idx = rng.integers(0, 3, B).astype(np.int32)
    luts = E.build_lut_accumulators(p, tables)
    for t in range(3):  # host LUT expansion == oracle's
        assert np.array_equal(luts[t, k * N:], oracle.build_lut_poly(tables[t], p_bits, p_bits, N))
    ct = E.encrypt(ks, m)
    small = E.keyswitch(ev, ct)
    d_luts = E.from_np_u64(luts, cuda_device)
    d_idx = torch.from_numpy(lut_idx).to(cuda_device)
    out = E.bootstrap(ev, small, d_luts, d_idx)
    got = E.decrypt(ks, out)
    exp = tables[lut_idx, m.astype(np.int64)].astype(np.uint64)
    assert np.array_equal(got, exp)
    # Same inputs through the oracle's bootstrap.  Raw ciphertext words are NOT comparable between
    # two fp64 FFT pipelines: a 1-ulp difference flips a decomposition rounding somewhere and injects a
    # different (equally valid) BSK row, so the comparison is made on the decrypted phase: identical
    # messages, and a phase error of the same size (tolerance: north_star, VARIANCE <= 2x the oracle's, times the
    # 3-sigma sampling error of a ratio of two B-sample variance estimates, sqrt(4 / (B - 1)) relative).
    bsk_f = oracle.bsk_to_fourier(E.to_np_u64(ks.bsk_std))
    ref = oracle.pbs(E.to_np_u64(small), bsk_f, luts, lut_idx, base_log)
    sk_b = E.to_np_u64(ks.sk_big)
    pt = exp << np.uint64(63 - p_bits)
    ph_ref = oracle.decrypt_phase(ref, sk_b)
    _, ph_gpu = E.decrypt(ks, out, return_phase=True)
    assert np.array_equal(oracle.decode(ph_ref, p_bits), exp)
    err_ref = (ph_ref - pt).view(np.int64).astype(np.float64)
    err_gpu = (ph_gpu - pt).view(np.int64).astype(np.float64)
    assert err_gpu.var() <= 2.0 * variance_ratio_slack(B) * err_ref.var() + 2.0 ** 40, (err_gpu.var(), err_ref.var())


@pytest.mark.parametrize("p_bits,n,k,N,level,base_log", [(3, 16, 1, 1024, 2, 10), (4, 20, 1, 2048, 1, 23),
                                                       (5, 16, 1, 4096, 2, 13), (6, 12, 1, 8192, 2, 13),
                                                       (7, 8, 1, 16384, 2, 15), (8, 8, 1, 32768, 2, 15)])
def test_pbs_plan_variants_agree(cuda_device, p_bits, n, k, N, level, base_log, monkeypatch):
    """Every plan variant (cluster size) of a parameter set bootstraps to the same messages from the
    Fourier BSK permuted into its layout (fhe_bsk_relayout), with the same noise level."""
    p = E.CryptoParams(p=p_bits, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=8,
                       ks_base_log=3, lwe_std=2.0 ** -40, glwe_std=2.0 ** -55)
    ks = E.KeySet(p, cuda_device, seed=9)
    ev = E.EvalKeys.of(ks)
    assert len(ev.variant_clusters) >= 2 and ev.variant_clusters[0] == _lib.load().fhe_pbs_cluster_size(N, k, level)
    rng = np.random.default_rng(N + 1)
    B = 64
    m = rng.integers(0, 1 << p_bits, B).astype(np.uint64)
    table = rng.integers(0, 1 << p_bits, (1, 1 << p_bits))
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), cuda_device)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    exp = table[0, m.astype(np.int64)].astype(np.uint64)
    stds = []
    for v in range(len(ev.variant_clusters)):
        out = E.bootstrap(ev, small, luts, None, variant=v)
        got, ph = E.decrypt(ks, out, return_phase=True)
        assert np.array_equal(got, exp), (v, ev.variant_clusters[v])
        stds.append((ph - (exp << np.uint64(63 - p_bits))).view(np.int64).astype(np.float64).std())
    # The variants factor the FFT differently (more, smaller passes on larger clusters), so the fp64 rounding term of
    # the output noise differs between them; each must stay within the tolerance (variance <= 2x) of the noise model
    # that the parameter search relies on (params.variance_pbs_out, calibrated on the published fp64-FFT law)
    from concrete_numpy_b200 import params as P

    model_std = (P.variance_pbs_out(n, k, N, level, base_log, p.glwe_std ** 2) ** 0.5) * 2.0 ** 64
    assert max(stds) <= (2.0 ** 0.5) * model_std + 2.0 ** 20, (stds, model_std)
    assert max(stds) <= 4.0 * min(stds) + 2.0 ** 20, stds
    best = max(v for v, c in enumerate(ev.variant_clusters) if v == 0 or c <= 8)
    # a pipelined variant with at least half the CTAs of the largest cluster is the latency plan (N = 4096: 4 vs 8)
    lib = _lib.load()
    piped = [v for v, c in enumerate(ev.variant_clusters)
             if v > 0 and level == 2 and lib.fhe_pbs_variant_pipelined(N, k, level, v) == 1
             and 2 * c >= ev.variant_clusters[best] > c]
    # ... and the grouped plan (one CTA per SM, k_pbs_group) is the latency plan of its size (N = 2048, up to two levels)
    grouped0 = lib.fhe_pbs_variant_grouped(N, k, level, 0) == 1
    assert grouped0 == (N == 2048 and k == 1 and level <= 2)
    monkeypatch.setenv("CNP_B200_VARIANT_CALIB", "0")           # the static rule (the default measures, see the next test)
    assert ev.variant_for_batch(1) == (0 if grouped0 else piped[0] if piped else best) and ev.variant_for_batch(100000) == 0
    monkeypatch.delenv("CNP_B200_VARIANT_CALIB")
    assert 0 <= ev.variant_for_batch(3) < len(ev.variant_clusters) and ev.variant_for_batch(100000) == 0
    if N == 4096:
        assert piped and ev.variant_clusters[piped[0]] == 4
    assert all(c >= 1 for c in ev._capacity)


def test_variant_selection_is_measured(cuda_device):
    """Single-wave batches run the plan variant that was FASTEST when that batch size was first seen (engine.EvalKeys.
    _calibrate): at N = 4096 a lone ciphertext belongs on the pipelined 4-CTA plan, and the choice for 60 ciphertexts must
    be at least as fast as the throughput plan measured here on the same batch."""
    p = E.CryptoParams(p=5, n=786, k=1, N=4096, pbs_level=2, pbs_base_log=13, ks_level=5, ks_base_log=3,
                       lwe_std=2.0 ** -20, glwe_std=2.0 ** -50)
    ks = E.KeySet(p, cuda_device, seed=23)
    ev = E.EvalKeys.of(ks)
    table = np.arange(32)[None, :]
    luts = E.from_np_u64(E.build_lut_accumulators(p, table), cuda_device)
    m = (np.arange(60) % 32).astype(np.uint64)
    small = E.keyswitch(ev, E.encrypt(ks, m))

    def timed(v):
        E.bootstrap(ev, small, luts, None, variant=v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = E.bootstrap(ev, small, luts, None, variant=v)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out

    assert ev.variant_clusters[ev.variant_for_batch(1)] == 4
    v60 = ev.variant_for_batch(60)
    t_best, out = timed(v60)
    t_0, _ = timed(0)
    assert t_best <= 1.05 * t_0, (v60, t_best, t_0)
    assert np.array_equal(E.decrypt(ks, out), m)
    assert ev.variant_for_batch(60) == v60                       # cached


@pytest.mark.parametrize("p_bits,N,base_log,cluster", [(8, 32768, 15, 8), (7, 16384, 15, 4), (6, 8192, 13, 2),
                                                       (5, 4096, 13, 4)])     # the last one is a latency variant
def test_pipelined_bootstrap_is_bit_identical_to_the_generic_kernel(cuda_device, p_bits, N, base_log, cluster):
    """k_pbs_pipe (pbs_pipe.cuh) replaces the cluster barriers of k_pbs by per-buffer mbarriers and split handshakes; it
    performs the same floating-point operations in the same order, so every output WORD must equal the generic kernel's
    (fhe_pbs_set_pipeline switches between them at run time).  A lost or early exchange -- the class of error a missing
    handshake causes -- corrupts an accumulator and shows up here as a mismatch.  CNP_B200_STRESS=1 runs the long form
    (the race fixed in round 2 appeared about once per 10^7 CMUX iterations)."""
    stress = os.environ.get("CNP_B200_STRESS", "0") not in ("", "0")
    n, B = (996, 3840) if stress else (160, 480)
    p = E.CryptoParams(p=p_bits, n=n, k=1, N=N, pbs_level=2, pbs_base_log=base_log, ks_level=8, ks_base_log=3,
                       lwe_std=2.0 ** -40, glwe_std=2.0 ** -55)
    ks = E.KeySet(p, cuda_device, seed=21)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(N)
    m = rng.integers(0, 1 << p_bits, B).astype(np.uint64)
    tables = rng.integers(0, 1 << p_bits, (2, 1 << p_bits))
    lut_idx = torch.from_numpy(rng.integers(0, 2, B).astype(np.int32)).to(cuda_device)
    luts = E.from_np_u64(E.build_lut_accumulators(p, tables), cuda_device)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    lib = _lib.load()
    v = ev.variant_clusters.index(cluster)
    meta = (C.c_int * 8)()
    assert lib.fhe_pbs_variant_plan(N, 1, 2, v, meta) == 0 and meta[6] == 1      # the pipelined plans run one CTA per SM
    prev = lib.fhe_pbs_set_pipeline(1)
    try:
        piped = [E.to_np_u64(E.bootstrap(ev, small, luts, lut_idx, variant=v)) for _ in range(2)]
        lib.fhe_pbs_set_pipeline(0)
        generic = E.to_np_u64(E.bootstrap(ev, small, luts, lut_idx, variant=v))
    finally:
        lib.fhe_pbs_set_pipeline(prev)
    exp = tables[lut_idx.cpu().numpy(), m.astype(np.int64)].astype(np.uint64)
    assert np.array_equal(E.decrypt(ks, E.from_np_u64(generic, cuda_device)), exp)
    for out in piped:
        bad = np.nonzero((out != generic).any(axis=1))[0]
        assert bad.size == 0, (bad[:8].tolist(), int(bad.size))


@pytest.mark.parametrize("B,level,base_log", [(1, 1, 23), (100, 1, 23), (148, 1, 23), (149, 1, 23), (296, 1, 23),
                                              (300, 1, 23), (1000, 1, 23), (300, 2, 12), (449, 3, 8), (200, 4, 6)])
def test_grouped_bootstrap_is_bit_identical_to_the_generic_kernel(cuda_device, B, level, base_log):
    """k_pbs_group (pbs_group.cuh: one CTA per SM carries two N = 2048 ciphertexts; mode 1 streams GGSW_i from L2, mode 2
    stages it once per SM by bulk copies into a shared-memory ring read by both) performs the floating-point operations of
    k_pbs in the same order: every output WORD must be equal (fhe_pbs_set_group switches at run time).  The batch sizes
    cover one slot per CTA, ragged second slots, and several rounds with a ragged last one (the ring runs on across
    rounds).  With two levels the GGSW does not fit next to two ciphertexts (mode 2 streams as well); with three levels only
    one ciphertext fits an SM: that plan runs the grouped kernel with a single 256-thread slot."""
    stress = os.environ.get("CNP_B200_STRESS", "0") not in ("", "0")
    n = 742 if stress else 96
    p = E.CryptoParams(p=4, n=n, k=1, N=2048, pbs_level=level, pbs_base_log=base_log, ks_level=5, ks_base_log=3,
                       lwe_std=2.0 ** -40, glwe_std=2.0 ** -52)
    ks = E.KeySet(p, cuda_device, seed=22)
    ev = E.EvalKeys.of(ks)
    rng = np.random.default_rng(B)
    m = rng.integers(0, 16, B).astype(np.uint64)
    tables = rng.integers(0, 16, (2, 16))
    lut_idx = torch.from_numpy(rng.integers(0, 2, B).astype(np.int32)).to(cuda_device)
    luts = E.from_np_u64(E.build_lut_accumulators(p, tables), cuda_device)
    small = E.keyswitch(ev, E.encrypt(ks, m))
    lib = _lib.load()
    prev = lib.fhe_pbs_set_group(1)
    try:
        grouped = [E.to_np_u64(E.bootstrap(ev, small, luts, lut_idx)) for _ in range(2)]
        lib.fhe_pbs_set_group(2)
        grouped.append(E.to_np_u64(E.bootstrap(ev, small, luts, lut_idx)))
        lib.fhe_pbs_set_group(0)
        generic = E.to_np_u64(E.bootstrap(ev, small, luts, lut_idx))
    finally:
        lib.fhe_pbs_set_group(prev)
    exp = tables[lut_idx.cpu().numpy(), m.astype(np.int64)].astype(np.uint64)
    assert np.array_equal(E.decrypt(ks, E.from_np_u64(generic, cuda_device)), exp)
    for out in grouped:
        bad = np.nonzero((out != generic).any(axis=1))[0]
        assert bad.size == 0, (bad[:8].tolist(), int(bad.size))


def test_linear_ops_bit_exact(oracle, cuda_device):
    rng = np.random.default_rng(3)
    L = 2049
    a = rng.integers(0, 1 << 64, (37, L), dtype=np.uint64)
    b = rng.integers(0, 1 << 64, (37, L), dtype=np.uint64)
    da, db = E.from_np_u64(a, cuda_device), E.from_np_u64(b, cuda_device)
    assert np.array_equal(E.to_np_u64(E.lin_add(da, db, L)), oracle.lin_add(a, b))
    assert np.array_equal(E.to_np_u64(E.lin_sub(da, db, L)), oracle.lin_sub(a, b))
    assert np.array_equal(E.to_np_u64(E.lin_neg(da, L)), oracle.lin_neg(a))
    # broadcast through row maps
    ia = torch.from_numpy(rng.integers(0, 37, 50).astype(np.int32)).to(cuda_device)
    ib = torch.from_numpy(rng.integers(0, 37, 50).astype(np.int32)).to(cuda_device)
    got = E.to_np_u64(E.lin_add(da, db, L, ia, ib, E=50))
    assert np.array_equal(got, a[ia.cpu().numpy()] + b[ib.cpu().numpy()])
    got = E.to_np_u64(E.lin_sub(da, db, L, ia, ib, E=50))
    assert np.array_equal(got, a[ia.cpu().numpy()] - b[ib.cpu().numpy()])
    pt = rng.integers(0, 1 << 64, 37, dtype=np.uint64)
    dpt = E.from_np_u64(pt, cuda_device)
    assert np.array_equal(E.to_np_u64(E.lin_add_plain(da, dpt, L)), oracle.lin_add_plain(a, pt))
    exp = oracle.lin_neg(a)
    exp[:, -1] += pt
    assert np.array_equal(E.to_np_u64(E.lin_plain_sub(da, dpt, L)), exp)
    c = rng.integers(-100, 100, 37).astype(np.int64)
    dc = torch.from_numpy(c).to(cuda_device)
    assert np.array_equal(E.to_np_u64(E.lin_mul_clear(da, dc, L)), oracle.lin_mul_clear(a, c))
    # gather-mac (general linear map) incl. "no term" entries and bias
    idx = rng.integers(-1, 37, (21, 9)).astype(np.int32)
    w = rng.integers(-3, 4, (21, 9)).astype(np.int64)
    bias = rng.integers(0, 1 << 64, 21, dtype=np.uint64)
    got = E.lin_gather_mac(da, torch.from_numpy(idx).to(cuda_device), torch.from_numpy(w).to(cuda_device), L,
                           E.from_np_u64(bias, cuda_device))
    assert np.array_equal(E.to_np_u64(got), oracle.lin_gather_mac(a, idx, w, bias))
    # dense matmul kernel == gather formulation
    Mr, K, Nc = 5, 7, 19
    x = rng.integers(0, 1 << 64, (Mr * K, L), dtype=np.uint64)
    W = rng.integers(-4, 5, (K, Nc)).astype(np.int64)
    gi = np.zeros((Mr * Nc, K), dtype=np.int32)
    gw = np.zeros((Mr * Nc, K), dtype=np.int64)
    for mm in range(Mr):
        for nn in range(Nc):
            gi[mm * Nc + nn] = mm * K + np.arange(K)
            gw[mm * Nc + nn] = W[:, nn]
    got = E.matmul_ct_clear(E.from_np_u64(x, cuda_device), torch.from_numpy(W).to(cuda_device), Mr, K, Nc, L)
    assert np.array_equal(E.to_np_u64(got), oracle.lin_gather_mac(x, gi, gw))


@pytest.mark.parametrize("L", [1, 2, 513, 2049])
def test_linear_kernels_all_paths_bit_exact(oracle, cuda_device, L):
    """Every code path of the round-2 linear kernels: rows on both 16-byte phases (vector path with scalar head / tail,
    scalar path when operand rows differ in phase), the staged matmul (x read once, sparse weights compacted) and its
    dense fallback for weights beyond 32 bits, gather-MAC with more taps than one shared-memory round holds."""
    rng = np.random.default_rng(L)
    Erows = 11
    a = rng.integers(0, 1 << 64, (Erows, L), dtype=np.uint64)
    b = rng.integers(0, 1 << 64, (Erows, L), dtype=np.uint64)
    da, db = E.from_np_u64(a, cuda_device), E.from_np_u64(b, cuda_device)
    for ia_np, ib_np in [(np.arange(Erows), np.arange(Erows)),              # same phase everywhere: vector path
                         (np.arange(Erows)[::-1].copy(), np.arange(Erows)),  # reversed rows: phases agree only for odd counts
                         (rng.integers(0, Erows, Erows), rng.integers(0, Erows, Erows))]:
        ia = torch.from_numpy(ia_np.astype(np.int32)).to(cuda_device)
        ib = torch.from_numpy(ib_np.astype(np.int32)).to(cuda_device)
        assert np.array_equal(E.to_np_u64(E.lin_add(da, db, L, ia, ib, E=Erows)), a[ia_np] + b[ib_np])
        assert np.array_equal(E.to_np_u64(E.lin_sub(da, db, L, ia, ib, E=Erows)), a[ia_np] - b[ib_np])
        pt = rng.integers(0, 1 << 64, Erows, dtype=np.uint64)
        exp = a[ia_np].copy()
        exp[:, -1] += pt
        assert np.array_equal(E.to_np_u64(E.lin_add_plain(da, E.from_np_u64(pt, cuda_device), L, ia)), exp)
        exp = np.uint64(0) - a[ia_np]
        exp[:, -1] += pt
        assert np.array_equal(E.to_np_u64(E.lin_plain_sub(da, E.from_np_u64(pt, cuda_device), L, ia)), exp)
        c = rng.integers(-(1 << 40), 1 << 40, Erows).astype(np.int64)
        assert np.array_equal(E.to_np_u64(E.lin_mul_clear(da, torch.from_numpy(c).to(cuda_device), L, ia)),
                              a[ia_np] * c.view(np.uint64)[:, None])
    # matmul: sparse ternary weights (tap-list kernel), sparse weights mixing +-1 with small and 40-bit values (odd tap
    # counts, a ragged last column chunk), dense small weights and dense 40-bit weights (register-blocked kernel)
    for Mr, K, Nc, wgen in [(3, 128, 64, lambda: rng.choice([-1, 0, 0, 0, 0, 0, 0, 1], (128, 64))),
                            (2, 37, 9, lambda: rng.choice([-1, 0, 0, 0, 0, 0, 0, 1, 5, -(1 << 40)], (37, 9))),
                            (4, 5, 3, lambda: rng.integers(-9, 10, (5, 3))),
                            (2, 6, 17, lambda: rng.integers(-(1 << 40), 1 << 40, (6, 17)))]:
        x = rng.integers(0, 1 << 64, (Mr * K, L), dtype=np.uint64)
        W = wgen().astype(np.int64)
        got = E.matmul_ct_clear(E.from_np_u64(x, cuda_device), torch.from_numpy(W).to(cuda_device), Mr, K, Nc, L)
        exp = np.zeros((Mr, Nc, L), dtype=np.uint64)
        xr = x.reshape(Mr, K, L)
        for kk in range(K):
            exp += W[kk].view(np.uint64)[None, :, None] * xr[:, kk][:, None, :]
        assert np.array_equal(E.to_np_u64(got), exp.reshape(Mr * Nc, L)), (Mr, K, Nc)
    # gather-MAC with 1500 taps per output (two shared-memory rounds), zero weights and "no term" entries
    Kt = 1500
    idx = rng.integers(-1, Erows, (3, Kt)).astype(np.int32)
    w = rng.integers(-2, 3, (3, Kt)).astype(np.int64)
    bias = rng.integers(0, 1 << 64, 3, dtype=np.uint64)
    got = E.lin_gather_mac(da, torch.from_numpy(idx).to(cuda_device), torch.from_numpy(w).to(cuda_device), L,
                           E.from_np_u64(bias, cuda_device))
    assert np.array_equal(E.to_np_u64(got), oracle.lin_gather_mac(a, idx, w, bias))


def test_linear_ops_on_misaligned_row_slices(oracle, cuda_device):
    """Row slices of [E, L] tensors with odd L start 8 bytes off a 16-byte boundary: the flat vectorised path
    must not be taken for them (the CRT lanes of the wide path hand such slices around)."""
    rng = np.random.default_rng(9)
    L = 513
    a = rng.integers(0, 1 << 63, (7, L), dtype=np.int64).astype(np.uint64)
    b = rng.integers(0, 1 << 63, (7, L), dtype=np.int64).astype(np.uint64)
    da, db = E.from_np_u64(a, cuda_device), E.from_np_u64(b, cuda_device)
    for lo in (1, 2, 3):
        sa, sb = da[lo:lo + 4], db[lo:lo + 4]
        assert np.array_equal(E.to_np_u64(E.lin_add(sa, sb, L)), oracle.lin_add(a[lo:lo + 4], b[lo:lo + 4]))
        assert np.array_equal(E.to_np_u64(E.lin_sub(sa, sb, L)), oracle.lin_sub(a[lo:lo + 4], b[lo:lo + 4]))
        assert np.array_equal(E.to_np_u64(E.lin_neg(sa, L)), oracle.lin_neg(a[lo:lo + 4]))
