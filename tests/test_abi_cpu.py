"""CPU-side checks: the C-ABI library loads without a GPU and exports every symbol that
include/fhe_b200.h declares; the oracle matches its own known answers."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fhe_b200.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fhe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from concrete_numpy_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/fhe_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.load().fhe_abi_version() == 1
    # geometry helper needs no device
    assert [_lib.load().fhe_pbs_cluster_size(N, 1, 2) for N in (2048, 8192, 32768)] == [1, 2, 8]


def test_product_path_fails_loudly_without_gpu():
    import pytest
    import torch

    from concrete_numpy_b200 import compiler

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compiler.default_device()


def test_oracle_chacha_known_answer(oracle):
    """RFC 8439 section 2.3.2 test vector does not apply (custom key schedule / 64-bit counter), so
    pin the generator against its own first words: a change of the RNG breaks key parity."""
    w = oracle.rng_u64(42, (3 << 56) | 5, 0, 4)
    assert w.dtype == np.uint64 and len(set(w.tolist())) == 4
    again = oracle.rng_u64(42, (3 << 56) | 5, 2, 2)
    assert np.array_equal(w[2:], again)          # random access == sequential


def test_oracle_decomposition_recomposes(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 1 << 64, 1000, dtype=np.uint64)
    for base_log, level in [(3, 5), (15, 2), (23, 1), (4, 4)]:
        d = oracle.signed_decompose(a, base_log, level)
        assert d.min() >= -(1 << (base_log - 1)) and d.max() <= (1 << (base_log - 1))
        rec = np.zeros(a.size, dtype=np.uint64)
        for lv in range(level):
            rec += d[:, lv].astype(np.uint64) << np.uint64(64 - (lv + 1) * base_log)
        err = (rec - a).view(np.int64)
        assert np.abs(err).max() <= 1 << (64 - base_log * level - 1)   # closest representable


def test_oracle_table_lookup_end_to_end(oracle):
    """keygen -> encrypt -> keyswitch -> bootstrap -> decrypt on the CPU oracle (small set)."""
    n, k, N, lb, bb, lk, bk, p = 64, 1, 512, 2, 10, 4, 4, 3
    sk_s, sk_b = oracle.keygen_lwe_secret(n, 7, 1), oracle.keygen_lwe_secret(k * N, 7, 2)
    bsk = oracle.keygen_bsk(sk_s, sk_b, k, N, lb, bb, 2.0 ** -45, 7)
    ksk = oracle.keygen_ksk(sk_b, sk_s, lk, bk, 2.0 ** -30, 7)
    rng = np.random.default_rng(1)
    m = rng.integers(0, 8, 16).astype(np.uint64)
    table = rng.integers(0, 8, 8)
    ct = oracle.encrypt(m << np.uint64(63 - p), sk_b, 2.0 ** -45, 9)
    lut = np.zeros((1, 2 * N), dtype=np.uint64)
    lut[0, N:] = oracle.build_lut_poly(table, p, p, N)
    out = oracle.pbs(oracle.keyswitch(ct, ksk, lk, bk), oracle.bsk_to_fourier(bsk), lut, None, bb)
    got = oracle.decode(oracle.decrypt_phase(out, sk_b), p)
    assert np.array_equal(got, table[m.astype(np.int64)].astype(np.uint64))


def test_plan_selection_and_python_mirror():
    """fhe_pbs_variant_plan (pure host logic, no GPU needed): variant 0 is the smallest cluster that fits, N = 2048
    with >= 3 levels takes the 256-thread plan, and params.pbs_cluster_size mirrors the cluster size for every shape."""
    import ctypes as C

    from concrete_numpy_b200 import _lib, params as P

    lib = _lib.load()

    def plan(N, k, level, v=0):
        m = (C.c_int * 8)()
        assert lib.fhe_pbs_variant_plan(N, k, level, v, m) == 0
        return tuple(m)

    assert plan(2048, 1, 1)[5:7] == (256, 2) and plan(2048, 1, 2)[5:7] == (256, 2)   # two 256-thread CTAs per SM at 128 registers
    assert plan(2048, 1, 3)[5:7] == (256, 1) and plan(2048, 1, 4)[5:7] == (256, 1)   # one CTA per SM by shared memory
    assert plan(32768, 1, 2)[0] == 8 and plan(32768, 1, 3)[0] == 16
    for logN in range(8, 16):
        for k in (1, 2, 3):
            for level in (1, 2, 3, 4):
                nv = lib.fhe_pbs_num_variants(1 << logN, k, level)
                c = lib.fhe_pbs_cluster_size(1 << logN, k, level)
                assert P.pbs_cluster_size_mirror(1 << logN, k, level) == c == P.pbs_cluster_size(1 << logN, k, level)
                assert (nv == 0) == (c < 0)
                if nv:
                    sizes = [lib.fhe_pbs_variant_cluster_size(1 << logN, k, level, v) for v in range(nv)]
                    assert sizes == sorted(sizes) and sizes[0] == c == plan(1 << logN, k, level)[0]
    assert lib.fhe_pbs_variant_plan(2048, 1, 1, 99, (C.c_int * 8)()) != 0
    # which kernels serve which plan (round 2): N = 2048 up to two levels = the grouped kernel with two ciphertext slots per CTA,
    # N = 4096 = its single-slot 512-thread plan (radix 8 x 8 x 4 x 8), the 2048-point cluster plans and the 4-CTA latency plan
    # of N = 4096 = the pipelined kernel (two levels only)
    assert [lib.fhe_pbs_variant_grouped(2048, 1, lv, 0) for lv in (1, 2, 3, 4)] == [1, 1, 0, 0]
    assert lib.fhe_pbs_variant_grouped(2048, 2, 1, 0) <= 0 and lib.fhe_pbs_variant_grouped(2048, 1, 1, 99) < 0
    assert plan(4096, 1, 2) == (1, 3, 3, 2, 3, 512, 1, 1) and lib.fhe_pbs_variant_grouped(4096, 1, 2, 0) == 0
    assert [lib.fhe_pbs_variant_pipelined(N, 1, 2, 0) for N in (32768, 16384, 8192, 4096, 2048)] == [1, 1, 1, 0, 0]
    assert lib.fhe_pbs_variant_pipelined(32768, 1, 3, 0) == 0
    clusters = [lib.fhe_pbs_variant_cluster_size(4096, 1, 2, v) for v in range(lib.fhe_pbs_num_variants(4096, 1, 2))]
    assert [lib.fhe_pbs_variant_pipelined(4096, 1, 2, v) for v in range(len(clusters))] == [int(c == 4) for c in clusters]


def test_key_material_and_cache_file_mode(oracle, tmp_path):
    """Key material is 512 bits (secret ChaCha key + separate public-mask key): 64 caller bytes are used as they are,
    an int is the deterministic TEST expansion (identical in the library and the oracle); the insecure key cache keeps
    its file private to the owner."""
    import stat

    import pytest

    from concrete_numpy_b200 import _lib, compiler as cc, engine as E

    k1, k2 = _lib.key_material(5), _lib.key_material(5)
    assert list(k1) == list(k2) and list(k1) == list(oracle.key_material(5))
    assert list(k1)[:8] != list(k1)[8:] and list(k1) != list(_lib.key_material(6))
    raw = os.urandom(64)
    assert bytes(_lib.key_material(raw)) == raw and bytes(oracle.key_material(raw)) == raw
    with pytest.raises(ValueError):
        _lib.key_material(b"too short")
    assert E.derive_key(7, 3) == 4
    d = E.derive_key(raw, 3)
    assert len(d) == 64 and d != raw and d == E.derive_key(raw, 3) and d != E.derive_key(raw, 4)
    assert len(E.fresh_key()) == 64 and E.fresh_key() != E.fresh_key()

    class FakeParams:
        @staticmethod
        def cache_key():
            return "abc"

    cache = cc.KeySetCache.new(str(tmp_path / "cache"))
    s1, s2 = cache.seed_for(FakeParams()), cache.seed_for(FakeParams())
    assert s1 == s2 and len(s1) == 64
    assert stat.S_IMODE(os.stat(tmp_path / "cache" / "abc.key").st_mode) == 0o600
