"""CPU tests of the wide-integer path: the oracle's restatement of the CRT / without-padding lookup decrypts to the
clear table value, and the product's host helpers (CRT layout, moduli, table expansion) agree with the oracle's."""
import numpy as np
import pytest

from oracle import wop as W
from concrete_numpy_b200 import wide


@pytest.mark.parametrize("m", [2, 3, 4, 5, 7, 8, 9, 11, 13, 15, 16, 17, 25, 31, 32])
def test_block_layout_matches_oracle_brute_force(m):
    b, off, dec, margin = wide.block_layout(m)
    ob, oc64, odec, omargin = W.block_layout(m)
    assert b == ob and list(dec) == odec.tolist()
    assert off == ((oc64 << (64 - ob)) >> 6)
    assert abs(margin - omargin) < 1e-9 and margin > 0


def test_residue_codec_and_crt():
    rng = np.random.default_rng(0)
    for moduli in [(7, 8, 9, 11, 13), (3, 4, 5), (16, 9, 5)]:
        M = int(np.prod(moduli))
        x = rng.integers(0, M, 500)
        res = []
        for m in moduli:
            enc = wide.encode_residue(x, m)
            assert np.array_equal(enc, W.encode_residue(x, m))
            noise = rng.integers(-(1 << 50), 1 << 50, x.size).astype(np.int64).view(np.uint64)
            got = wide.decode_residue(enc + noise, m)
            assert np.array_equal(got, x % m) and np.array_equal(got, W.decode_residue(enc + noise, m))
            res.append(got)
        assert np.array_equal(wide.crt_combine(res, moduli), x)


@pytest.mark.parametrize("p", range(9, 17))
def test_crt_moduli_cover_the_width(p):
    ms = wide.crt_moduli(p)
    assert int(np.prod(ms)) >= 1 << p
    for i, a in enumerate(ms):
        for b in ms[i + 1:]:
            assert np.gcd(a, b) == 1
    assert wide.total_bits(ms) <= p + 3 and max(ms) <= 16


def test_crt_table_expansion_matches_oracle():
    moduli = (3, 4, 5)
    table = (np.arange(60) * 7 + 3) % 60 - 20
    luts = wide.build_crt_luts(table, moduli)
    assert luts.shape == (1, 3, 1 << wide.total_bits(moduli))
    assert np.array_equal(luts[0], W.build_crt_lut(table, moduli, moduli))


def test_oracle_wide_lookup_decrypts_to_table(oracle):
    """Toy parameters (tiny noise), N = 64 so that the 7 extracted bits need one CMUX-tree level."""
    p = W.WopParams(n=16, k=1, N=64, lwe_std=2.0 ** -40, glwe_std=2.0 ** -50, ks_level=8, ks_base_log=4,
                    pbs_level=4, pbs_base_log=8, cbs_level=3, cbs_base_log=6, pf_level=6, pf_base_log=7)
    keys = W.WopKeys(oracle, p, 7)
    moduli = [3, 4, 5]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.integers(0, 60, 10), [0, 59]])
    table = (np.arange(60) ** 2 + 7) % 60
    cts = [oracle.encrypt(W.encode_residue(x % m, m), keys.sk_big, p.glwe_std, 11 + i) for i, m in enumerate(moduli)]
    outs = W.wide_lookup(oracle, keys, cts, moduli, table)
    res = [W.decode_residue(oracle.decrypt_phase(o, keys.sk_big), m) for o, m in zip(outs, moduli)]
    assert np.array_equal(W.crt_combine(res, moduli), table[x])
