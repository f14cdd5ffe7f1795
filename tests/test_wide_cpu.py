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


def test_wide_parameter_search_meets_its_own_error_budget():
    from concrete_numpy_b200 import params as P

    for p, norm2 in [(9, 1.0), (12, 4.0), (16, 1.0), (16, 256.0)]:
        prm = P.select_wide_parameters(p, norm2, 1e-6)
        assert prm.wide and prm.p == p and int(np.prod(prm.moduli)) >= 1 << p
        assert P.pbs_cluster_size(prm.N, prm.k, prm.pbs_level) > 0 and P.pbs_cluster_size(prm.N, prm.k, prm.cbs_level) > 0
        assert prm.pf_base_log <= 7                      # packing keyswitch stays on the exact int8 tensor-core path
        assert P.estimate_wide_p_error(prm, norm2) <= 1e-6
        assert abs(np.log2(prm.glwe_std) - P.secure_log2_std(prm.k * prm.N)) < 1e-9
    with pytest.raises(ValueError):
        P.select_wide_parameters(17)


def test_wide_circuits_compile_with_crt_client_parameters():
    """The reference frontend's MLIR for 9..16-bit circuits (tests/golden/wide_*.json) compiles to CRT parameters;
    the client parameters announce the moduli where the reference's client looks for them (client.py:213-215)."""
    import json

    from golden_util import golden_names, load_golden
    from concrete_numpy_b200 import compiler as cc

    names = golden_names("wide_")
    assert len(names) >= 6
    for name in names:
        doc = load_golden(name)
        opt = cc.CompilationOptions.new("main")
        opt.set_global_p_error(doc["global_p_error"])
        res = cc.JITSupport.new().compile(doc["mlir"], opt)
        if res.info.n_pbs == 0:          # no table lookup: one ciphertext per value, no evaluation keys
            assert res.crypto.linear_only and not res.crypto.wide
            continue
        assert res.crypto.wide and 9 <= res.crypto.p <= 16
        cp = json.loads(res.client_parameters.serialize())
        for gate in cp["inputs"] + cp["outputs"]:
            if gate["encryption"] is not None:
                assert gate["encryption"]["encoding"]["crt"] == list(res.crypto.moduli)
        assert res.feedback.p_error <= 1e-4 and (res.feedback.complexity > 0 or res.info.n_pbs == 0)
        # round trip of the parameter document (Server.save / load path)
        assert cc.ClientParameters.unserialize(res.client_parameters.serialize()).crypto == res.crypto


def test_analysis_reduces_weights_per_modulus():
    from concrete_numpy_b200.executor import _centered

    c = np.array([100, -7, 250, 9, 0, 13])
    for m in (5, 7, 11, 13, 16):
        r = _centered(c, m)
        assert np.all((r - c) % m == 0) and r.min() >= -(m // 2) - (m % 2 == 0) * 0 and r.max() <= m // 2


def test_linear_only_circuits_take_any_width():
    """Circuits without table lookups are not limited to 16 bits in the reference (graph_converter.py:288-306;
    tests/execution/test_add.py runs x + 42 with x up to 10^6): one ciphertext per value, no bootstrap keys."""
    from concrete_numpy_b200 import params as P

    for p, norm2 in [(4, 1.0), (20, 1.0), (32, 100.0), (50, 1.0)]:
        prm = P.select_linear_parameters(p, norm2, 1e-9)
        assert prm.linear_only and not prm.wide and prm.n == 0 and prm.N % 16 == 0
        assert P.p_error_of(p, norm2 * prm.glwe_std ** 2) <= 1e-9
        assert abs(np.log2(prm.glwe_std) - P.secure_log2_std(prm.N)) < 1e-9
    assert P.select_linear_parameters(20).N < P.select_linear_parameters(40).N
    with pytest.raises(ValueError):
        P.select_linear_parameters(60)
