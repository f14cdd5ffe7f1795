"""
Clear stand-in for concrete_numpy_b200.engine used by the CPU tests of the host logic
(parser + executor plans + encodings): a "ciphertext" is a single torus word (LWE of dimension 0,
ct_words = 1) holding `m << (63 - p)`, so every linear op of the executor runs the same wrapping
u64 formulas as the CUDA kernels and a table lookup is decode -> (negacyclic) table -> encode.
TEST INFRASTRUCTURE ONLY; never imported by the product.
"""
import numpy as np
import torch


def _u(t):
    return t.detach().cpu().numpy().view(np.uint64)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64))


def from_np_u64(a, device):
    return _t(np.ascontiguousarray(a, dtype=np.uint64))


def build_lut_accumulators(params, tables):
    return np.asarray(tables).astype(np.int64).view(np.uint64)


def _rows(x, idx):
    return x if idx is None else x[idx.numpy()]


def lin_add(a, b, L, ia=None, ib=None, E=None):
    return _t(_rows(_u(a), ia) + _rows(_u(b), ib))


def lin_sub(a, b, L, ia=None, ib=None, E=None):
    return _t(_rows(_u(a), ia) - _rows(_u(b), ib))


def lin_neg(a, L, ia=None, E=None):
    return _t(np.uint64(0) - _rows(_u(a), ia))


def lin_add_plain(a, pt, L, ia=None):
    out = _rows(_u(a), ia).copy()
    out[:, -1] += _u(pt)
    return _t(out)


def lin_plain_sub(a, pt, L, ia=None):
    out = np.uint64(0) - _rows(_u(a), ia)
    out[:, -1] += _u(pt)
    return _t(out)


def lin_mul_clear(a, c, L, ia=None):
    return _t(_rows(_u(a), ia) * c.numpy().view(np.uint64)[:, None])


def lin_gather_mac(x, idx, w, L, bias=None):
    xu = _u(x)
    idx, w = idx.numpy(), w.numpy().view(np.uint64)
    out = np.zeros((idx.shape[0], xu.shape[1]), dtype=np.uint64)
    for j in range(idx.shape[1]):
        ok = idx[:, j] >= 0
        out[ok] += w[ok, j][:, None] * xu[idx[ok, j]]
    if bias is not None:
        out[:, -1] += _u(bias)
    return _t(out)


def matmul_ct_clear(x, W, Mr, K, Nc, L):
    xu = _u(x).reshape(Mr, K, L)
    Wu = W.numpy().view(np.uint64)
    out = np.zeros((Mr, Nc, L), dtype=np.uint64)
    for kk in range(K):
        out += Wu[kk][None, :, None] * xu[:, kk][:, None, :]
    return _t(out.reshape(Mr * Nc, L))


class ClearEval:
    """What executor.py receives as `ev`; carries the precision for table lookups."""

    def __init__(self, params):
        self.params = params


def decode(body_u64, p):
    t = body_u64 >> np.uint64(64 - p - 2)
    return ((t >> np.uint64(1)) + (t & np.uint64(1))) & np.uint64((1 << (p + 1)) - 1)


def _trivial_rows(bodies_u64, L):
    """[E, L] rows with zero masks and the given bodies (a clear 'ciphertext' of L torus words)."""
    out = np.zeros((bodies_u64.size, L), dtype=np.uint64)
    out[:, -1] = bodies_u64
    return _t(out)


def table_lookup(ev, ct, luts, lut_idx=None):
    p = ev.params.p
    cu = _u(ct)
    m = decode(cu[:, -1], p).astype(np.int64)
    tables = _u(luts)
    which = np.zeros(m.size, dtype=np.int64) if lut_idx is None else lut_idx.numpy().astype(np.int64)
    neg = m >= (1 << p)                       # padding bit set: the bootstrap is negacyclic
    val = tables[which, np.where(neg, m - (1 << p), m)]
    val = np.where(neg, np.uint64(0) - val, val)
    return _trivial_rows(val << np.uint64(63 - p), cu.shape[1])


# ---- client side + keys: stand-ins for engine.KeySet / EvalKeys / encrypt / decrypt, so that compiler.py (and through
# it the reference's own Client / Server / Circuit classes) can run end to end on the CPU in the host-logic tests ----

from concrete_numpy_b200.engine import CryptoParams, derive_key, fresh_key  # noqa: E402,F401  (pure-Python helpers)


class KeySet:
    def __init__(self, params, device, seed=None, keep_std_bsk=False):
        self.params = params
        self.device = torch.device("cpu")
        self.seed = fresh_key() if seed is None else seed
        self.deterministic = isinstance(self.seed, int)
        self.sk_big = torch.zeros(max(params.big_dim, 1), dtype=torch.int64)
        self.sk_small = torch.zeros(max(params.n, 1), dtype=torch.int64)
        self.ksk = torch.zeros((1, 1, 2), dtype=torch.int64)
        self.bsk_std = None
        self.bsk_f = torch.zeros((4, 2), dtype=torch.float64)
        self.pfksk = None


class EvalKeys:
    def __init__(self, params, bsk_f, ksk, pfksk=None):
        self.params, self.bsk_f, self.ksk, self.pfksk = params, bsk_f, ksk, pfksk

    @staticmethod
    def of(keyset):
        return EvalKeys(keyset.params, keyset.bsk_f, keyset.ksk, keyset.pfksk)


def encrypt(keyset, messages, first_index=0, seed=None):
    p = keyset.params
    m = np.ascontiguousarray(np.asarray(messages).reshape(-1).astype(np.uint64))
    return _trivial_rows(m << np.uint64(63 - p.p), p.ct_words)


def encrypt_plaintexts(keyset, plaintexts, first_index=0, seed=None):
    pt = np.ascontiguousarray(np.asarray(plaintexts).reshape(-1).astype(np.uint64))
    return _trivial_rows(pt, keyset.params.ct_words)


def decrypt(keyset, ct, return_phase=False):
    phase = _u(ct.contiguous())[:, -1].copy()
    msg = decode(phase, keyset.params.p)
    return (msg, phase) if return_phase else msg


# ---- wide integers (CRT residues): clear stand-ins for concrete_numpy_b200.wide.{lut_pool, wide_table_lookup} ----

def wide_lut_pool(params, luts, device):
    return _t(np.ascontiguousarray(luts))          # [tables, residues, 2^bits] encoded outputs


def wide_table_lookup(ev, cts, pool, polys_per_residue, table_idx=None):
    """Cell index of every residue word (what bit extraction reads) -> table pattern -> encoded output residues."""
    from concrete_numpy_b200 import wide

    moduli = ev.params.moduli
    luts = _u(pool)
    pattern = np.zeros(_u(cts[0]).shape[0], dtype=np.int64)
    shift = 0
    for ct, m in zip(cts, moduli):
        b, off, _, _ = wide.block_layout(int(m))
        phase = _u(ct)[:, -1] + np.uint64(off)
        pattern |= (phase >> np.uint64(64 - b)).astype(np.int64) << shift
        shift += b
    sel = np.zeros(pattern.size, dtype=np.int64) if table_idx is None else table_idx.numpy().astype(np.int64)
    return [_t(luts[sel, o, pattern][:, None]) for o in range(len(moduli))]


def install_wide(monkeypatch_or_none=None):
    """Route the executor's wide lookups to the clear stand-ins (CPU tests of the CRT lanes)."""
    from concrete_numpy_b200 import wide

    if monkeypatch_or_none is None:
        wide.lut_pool, wide.wide_table_lookup = wide_lut_pool, wide_table_lookup
    else:
        monkeypatch_or_none.setattr(wide, "lut_pool", wide_lut_pool)
        monkeypatch_or_none.setattr(wide, "wide_table_lookup", wide_table_lookup)


def run_crt(X, doc, args, group=None):
    """Execute a golden circuit on clear CRT residues: client encoding (compiler.encrypt_arguments), executor lanes,
    client decoding (compiler.decrypt_result + api.Client.decrypt) without the encryption."""
    from concrete_numpy_b200 import wide
    from concrete_numpy_b200.engine import CryptoParams
    from concrete_numpy_b200.mlir_text.parser import parse_module

    program = parse_module(doc["mlir"])
    p = X.analyze(program).precision
    moduli = wide.crt_moduli(p)
    params = CryptoParams(p=p, n=1, k=0, N=1 << 11, pbs_level=1, pbs_base_log=1, ks_level=1, ks_base_log=1,
                          lwe_std=0.0, glwe_std=0.0, moduli=moduli, cbs_level=1, cbs_base_log=1, pf_level=1, pf_base_log=1)
    ex = X.Executor(program, params, "cpu", group=group)
    vals = []
    for (_, t), a, signed in zip(program.args, args, doc["input_signs"]):
        a = np.asarray(a, dtype=np.int64)
        if t.encrypted:
            off = (1 << (t.width - 1)) if signed else 0
            pts = np.concatenate([wide.encode_residue((a + off).reshape(-1), m) for m in moduli])
            vals.append(X.Ct(_t(pts[:, None]), t.dims))
        else:
            vals.append(a)
    outs = ex.run(ClearEval(params), vals)
    M = int(np.prod([int(m) for m in moduli], dtype=object))
    res = []
    for o, t, signed in zip(outs, program.result_types, doc["output_signs"]):
        per = _u(o.data)[:, -1].reshape(len(moduli), -1)
        m = wide.crt_combine([wide.decode_residue(per[i], mod) for i, mod in enumerate(moduli)], moduli)
        if signed:
            m = np.where(m < M // 2, m, m - M)
        res.append(int(m[0]) if t.dims == () else m.reshape(t.dims))
    return res[0] if len(res) == 1 else tuple(res)
