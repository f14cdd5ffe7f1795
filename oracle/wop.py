"""
CPU restatement of the wide-integer (9..16-bit) table lookup: CRT residues without padding bit, bit
extraction, circuit bootstrapping, vertical packing (CMUX tree + blind rotation).

TEST INFRASTRUCTURE ONLY (see oracle/tfhe_ref.c).  PARITY UNPINNED at ciphertext level: the reference
reaches this path inside the absent concrete-compiler wheel (`crt` decomposition announced in the client
parameters, concrete/numpy/compilation/client.py:213-235; widths up to MAXIMUM_TLU_BIT_WIDTH = 16,
concrete/numpy/mlir/utils.py:16).  The algorithm restated here is the published one (Bergerat, Boudi,
Bourgerie, Chillotti, Ligier, Orfila, Tap: "Parameter Optimization & Larger Precision for (T)FHE", 2022,
sections on WoP-PBS: bit extraction, circuit bootstrapping with private functional packing keyswitches,
vertical packing), pinned at the decrypted-output level only.

Everything is composed from the C primitives of tfhe_ref.c (keyswitch, bootstrap, CMUX) with numpy glue,
written for clarity, independent of the product's host code (concrete-numpy_b200/wide.py).
"""

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from .oracle import DOM_SK_BIG, DOM_SK_SMALL, Oracle

U = np.uint64
PF_SEED_XOR = 0x9E3779B97F4A7C15      # the packing key must not reuse the bootstrap key's mask streams
PF_PAD = 16                           # extended key length kN + 16 (tensor-core keyswitch wants kN % 16 == 0)


@dataclass(frozen=True)
class WopParams:
    n: int
    k: int
    N: int
    lwe_std: float
    glwe_std: float
    ks_level: int
    ks_base_log: int
    pbs_level: int
    pbs_base_log: int
    cbs_level: int
    cbs_base_log: int
    pf_level: int
    pf_base_log: int


# ----------------------------------------------------------------- CRT block layout -----

def block_bits(m: int) -> int:
    return max(1, int(m - 1).bit_length())


def block_layout(m: int):
    """(bits b, offset in 1/64ths of a cell, decode table pattern -> residue, margin in cells).
    A residue x sits at x * 2^b / m + c cells; the extracted b-bit pattern is the cell index.  c is chosen
    (brute force) to maximise the distance of every position to the cells that decode to another residue."""
    b = block_bits(m)
    cells = 1 << b
    s = cells / m
    best = None
    for c64 in range(64):
        c = c64 / 64.0
        pos = np.arange(m) * s + c
        centers = np.arange(cells) + 0.5
        d = np.abs(centers[:, None] - pos[None, :])
        d = np.minimum(d, cells - d)
        dec = np.argmin(d, axis=1)
        margin = np.inf
        for x in range(m):
            cell = int(np.floor(pos[x])) % cells
            if dec[cell] != x:
                margin = -1.0
                break
            lo = cell
            while dec[(lo - 1) % cells] == x and (lo - 1) % cells != cell:
                lo -= 1
            hi = cell + 1
            while dec[hi % cells] == x and hi % cells != cell:
                hi += 1
            margin = min(margin, pos[x] - lo, hi - pos[x])
        if best is None or margin > best[0] + 1e-12:
            best = (margin, c64, dec.copy())
    return b, best[1], best[2].astype(np.int64), float(best[0])


def encode_residue(x, m: int) -> np.ndarray:
    """x mod m -> round(x * 2^64 / m) as u64 (no padding bit)."""
    table = np.array([((int(v) << 64) + m // 2) // m % (1 << 64) for v in range(m)], dtype=np.uint64)
    return table[np.mod(np.asarray(x, dtype=np.int64), m)]


def decode_residue(phase: np.ndarray, m: int) -> np.ndarray:
    out = np.empty(phase.shape, dtype=np.int64)
    flat = phase.reshape(-1)
    o = out.reshape(-1)
    for i, ph in enumerate(flat):
        o[i] = ((int(ph) * m + (1 << 63)) >> 64) % m
    return out


def crt_combine(residues: Sequence[np.ndarray], moduli: Sequence[int]) -> np.ndarray:
    M = int(np.prod([int(m) for m in moduli], dtype=object))
    acc = np.zeros(np.asarray(residues[0]).shape, dtype=np.int64)
    for r, m in zip(residues, moduli):
        Mi = M // m
        acc = (acc + np.asarray(r, dtype=np.int64) * (Mi * pow(Mi, -1, m) % M)) % M
    return acc


# ----------------------------------------------------------------- keys ----------------

class WopKeys:
    def __init__(self, orc: Oracle, p: WopParams, seed: int):
        self.p = p
        kN = p.k * p.N
        self.sk_small = orc.keygen_lwe_secret(p.n, seed, DOM_SK_SMALL)
        self.sk_big = orc.keygen_lwe_secret(kN, seed, DOM_SK_BIG)
        self.ksk = orc.keygen_ksk(self.sk_big, self.sk_small, p.ks_level, p.ks_base_log, p.lwe_std, seed)
        bsk = orc.keygen_bsk(self.sk_small, self.sk_big, p.k, p.N, p.pbs_level, p.pbs_base_log, p.glwe_std, seed)
        self.bsk_f = orc.bsk_to_fourier(bsk)
        # packing keyswitch key: the "bootstrap key" of the extended big key (s, -1, 0...0): row (i, lv, r) is a
        # GLWE encryption of zero with s'_i * 2^(64-(lv+1)*b) added on polynomial r -- all k+1 private functional
        # keyswitches (multiplication by -S_r, r < k, and by 1) of circuit bootstrapping in one key
        ext = np.zeros(kN + PF_PAD, dtype=np.uint64)
        ext[:kN] = self.sk_big
        ext[kN] = U(0xFFFFFFFFFFFFFFFF)
        pf = orc.keygen_bsk(ext, self.sk_big, p.k, p.N, p.pf_level, p.pf_base_log, p.glwe_std, seed ^ PF_SEED_XOR)
        self.pfksk = pf.reshape(kN + PF_PAD, p.pf_level, (p.k + 1) * (p.k + 1) * p.N)


# ----------------------------------------------------------------- the lookup ----------

def bootstrap_theta(p: WopParams) -> int:
    """log2 of the functions one blind rotation evaluates: the extraction weight of a bit and its cbs_level
    circuit-bootstrap levels (multi-output bootstrap, ref_pbs_batch_ex)."""
    return max(1, int(1 + p.cbs_level - 1).bit_length())


def sign_bootstrap_multi(orc: Oracle, keys: WopKeys, small: np.ndarray, ext_weights: Sequence[int], idx) -> np.ndarray:
    """One blind rotation per row of `small` (phase in [0, 1/2) = bit 0, [1/2, 1) = bit 1), 1 + cbs_level results
    [R, 1 + cbs_level, kN+1]: bit * ext_weights[idx[e]] and bit * 2^(64-(lv+1)*cbs_base_log).  Function u has the
    constant test polynomial -w_u/2 on the coefficients congruent to u mod 2^theta; + w_u/2 afterwards."""
    p = keys.p
    theta, nout = bootstrap_theta(p), 1 + p.cbs_level
    cbs_w = [1 << (64 - (lv + 1) * p.cbs_base_log) for lv in range(p.cbs_level)]
    luts = np.zeros((len(ext_weights), (p.k + 1) * p.N), dtype=np.uint64)
    half = np.zeros((len(ext_weights), nout), dtype=np.uint64)
    for a, w in enumerate(ext_weights):
        ws = [w] + cbs_w
        for u, wu in enumerate(ws):
            luts[a, p.k * p.N + u::1 << theta] = U((-(wu >> 1)) % (1 << 64))
            half[a, u] = U((wu >> 1) % (1 << 64))
    idx = np.asarray(idx, dtype=np.int32)
    out = orc.pbs_ex(small, keys.bsk_f, luts, idx, p.pbs_base_log, p.n, 0, theta, nout)
    out[:, :, -1] += half[idx]
    return out


def extract_and_bootstrap(orc: Oracle, keys: WopKeys, ct: np.ndarray, m: int):
    """Bits (LSB first) of the cell index of the residues mod m in `ct` [E, kN+1].  Returns (bits, levels): per bit the
    small-key ciphertext that was tested (phase in [0, 1/2) = 0) and the cbs_level LWEs bit * 2^(64-(lv+1)*b) under the
    big key [E, cbs_level, kN+1].  The bit is subtracted from the residue with the result of the SAME blind rotation,
    so extraction and circuit bootstrapping can never disagree on a residue sitting on a cell boundary."""
    p = keys.p
    b, c64, _, _ = block_layout(m)
    buf = ct.copy()
    buf[:, -1] += U((c64 << (64 - b)) >> 6)
    bits, levels = [], []
    for t in range(b):
        sh = buf * U(1 << (b - 1 - t))
        ks = orc.keyswitch(sh, keys.ksk, p.ks_level, p.ks_base_log)
        if t:
            # the already removed lower bits leave the fraction in [0, 2^-t): centre it inside the half torus
            ks[:, -1] += U((1 << 62) - (1 << (62 - t)))
        got = sign_bootstrap_multi(orc, keys, ks, [1 << (64 - b + t)], np.zeros(len(ks), dtype=np.int32))
        bits.append(ks)
        levels.append(got[:, 1:])
        if t < b - 1:
            buf = buf - got[:, 0]
    return bits, levels


def circuit_bootstrap(orc: Oracle, keys: WopKeys, lwe: np.ndarray) -> np.ndarray:
    """lwe [G, cbs_level, kN+1] (bit * gadget factor per level) -> Fourier GGSWs [G, cbs_level, k+1, k+1, M, 2]:
    one packing keyswitch per level turns the LWE into the k+1 GLWE rows of that level."""
    p = keys.p
    G = lwe.shape[0]
    kN = p.k * p.N
    flat = lwe.reshape(G * p.cbs_level, kN + 1)
    ext = np.zeros((flat.shape[0], kN + PF_PAD + 1), dtype=np.uint64)
    ext[:, :kN + 1] = flat                                            # the body becomes mask word kN (key entry -1)
    rows = orc.keyswitch(ext, keys.pfksk, p.pf_level, p.pf_base_log)  # [G*lc, (k+1)^2 N]
    ggsw = rows.reshape(G, p.cbs_level, p.k + 1, p.k + 1, p.N)
    return orc.bsk_to_fourier(ggsw)


def vertical_packing(orc: Oracle, keys: WopKeys, ggsw_f: np.ndarray, E: int, T: int, lut: np.ndarray,
                     lut_sel=None) -> np.ndarray:
    """ggsw_f [E*T, ...] (bit t of ciphertext e at e*T+t, LSB first); lut [tables, 2^T] u64 -> LWE [E, kN+1] holding
    lut[sel[e]][index of e]."""
    p = keys.p
    N, k = p.N, p.k
    logN = N.bit_length() - 1
    nrot = min(T, logN)
    polys = 1 << (T - nrot)
    ntab = lut.shape[0]
    sel = np.zeros(E, dtype=np.int64) if lut_sel is None else np.asarray(lut_sel, dtype=np.int64)
    pool = np.zeros((ntab * polys, (k + 1) * N), dtype=np.uint64)
    pool[:, k * N:k * N + min(N, 1 << T)] = lut.reshape(ntab * polys, -1)
    # CMUX tree over bits nrot..T-1 (bit nrot selects between neighbouring polynomials)
    cur_idx = (sel[:, None] * polys + np.arange(polys)[None, :])     # [E, polys] indices into pool
    cur_pool = pool
    for lvl in range(T - nrot):
        cnt = cur_idx.shape[1] // 2
        i0 = cur_idx[:, 0::2].reshape(-1)
        i1 = cur_idx[:, 1::2].reshape(-1)
        gi = np.repeat(np.arange(E) * T + nrot + lvl, cnt)
        cur_pool = orc.cmux_batch(cur_pool, i0, i1, ggsw_f, gi, p.cbs_base_log)
        cur_idx = np.arange(E * cnt).reshape(E, cnt)
    acc_idx = cur_idx[:, 0]
    # blind rotation by the low bits: a synthetic LWE input whose mask words modulus-switch to 2N - 2^t
    lwe = np.zeros((E, nrot + 1), dtype=np.uint64)
    for t in range(nrot):
        lwe[:, t] = U(((2 * N - (1 << t)) << (64 - (logN + 1))) % (1 << 64))
    M = N // 2
    stride = T * p.cbs_level * (k + 1) * (k + 1) * M
    return orc.pbs_ex(lwe, ggsw_f, cur_pool, acc_idx, p.cbs_base_log, nrot, stride)


def build_crt_lut(table: np.ndarray, moduli_in: Sequence[int], moduli_out: Sequence[int]) -> np.ndarray:
    """[len(moduli_out), 2^T] u64: for every pattern of extracted bits the encoded output residues of table[x]."""
    lay = [block_layout(m) for m in moduli_in]
    T = sum(l[0] for l in lay)
    idx = np.arange(1 << T, dtype=np.int64)
    res, shift = [], 0
    for (b, _, dec, _), m in zip(lay, moduli_in):
        res.append(dec[(idx >> shift) & ((1 << b) - 1)])
        shift += b
    x = crt_combine(res, moduli_in)
    table = np.asarray(table, dtype=np.int64)
    y = np.where(x < len(table), table[np.minimum(x, len(table) - 1)], 0)
    return np.stack([encode_residue(y, m) for m in moduli_out])


def wide_lookup(orc: Oracle, keys: WopKeys, cts: List[np.ndarray], moduli: Sequence[int], table: np.ndarray,
                moduli_out=None) -> List[np.ndarray]:
    """cts: one [E, kN+1] array per CRT block of the inputs; returns one array per output block."""
    moduli_out = list(moduli) if moduli_out is None else list(moduli_out)
    E = cts[0].shape[0]
    levels = []
    for ct, m in zip(cts, moduli):
        levels += extract_and_bootstrap(orc, keys, ct, m)[1]
    T = len(levels)
    stacked = np.stack(levels, axis=1).reshape(E * T, keys.p.cbs_level, -1)      # [e][t][lv]
    ggsw_f = circuit_bootstrap(orc, keys, stacked)
    lut = build_crt_lut(table, moduli, moduli_out)
    return [vertical_packing(orc, keys, ggsw_f, E, T, lut[o:o + 1]) for o in range(len(moduli_out))]
