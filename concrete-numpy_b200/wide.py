"""
Wide integers (9..16 bits): CRT residues without padding bit and the lookup that goes with them.

The reference allows table lookups up to `MAXIMUM_TLU_BIT_WIDTH = 16` (concrete/numpy/mlir/utils.py:16); beyond
the native bootstrap range its compiler switches the circuit to a CRT encoding announced in the client
parameters (`encoding.crt`, concrete/numpy/compilation/client.py:213-235) and evaluates tables with the
"without padding" bootstrap of Bergerat et al., "Parameter Optimization & Larger Precision for (T)FHE" (2022).
That arithmetic lives in the absent wheel; this module is the B200 host composition of the published
algorithm on top of the C ABI:

  value x  ->  residues x mod m_i, each one LWE ciphertext of phase (x mod m_i) * 2^64 / m_i (no padding bit);
               add / sub / neg / clear multiplication act residue-wise on the torus (mod m_i for free)
  lookup   ->  1. bit extraction per residue   keyswitch + ONE multi-output sign bootstrap per bit (the bit's weight
                                               for the subtraction and its cbs_level gadget levels come out of the
                                               same blind rotation)                    fhe_keyswitch_batch*, fhe_pbs_batch_multi
               2. circuit bootstrapping        per level ONE packing keyswitch with (k+1)^2 N columns (all k+1
                                               private functional keyswitches at once: the key is the bootstrap-key
                                               construction applied to the extended big key (s, -1)),
                                               Fourier conversion                       fhe_bsk_to_fourier
               3. vertical packing             CMUX tree over the high bits            fhe_cmux_batch
                                               blind rotation by the low log2(N) bits  fhe_blind_rotate_batch
             one pass of step 3 per output residue; steps 1-2 are shared.

PyTorch only allocates buffers and moves rows; every ciphertext operation is a C-ABI call.
"""

import os
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import engine as E

PF_SEED_XOR = 0x9E3779B97F4A7C15      # packing key: own mask / noise streams
PF_PAD = 16                           # extended big key (s, -1, 0 x 15): the tensor-core keyswitch wants kN % 16 == 0


# ------------------------------------------------------------------ CRT encoding (host) ----

def block_bits(m: int) -> int:
    return max(1, int(m - 1).bit_length())


@lru_cache(maxsize=None)
def block_layout(m: int) -> Tuple[int, int, Tuple[int, ...], float]:
    """(bits b, body offset, decode table, margin) of residues mod m.

    Residue x has phase x/m, i.e. sits x * 2^b / m cells into the b-bit cell grid that bit extraction reads.
    Powers of two are shifted by half a cell (offset) so that noise stays inside the cell; for the other moduli
    neighbouring residues are more than one cell apart and every cell is decoded to the nearest residue
    (cyclically); `margin` is the smallest distance (in cells) of a residue to a cell decoded differently."""
    b = block_bits(m)
    cells = 1 << b
    if m == cells:
        return b, 1 << (64 - b - 1), tuple(range(m)), 0.5
    dec = tuple(int(((2 * j + 1) * m + cells) // (2 * cells)) % m for j in range(cells))
    margin = float(cells)
    for x in range(m):
        pos = x * cells / m
        lo = hi = int(pos)                      # the cell the position falls in decodes to x
        while dec[(lo - 1) % cells] == x:
            lo -= 1
        while dec[(hi + 1) % cells] == x:
            hi += 1
        margin = min(margin, pos - lo, hi + 1 - pos)
    return b, 0, dec, margin


def crt_moduli(p: int) -> Tuple[int, ...]:
    """Pairwise coprime moduli (odd numbers and at most one power of two) with product >= 2^p, fewest extracted
    bits first, then the largest margin (as a fraction of the torus)."""
    return _crt_moduli(int(p))


@lru_cache(maxsize=None)
def _crt_moduli(p: int) -> Tuple[int, ...]:
    # residues of at most 4 bits: bit extraction amplifies the input noise by 2^(bits-1)
    # (15 is left out: 16/15 cells per residue leave almost no margin)
    cands = [m for m in range(3, 15) if m % 2 == 1] + [2, 4, 8, 16]
    best = None
    target = 1 << p

    def rec(start, chosen, prod, bits, margin):
        nonlocal best
        if prod >= target:
            key = (bits, -margin, len(chosen))
            if best is None or key < best[0]:
                best = (key, tuple(sorted(chosen)))
            return
        if best is not None and bits >= best[0][0] + 1:
            return
        for i in range(start, len(cands)):
            m = cands[i]
            if any(np.gcd(m, c) != 1 for c in chosen):
                continue
            b, _, _, cells = block_layout(m)
            rec(i + 1, chosen + [m], prod * m, bits + b, min(margin, cells / (1 << b)))

    rec(0, [], 1, 0, 1.0)
    return best[1]


def encode_residue(x, m: int) -> np.ndarray:
    """(x mod m) -> round((x mod m) * 2^64 / m), u64."""
    table = np.array([(((v << 64) + m // 2) // m) % (1 << 64) for v in range(m)], dtype=np.uint64)
    return table[np.mod(np.asarray(x, dtype=np.int64), m)]


def decode_residue(phase: np.ndarray, m: int) -> np.ndarray:
    """round(phase * m / 2^64) mod m (exact integer arithmetic on the high / low halves)."""
    ph = np.asarray(phase, dtype=np.uint64)
    hi = (ph >> np.uint64(32)).astype(np.uint64)
    lo = (ph & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    t = hi * np.uint64(m) + ((lo * np.uint64(m)) >> np.uint64(32))          # floor(ph * m / 2^32), < 2^37
    return (((t + np.uint64(1 << 31)) >> np.uint64(32)) % np.uint64(m)).astype(np.int64)


def crt_combine(residues: Sequence[np.ndarray], moduli: Sequence[int]) -> np.ndarray:
    M = 1
    for m in moduli:
        M *= int(m)
    acc = np.zeros(np.asarray(residues[0]).shape, dtype=np.int64)
    for r, m in zip(residues, moduli):
        Mi = M // int(m)
        acc = (acc + np.asarray(r, dtype=np.int64) * ((Mi * pow(Mi, -1, int(m))) % M)) % M
    return acc


def total_bits(moduli: Sequence[int]) -> int:
    return sum(block_bits(int(m)) for m in moduli)


def build_crt_luts(tables: np.ndarray, moduli: Sequence[int]) -> np.ndarray:
    """[T, 2^p] integer tables -> [T, r, 2^bits] u64: for every pattern of extracted bits (residue 0 in the low
    bits) the encoded output residues of table[x], x = CRT of the decoded input residues (0 outside the table)."""
    tables = np.asarray(tables, dtype=np.int64)
    if tables.ndim == 1:
        tables = tables[None, :]
    bits = total_bits(moduli)
    idx = np.arange(1 << bits, dtype=np.int64)
    res, shift = [], 0
    for m in moduli:
        b, _, dec, _ = block_layout(int(m))
        res.append(np.asarray(dec, dtype=np.int64)[(idx >> shift) & ((1 << b) - 1)])
        shift += b
    x = crt_combine(res, moduli)
    inside = x < tables.shape[1]
    y = np.where(inside[None, :], tables[:, np.minimum(x, tables.shape[1] - 1)], 0)
    return np.stack([encode_residue(y, int(m)) for m in moduli], axis=1)


# ------------------------------------------------------------------ keys -------------------

def keygen_packing_key(params, sk_big: torch.Tensor, seed) -> torch.Tensor:
    """[(kN + PF_PAD), pf_level, (k+1)^2 N]: fhe_keygen_bsk on the extended big key (s, -1, 0...)."""
    p = params
    dev = sk_big.device
    kN = p.big_dim
    ext = torch.zeros(kN + PF_PAD, dtype=torch.int64, device=dev)
    ext[:kN] = sk_big
    ext[kN] = -1
    key = torch.empty((kN + PF_PAD, p.pf_level, (p.k + 1) * (p.k + 1) * p.N), dtype=torch.int64, device=dev)
    scratch = torch.empty(p.k * p.N + p.k + 8, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fhe_keygen_bsk", key.data_ptr(), ext.data_ptr(), kN + PF_PAD, sk_big.data_ptr(), p.k, p.N,
                  p.pf_level, p.pf_base_log, p.glwe_std, _lib.key_material(E.derive_key(seed, PF_SEED_XOR)),
                  scratch.data_ptr(), E._stream())
    return key


class PackingKey:
    """The packing keyswitch as seen by the keyswitch kernels: kN + PF_PAD inputs, (k+1)^2 N - 1 'mask' outputs."""

    def __init__(self, params, key: torch.Tensor):
        self.params = params
        self.key = key
        self.kin = params.big_dim + PF_PAD
        self.nout = (params.k + 1) * (params.k + 1) * params.N - 1
        lib = _lib.load()
        self.tc_ok = bool(lib.fhe_ks_tc_supported(self.kin, self.nout, params.pf_level, params.pf_base_log))
        self._tc = None

    def tc(self) -> torch.Tensor:
        if self._tc is None:
            lib = _lib.load()
            p = self.params
            buf = torch.empty(lib.fhe_ksk_tc_bytes(self.kin, self.nout, p.pf_level), dtype=torch.uint8, device=self.key.device)
            with torch.cuda.device(self.key.device):
                _lib.call("fhe_ksk_to_tc", buf.data_ptr(), self.key.data_ptr(), self.kin, self.nout, p.pf_level, E._stream())
            self._tc = buf
        return self._tc

    def apply(self, lwe_ext: torch.Tensor) -> torch.Tensor:
        """lwe_ext [B, kin + 1] (body moved to mask slot kN, last word 0) -> [B, (k+1)^2 N] GGSW level rows."""
        p = self.params
        B = lwe_ext.shape[0]
        out = torch.empty((B, self.nout + 1), dtype=torch.int64, device=lwe_ext.device)
        lib = _lib.load()
        with torch.cuda.device(lwe_ext.device):
            if self.tc_ok and os.environ.get("CNP_B200_KS_TC", "1") not in ("", "0"):
                key = self.tc()
                rows = max(128, ((1 << 30) // max(lib.fhe_keyswitch_tc_scratch_bytes(128, self.kin, p.pf_level), 1)) * 128)
                for lo in range(0, B, rows):
                    hi = min(B, lo + rows)
                    scratch = torch.empty(lib.fhe_keyswitch_tc_scratch_bytes(hi - lo, self.kin, p.pf_level),
                                          dtype=torch.uint8, device=lwe_ext.device)
                    _lib.call("fhe_keyswitch_batch_tc", out[lo:hi].data_ptr(), lwe_ext[lo:hi].data_ptr(), hi - lo,
                              key.data_ptr(), scratch.data_ptr(), self.kin, self.nout, p.pf_level, p.pf_base_log, E._stream())
            else:
                _lib.call("fhe_keyswitch_batch", out.data_ptr(), lwe_ext.data_ptr(), B, self.key.data_ptr(), self.kin,
                          self.nout, p.pf_level, p.pf_base_log, E._stream())
        return out


# ------------------------------------------------------------------ the lookup --------------

def _i64(v: int) -> int:
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def bootstrap_theta(params) -> int:
    """log2 of the functions one blind rotation of the lookup evaluates (fhe_pbs_batch_multi): the extraction weight
    of a bit plus its cbs_level circuit-bootstrap levels."""
    return max(1, int(params.cbs_level).bit_length())


def sign_bootstrap_multi(ev, small: torch.Tensor, ext_weights: Sequence[int], idx: torch.Tensor) -> torch.Tensor:
    """One blind rotation per row of `small` (phase in [0, 1/2) = bit 0, [1/2, 1) = bit 1) -> [R, 1 + cbs_level, kN+1]:
    bit * ext_weights[idx[e]] and bit * 2^(64-(lv+1)*cbs_base_log).  Function u owns the accumulator coefficients
    congruent to u mod 2^theta (constant -w_u/2); + w_u/2 on the bodies afterwards."""
    p = ev.params
    dev = small.device
    theta, nout = bootstrap_theta(p), 1 + p.cbs_level
    key = ("msign", tuple(ext_weights))
    if key not in ev.wide_cache:
        cbs_w = [1 << (64 - (lv + 1) * p.cbs_base_log) for lv in range(p.cbs_level)]
        acc = np.zeros((len(ext_weights), (p.k + 1) * p.N), dtype=np.uint64)
        half = np.zeros((len(ext_weights), nout), dtype=np.int64)
        for a, w in enumerate(ext_weights):
            for u, wu in enumerate([w] + cbs_w):
                acc[a, p.k * p.N + u::1 << theta] = np.uint64((-(wu >> 1)) % (1 << 64))
                half[a, u] = _i64(wu >> 1)
        ev.wide_cache[key] = (E.from_np_u64(acc, dev), torch.from_numpy(half).to(dev))
    luts, half = ev.wide_cache[key]
    R = small.shape[0]
    out = torch.empty((R, nout, p.ct_words), dtype=torch.int64, device=dev)
    v = ev.variant_for_batch(R)
    with torch.cuda.device(dev):
        _lib.call("fhe_pbs_batch_multi", out.data_ptr(), small.data_ptr(), R, ev.bsk_for_variant(v).data_ptr(),
                  luts.data_ptr(), idx.data_ptr(), p.n, p.k, p.N, p.pbs_level, p.pbs_base_log, theta, nout, v, E._stream())
    pt = half[idx.long()].reshape(-1).contiguous()
    return E.lin_add_plain(out.view(R * nout, p.ct_words), pt, p.ct_words).view(R, nout, p.ct_words)


def extract_and_bootstrap(ev, cts: Sequence[torch.Tensor], moduli: Sequence[int]):
    """Bits (LSB first) of the cell index of every residue.  Returns (bits, levels), one list per residue: the
    small-key ciphertexts that were tested (phase in [0, 1/2) = 0; kept for tests) and, per bit, the cbs_level LWEs
    bit * 2^(64-(lv+1)*b) under the big key, [E, cbs_level, kN+1].

    Bit t of all residues goes through one keyswitch and one multi-output bootstrap launch; the bit is subtracted
    from the residue with the result of the SAME blind rotation that feeds circuit bootstrapping, so the two can
    never disagree on a residue sitting on a cell boundary."""
    p = ev.params
    L = p.ct_words
    dev = cts[0].device
    n_ct = cts[0].shape[0]
    lay = [block_layout(int(m)) for m in moduli]

    def const(v, rows=n_ct):
        return torch.full((rows,), _i64(v), dtype=torch.int64, device=dev)

    bufs = [E.lin_add_plain(ct, const(off), L) if off else ct for ct, (_, off, _, _) in zip(cts, lay)]
    bits: List[List[torch.Tensor]] = [[] for _ in moduli]
    levels: List[List[torch.Tensor]] = [[] for _ in moduli]
    for t in range(max(b for b, _, _, _ in lay)):
        act = [i for i, (b, _, _, _) in enumerate(lay) if b > t]
        sh = torch.cat([E.lin_mul_clear(bufs[i], const(1 << (lay[i][0] - 1 - t)), L) if lay[i][0] - 1 - t else bufs[i]
                        for i in act], dim=0)
        ks = E.keyswitch(ev, sh)
        if t:
            # the removed lower bits leave a fraction in [0, 2^-t) of the half torus: centre it
            ks = E.lin_add_plain(ks, const((1 << 62) - (1 << (62 - t)), ks.shape[0]), p.n + 1)
        weights = [1 << (64 - lay[i][0] + t) for i in act]
        idx = torch.arange(len(act), dtype=torch.int32, device=dev).repeat_interleave(n_ct)
        got = sign_bootstrap_multi(ev, ks, weights, idx)                 # [len(act) * n_ct, 1 + lc, L]
        for j, i in enumerate(act):
            part = got[j * n_ct:(j + 1) * n_ct]
            bits[i].append(ks[j * n_ct:(j + 1) * n_ct])
            levels[i].append(part[:, 1:])
            if t < lay[i][0] - 1:
                bufs[i] = E.lin_sub(bufs[i], part[:, 0].contiguous(), L)
    return bits, levels


def circuit_bootstrap(ev, lwe: torch.Tensor) -> torch.Tensor:
    """lwe [G, cbs_level, kN+1] (bit * gadget factor per level) -> Fourier GGSWs (G * fhe_bsk_fourier_elems(1, ...)
    complex values, plan of the vertical-packing external products): one packing keyswitch row per level."""
    p = ev.params
    G = lwe.shape[0]
    dev = lwe.device
    kN = p.big_dim
    lc = p.cbs_level
    ext = torch.zeros((G * lc, kN + PF_PAD + 1), dtype=torch.int64, device=dev)
    ext[:, :kN + 1] = lwe.reshape(G * lc, kN + 1)                    # body -> mask slot kN (key entry -1)
    rows = ev.packing.apply(ext)                                     # [G*lc, (k+1)^2 N] = [G][lv][r][(k+1)N]
    lib = _lib.load()
    out = torch.empty((lib.fhe_bsk_fourier_elems(G, p.k, p.N, lc), 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fhe_bsk_to_fourier", out.data_ptr(), rows.data_ptr(), G, p.k, p.N, lc, p.cbs_base_log, E._stream())
    return out


@lru_cache(maxsize=None)
def _rotation_masks(N: int, nrot: int) -> np.ndarray:
    logN = N.bit_length() - 1
    lwe = np.zeros(nrot + 1, dtype=np.uint64)
    for t in range(nrot):
        lwe[t] = np.uint64(((2 * N - (1 << t)) << (64 - (logN + 1))) % (1 << 64))
    return lwe


def lut_pool(params, luts: np.ndarray, device) -> torch.Tensor:
    """[tables, r, 2^bits] u64 -> trivial GLWE pool [tables * r * polys, (k+1)N] (polynomial t of a table holds
    entries [t*N, (t+1)*N))."""
    N, k = params.N, params.k
    ntab, r, size = luts.shape
    polys = max(size // N, 1)
    pool = np.zeros((ntab * r * polys, (k + 1) * N), dtype=np.uint64)
    pool[:, k * N:k * N + min(N, size)] = luts.reshape(ntab * r * polys, -1)
    return E.from_np_u64(pool, device)


def vertical_packing(ev, ggsw_f: torch.Tensor, n_ct: int, bits: int, pool: torch.Tensor, first_poly: torch.Tensor) -> torch.Tensor:
    """first_poly [n_ct] int64: index in `pool` of polynomial 0 of the table each ciphertext looks up.
    Returns LWE [n_ct, kN+1] (big key) of the selected entries."""
    p = ev.params
    N, k = p.N, p.k
    dev = ggsw_f.device
    logN = N.bit_length() - 1
    nrot = min(bits, logN)
    polys = 1 << (bits - nrot)
    glwe = (k + 1) * N
    ar = torch.arange(n_ct, dtype=torch.int64, device=dev)
    cur_idx = first_poly.to(torch.int64)[:, None] + torch.arange(polys, dtype=torch.int64, device=dev)[None, :]
    cur_pool = pool
    with torch.cuda.device(dev):
        for lvl in range(bits - nrot):
            cnt = cur_idx.shape[1] // 2
            i0 = cur_idx[:, 0::2].reshape(-1).to(torch.int32).contiguous()
            i1 = cur_idx[:, 1::2].reshape(-1).to(torch.int32).contiguous()
            gi = (ar * bits + nrot + lvl).repeat_interleave(cnt).to(torch.int32).contiguous()
            out = torch.empty((n_ct * cnt, glwe), dtype=torch.int64, device=dev)
            _lib.call("fhe_cmux_batch", out.data_ptr(), cur_pool.data_ptr(), i0.data_ptr(), i1.data_ptr(),
                      ggsw_f.data_ptr(), gi.data_ptr(), n_ct * cnt, k, N, p.cbs_level, p.cbs_base_log,
                      1 if lvl == 0 else 0, E._stream())      # the leaves are plaintext polynomials
            cur_pool = out
            cur_idx = torch.arange(n_ct * cnt, dtype=torch.int64, device=dev).reshape(n_ct, cnt)
        acc_idx = cur_idx[:, 0].to(torch.int32).contiguous()
        key = ("rot", nrot)
        if key not in ev.wide_cache:
            ev.wide_cache[key] = E.from_np_u64(_rotation_masks(N, nrot), dev)
        lwe = ev.wide_cache[key][None, :].expand(n_ct, nrot + 1).contiguous()
        lib = _lib.load()
        stride = bits * lib.fhe_bsk_fourier_elems(1, k, N, p.cbs_level)
        res = torch.empty((n_ct, p.ct_words), dtype=torch.int64, device=dev)
        _lib.call("fhe_blind_rotate_batch", res.data_ptr(), lwe.data_ptr(), n_ct, ggsw_f.data_ptr(), stride,
                  cur_pool.data_ptr(), acc_idx.data_ptr(), nrot, k, N, p.cbs_level, p.cbs_base_log, E._stream())
    return res


PROFILE = None            # development aid: set to a dict to accumulate per-phase milliseconds (synchronises)
_last_tick = [None]


def _tick(name: str) -> None:
    if PROFILE is None:
        return
    import time

    torch.cuda.synchronize()
    now = time.perf_counter()
    if name != "start" and _last_tick[0] is not None:
        PROFILE[name] = PROFILE.get(name, 0.0) + (now - _last_tick[0]) * 1e3
    _last_tick[0] = now


WIDE_CHUNK = 512          # ciphertexts per pass: the GGSWs of a chunk take bits * cbs_level * (k+1)^2 * N * 16 B each


def wide_table_lookup(ev, cts: Sequence[torch.Tensor], pool: torch.Tensor, polys_per_residue: int,
                      table_idx: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """cts: one [E, kN+1] tensor per residue (moduli ev.params.moduli).  pool: lut_pool(...) of the op's tables;
    table_idx [E] selects the table per element (mapped lookup) or None.  Returns one tensor per output residue."""
    p = ev.params
    moduli = p.moduli
    r = len(moduli)
    n_all = cts[0].shape[0]
    bits = total_bits(moduli)
    outs = [torch.empty((n_all, p.ct_words), dtype=torch.int64, device=cts[0].device) for _ in range(r)]
    for lo in range(0, n_all, WIDE_CHUNK):
        hi = min(n_all, lo + WIDE_CHUNK)
        n_ct = hi - lo
        parts = []
        for ct in cts:
            part = ct[lo:hi]
            if not part.is_contiguous() or part.data_ptr() % 16:       # the linear kernels vectorise on 16 bytes
                part = part.clone()
            parts.append(part)
        _tick("start")
        levels = [lv for per_residue in extract_and_bootstrap(ev, parts, moduli)[1] for lv in per_residue]
        stacked = torch.stack(levels, dim=1).reshape(n_ct * bits, p.cbs_level, p.ct_words)      # [e][t][lv]
        _tick("extract+bootstrap")
        ggsw_f = circuit_bootstrap(ev, stacked)
        _tick("packing+fourier")
        tsel = torch.zeros(n_ct, dtype=torch.int64, device=stacked.device) if table_idx is None else table_idx[lo:hi].to(torch.int64)
        for o in range(r):
            first = (tsel * r + o) * polys_per_residue
            outs[o][lo:hi] = vertical_packing(ev, ggsw_f, n_ct, bits, pool, first)
        _tick("vertical_packing")
    return outs
