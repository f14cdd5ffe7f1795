"""
Device engine: thin Python layer over the C ABI.  PyTorch is used only for device buffers
(`torch.empty(..., dtype=torch.int64, device=...)`, `.data_ptr()`) and for the current stream.

Ciphertext tensors are `torch.int64` of shape `[E, k*N+1]` (bit pattern of the u64 torus words,
mask first, body last) -- SURVEY.md Appendix A.1.
"""

import os
from dataclasses import asdict, dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

DOM_SK_SMALL, DOM_SK_BIG = 1, 2


@dataclass(frozen=True)
class CryptoParams:
    """One TFHE parameter set (the reference uses a single set per circuit because
    `_update_bit_widths` forces one width on every encrypted value,
    concrete/numpy/mlir/graph_converter.py:315-322)."""

    p: int                 # message bits carried by every ciphertext (one extra padding bit)
    n: int                 # small LWE dimension
    k: int                 # GLWE dimension
    N: int                 # polynomial size
    pbs_level: int
    pbs_base_log: int
    ks_level: int
    ks_base_log: int
    lwe_std: float         # noise std of the small key (relative to q)
    glwe_std: float        # noise std of the big key
    # wide integers (p > 8, wide.py): CRT moduli of the residues every value is split into, circuit-bootstrap and
    # packing-keyswitch decompositions; empty / zero for the native one-ciphertext-per-value encoding
    moduli: tuple = ()
    cbs_level: int = 0
    cbs_base_log: int = 0
    pf_level: int = 0
    pf_base_log: int = 0

    def __post_init__(self):
        object.__setattr__(self, "moduli", tuple(int(m) for m in self.moduli))

    @property
    def wide(self) -> bool:
        return len(self.moduli) > 0

    @property
    def linear_only(self) -> bool:
        """Circuits without table lookups carry no bootstrap / keyswitch keys (params.select_linear_parameters)."""
        return self.pbs_level == 0

    @property
    def big_dim(self) -> int:
        return self.k * self.N

    @property
    def ct_words(self) -> int:
        return self.k * self.N + 1

    def to_dict(self):
        return asdict(self)


def fresh_key() -> bytes:
    """64 bytes of OS entropy: a 256-bit secret ChaCha key and a separate 256-bit key for the public masks."""
    return os.urandom(_lib.KEY_BYTES)


def derive_key(seed, tweak: int):
    """Independent key material for another object under the same root: a TEST seed (int) is xored with the tweak
    (the oracle does the same), real key material is re-hashed with it."""
    if isinstance(seed, (bytes, bytearray)):
        import hashlib

        return hashlib.sha512(bytes(seed) + int(tweak).to_bytes(8, "little")).digest()
    return (int(seed) ^ int(tweak)) & ((1 << 64) - 1)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def to_np_u64(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)


def from_np_u64(a: np.ndarray, device) -> torch.Tensor:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).to(device, non_blocking=False)


class KeySet:
    """Device-resident secret + evaluation keys for one parameter set."""

    def __init__(self, params: CryptoParams, device, seed=None, keep_std_bsk: bool = False):
        """`seed`: None = 64 bytes of OS entropy (production); 64 bytes = that key material; an int = a TEST seed
        (deterministic expansion, bit-reproducible against the oracle and identical on every rank)."""
        self.params = params
        self.device = torch.device(device)
        self.seed = fresh_key() if seed is None else (bytes(seed) if isinstance(seed, (bytes, bytearray)) else int(seed))
        self.deterministic = isinstance(self.seed, int)
        key = _lib.key_material(self.seed)
        p = params
        dev = self.device
        with torch.cuda.device(dev):
            st = _stream()
            self.sk_big = torch.empty(p.big_dim, dtype=torch.int64, device=dev)
            _lib.call("fhe_keygen_lwe_secret", self.sk_big.data_ptr(), p.big_dim, key, DOM_SK_BIG, st)
            if p.linear_only:           # no table lookups: nothing to bootstrap or keyswitch with
                self.sk_small = torch.empty(0, dtype=torch.int64, device=dev)
                self.ksk = torch.empty((0, 0, 1), dtype=torch.int64, device=dev)
                self.bsk_std = None
                self.bsk_f = torch.empty((0, 2), dtype=torch.float64, device=dev)
                self.pfksk = None
                torch.cuda.current_stream().synchronize()
                return
            self.sk_small = torch.empty(p.n, dtype=torch.int64, device=dev)
            _lib.call("fhe_keygen_lwe_secret", self.sk_small.data_ptr(), p.n, key, DOM_SK_SMALL, st)
            self.ksk = torch.empty((p.big_dim, p.ks_level, p.n + 1), dtype=torch.int64, device=dev)
            _lib.call("fhe_keygen_ksk", self.ksk.data_ptr(), self.sk_big.data_ptr(), p.big_dim,
                      self.sk_small.data_ptr(), p.n, p.ks_level, p.ks_base_log, p.lwe_std, key, st)
            bsk_std = torch.empty((p.n, p.pbs_level, p.k + 1, p.k + 1, p.N), dtype=torch.int64, device=dev)
            scratch = torch.empty(p.k * p.N + p.k + 8, dtype=torch.int32, device=dev)
            _lib.call("fhe_keygen_bsk", bsk_std.data_ptr(), self.sk_small.data_ptr(), p.n,
                      self.sk_big.data_ptr(), p.k, p.N, p.pbs_level, p.pbs_base_log, p.glwe_std,
                      key, scratch.data_ptr(), st)
            self.bsk_std = bsk_std if keep_std_bsk else None
            self.bsk_f = bsk_std_to_fourier(bsk_std, p)
            self.pfksk = None
            if p.wide:
                from . import wide

                self.pfksk = wide.keygen_packing_key(p, self.sk_big, self.seed)
            torch.cuda.current_stream().synchronize()

    # sizes reported through CompilationFeedback (reference server.py:296-350)
    def secret_keys_bytes(self) -> int:
        return 8 * (self.params.n + self.params.big_dim)

    def bootstrap_keys_bytes(self) -> int:
        p = self.params
        return 8 * p.n * p.pbs_level * (p.k + 1) * (p.k + 1) * p.N

    def keyswitch_keys_bytes(self) -> int:
        p = self.params
        return 8 * p.big_dim * p.ks_level * (p.n + 1)


def bsk_std_to_fourier(bsk_std: torch.Tensor, p: CryptoParams) -> torch.Tensor:
    """Standard-domain BSK -> Fourier BSK on device (layout owned by csrc/pbs.cu)."""
    lib = _lib.load()
    if lib.fhe_pbs_cluster_size(p.N, p.k, p.pbs_level) < 0:
        raise _lib.FheError(f"PBS layout unsupported for N={p.N} k={p.k} level={p.pbs_level}")
    elems = lib.fhe_bsk_fourier_elems(p.n, p.k, p.N, p.pbs_level)
    out = torch.empty((elems, 2), dtype=torch.float64, device=bsk_std.device)
    with torch.cuda.device(bsk_std.device):
        _lib.call("fhe_bsk_to_fourier", out.data_ptr(), bsk_std.data_ptr(), p.n, p.k, p.N,
                  p.pbs_level, p.pbs_base_log, _stream())
    return out


class EvalKeys:
    """Evaluation keys only (what the server sees): Fourier BSK + KSK.

    The bootstrap has several plan variants per parameter set (csrc/pbs.cu): variant 0 maximises the
    ciphertexts in flight, higher variants use larger clusters per ciphertext.  A small batch (a step of a
    sequential circuit) is run on the largest cluster size that still bootstraps the whole batch at once;
    the Fourier BSK is permuted into that variant's layout on first use and cached."""

    def __init__(self, params: CryptoParams, bsk_f: torch.Tensor, ksk: torch.Tensor,
                 pfksk: Optional[torch.Tensor] = None):
        self.params = params
        self.bsk_f = bsk_f
        self.ksk = ksk
        self.pfksk = pfksk
        self.packing = None         # packing keyswitch of the wide-integer lookup (wide.py)
        self.wide_cache = {}        # small device constants of that lookup
        if pfksk is not None:
            from . import wide

            self.packing = wide.PackingKey(params, pfksk)
        self._variants = {0: bsk_f}
        lib = _lib.load()
        n = 0 if params.linear_only else lib.fhe_pbs_num_variants(params.N, params.k, params.pbs_level)
        self.variant_clusters = [lib.fhe_pbs_variant_cluster_size(params.N, params.k, params.pbs_level, v)
                                 for v in range(max(n, 0))]
        self._capacity = None      # ciphertexts each variant bootstraps in one wave (queried on the device)
        self._measured_best = {}   # batch size -> fastest variant, measured on first use (_calibrate)
        # tensor-core keyswitch (csrc/keyswitch_tc.cu): key bytes re-laid out once, on first use
        self.ks_tc_ok = (not params.linear_only) and bool(
            lib.fhe_ks_tc_supported(params.big_dim, params.n, params.ks_level, params.ks_base_log))
        self._ksk_tc = None

    def ksk_tc(self) -> torch.Tensor:
        if self._ksk_tc is None:
            p = self.params
            lib = _lib.load()
            buf = torch.empty(lib.fhe_ksk_tc_bytes(p.big_dim, p.n, p.ks_level), dtype=torch.uint8, device=self.ksk.device)
            with torch.cuda.device(self.ksk.device):
                _lib.call("fhe_ksk_to_tc", buf.data_ptr(), self.ksk.data_ptr(), p.big_dim, p.n, p.ks_level, _stream())
            self._ksk_tc = buf
        return self._ksk_tc

    def variant_for_batch(self, count: int) -> int:
        """Plan variant for a batch of `count` ciphertexts.  More than one wave of the throughput plan: variant 0.  A
        single wave: the variant measured fastest at this batch size (_calibrate, default).  With CNP_B200_VARIANT_CALIB=0 a
        static rule instead: the largest cluster that still bootstraps the batch in a single wave -- except that the grouped
        plan wins at its size and a PIPELINED variant (k_pbs_pipe) that holds the batch and is at most half that cluster is
        faster (scripts/variant_latency.py: N = 4096, 17 ciphertexts: 4.4 ms on 4 pipelined CTAs vs 5.25 ms on 8)."""
        if count > 0 and count in self._measured_best:
            return self._measured_best[count]
        if self._capacity is None:
            lib = _lib.load()
            p = self.params
            with torch.cuda.device(self.bsk_f.device):
                self._capacity = [lib.fhe_pbs_variant_capacity(p.N, p.k, p.pbs_level, v)
                                  for v in range(len(self.variant_clusters))]
            self._pipelined = [p.k == 1 and p.pbs_level == 2 and 2 * p.pbs_base_log <= 30 and
                               lib.fhe_pbs_variant_pipelined(p.N, p.k, p.pbs_level, v) == 1
                               for v in range(len(self.variant_clusters))]
            self._grouped0 = lib.fhe_pbs_variant_grouped(p.N, p.k, p.pbs_level, 0) == 1
        if count > self._capacity[0]:
            return 0              # several waves of the throughput plan: nothing else competes
        if count > 0 and len(self.variant_clusters) > 1 and os.environ.get("CNP_B200_VARIANT_CALIB", "1") not in ("", "0"):
            return self._calibrate(count)
        # static rule (CNP_B200_VARIANT_CALIB=0):
        # the grouped plan (k_pbs_group) is also the latency plan of its size: a lone ciphertext per SM has the SM and its
        # tensor memory to itself (measured at N = 2048, scripts/variant_latency.py: 3.13 ms on one SM, 4.6 ms on a
        # 2-CTA cluster, 3.96 ms on 4) -- as long as the batch leaves one slot per SM
        if self._grouped0 and 0 < count <= self._capacity[0] // 2:
            return 0
        best = 0
        for v, cap in enumerate(self._capacity):
            # 16-CTA clusters are a capacity fallback (N = 32768 with >= 3 levels), not a latency win
            if v > 0 and 0 < count <= cap and self.variant_clusters[v] <= 8:
                best = v
        for v, cap in enumerate(self._capacity):
            if (v > 0 and v != best and self._pipelined[v] and 0 < count <= cap
                    and 2 * self.variant_clusters[v] >= self.variant_clusters[best] > self.variant_clusters[v]):
                return v
        return best

    def _calibrate(self, count: int) -> int:
        """Fastest plan variant for a single-wave batch of `count` ciphertexts, MEASURED the first time this batch size is
        seen (one warm-up and one timed launch per variant on dummy ciphertexts, CUDA events; cached per size).  The wave
        time of a variant depends on how its clusters share SMs, so no closed form picks reliably: at N = 4096, 60
        ciphertexts take 9.1 ms on 1 CTA each, 11.0 on 2, 8.9 on 4 (two waves) and 7.0 on 8 (scripts/variant_latency.py).
        A circuit has a handful of distinct batch sizes; each costs a few tens of milliseconds once per key set."""
        p = self.params
        dev = self.bsk_f.device
        # random mask words: the rotation amounts decide which CTA every accumulator word is pushed to (all-zero inputs would
        # keep the cluster exchange of the pipelined plans local and flatter them)
        gen = torch.Generator(device=dev)
        gen.manual_seed(0x5eed)
        small = torch.randint(-(1 << 62), 1 << 62, (count, p.n + 1), dtype=torch.int64, device=dev, generator=gen)
        luts = torch.zeros((1, (p.k + 1) * p.N), dtype=torch.int64, device=dev)
        out = torch.empty((count, p.ct_words), dtype=torch.int64, device=dev)
        best, best_ms = 0, float("inf")
        with torch.cuda.device(dev):
            for v, c in enumerate(self.variant_clusters):
                if c > 8 and v > 0:
                    continue          # 16-CTA clusters are a capacity fallback, not a latency plan
                bsk = self.bsk_for_variant(v).data_ptr()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for timed in (False, True):
                    if timed:
                        e0.record()
                    _lib.call("fhe_pbs_batch_variant", out.data_ptr(), small.data_ptr(), count, bsk, luts.data_ptr(), None,
                              p.n, p.k, p.N, p.pbs_level, p.pbs_base_log, v, _stream())
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1)
                if ms < 0.97 * best_ms:          # a later (larger) cluster must win clearly
                    best, best_ms = v, ms
        self._measured_best[count] = best
        return best

    def bsk_for_variant(self, v: int) -> torch.Tensor:
        if v not in self._variants:
            p = self.params
            out = torch.empty_like(self.bsk_f)
            with torch.cuda.device(self.bsk_f.device):
                _lib.call("fhe_bsk_relayout", out.data_ptr(), self.bsk_f.data_ptr(), p.n, p.k, p.N, p.pbs_level,
                          0, v, _stream())
            self._variants[v] = out
        return self._variants[v]

    @staticmethod
    def of(keyset: KeySet) -> "EvalKeys":
        return EvalKeys(keyset.params, keyset.bsk_f, keyset.ksk, keyset.pfksk)


# ------------------------------------------------------------------ client side ----

def _enc_key(keyset: KeySet, seed):
    """Mask / noise key material of one encryption call: FRESH OS entropy unless the caller asks for reproducible
    streams (a TEST seed, or a deterministic key set with an explicit index window)."""
    if seed is None:
        seed = fresh_key()
    return _lib.key_material(seed)


def encrypt(keyset: KeySet, messages: np.ndarray, first_index: int = 0, seed=None) -> torch.Tensor:
    """Encode `m << (63 - p)` and LWE-encrypt each message under the big key.  `seed` None: fresh randomness."""
    p = keyset.params
    dev = keyset.device
    m = np.ascontiguousarray(np.asarray(messages).reshape(-1).astype(np.uint64))
    pt = from_np_u64(m << np.uint64(63 - p.p), dev)
    ct = torch.empty((m.size, p.ct_words), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fhe_encrypt_lwe_batch", ct.data_ptr(), pt.data_ptr(), m.size, keyset.sk_big.data_ptr(),
                  p.big_dim, p.glwe_std, _enc_key(keyset, seed), int(first_index), _stream())
    return ct


def encrypt_plaintexts(keyset: KeySet, plaintexts: np.ndarray, first_index: int = 0, seed=None) -> torch.Tensor:
    """LWE-encrypt already encoded u64 plaintexts under the big key (CRT residues of the wide path)."""
    p = keyset.params
    dev = keyset.device
    pt_np = np.ascontiguousarray(np.asarray(plaintexts).reshape(-1).astype(np.uint64))
    pt = from_np_u64(pt_np, dev)
    ct = torch.empty((pt_np.size, p.ct_words), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("fhe_encrypt_lwe_batch", ct.data_ptr(), pt.data_ptr(), pt_np.size, keyset.sk_big.data_ptr(),
                  p.big_dim, p.glwe_std, _enc_key(keyset, seed), int(first_index), _stream())
    return ct


def decrypt(keyset: KeySet, ct: torch.Tensor, return_phase: bool = False):
    """Decrypt `[E, kN+1]` ciphertexts to messages of p+1 bits (reference client.py:201 + A.2)."""
    p = keyset.params
    ct = ct.contiguous()
    E = ct.numel() // p.ct_words
    msg = torch.empty(E, dtype=torch.int64, device=ct.device)
    phase = torch.empty(E, dtype=torch.int64, device=ct.device) if return_phase else None
    with torch.cuda.device(ct.device):
        _lib.call("fhe_decrypt_lwe_batch", _ptr(phase), msg.data_ptr(), ct.data_ptr(), E,
                  keyset.sk_big.data_ptr(), p.big_dim, p.p, _stream())
    if return_phase:
        return to_np_u64(msg), to_np_u64(phase)
    return to_np_u64(msg)


def decrypt_small(keyset: KeySet, ct: torch.Tensor) -> np.ndarray:
    """Phase of ciphertexts under the small key (test / noise measurement helper)."""
    p = keyset.params
    ct = ct.contiguous()
    E = ct.numel() // (p.n + 1)
    phase = torch.empty(E, dtype=torch.int64, device=ct.device)
    with torch.cuda.device(ct.device):
        _lib.call("fhe_decrypt_lwe_batch", phase.data_ptr(), None, ct.data_ptr(), E,
                  keyset.sk_small.data_ptr(), p.n, p.p, _stream())
    return to_np_u64(phase)


def trivial(params: CryptoParams, messages: np.ndarray, device) -> torch.Tensor:
    m = np.ascontiguousarray(np.asarray(messages).reshape(-1).astype(np.uint64))
    pt = from_np_u64(m << np.uint64(63 - params.p), device)
    ct = torch.empty((m.size, params.ct_words), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        _lib.call("fhe_trivial_lwe_batch", ct.data_ptr(), pt.data_ptr(), m.size, params.big_dim, _stream())
    return ct


# ------------------------------------------------------------------ server side ----

def build_lut_accumulators(params: CryptoParams, tables: np.ndarray) -> np.ndarray:
    """Expand `[T, 2^p]` integer tables into `[T, (k+1)*N]` GLWE accumulators:
    every entry repeated N/2^p times, pre-rotated by half a box with negacyclic sign, outputs
    encoded at bit 63-p.  Host-side (tables are compile-time constants of a circuit)."""
    p, N, k = params.p, params.N, params.k
    tables = np.asarray(tables)
    T = tables.shape[0]
    assert tables.shape[1] == (1 << p), (tables.shape, p)
    box = N >> p
    if box < 1:
        raise ValueError(f"polynomial size {N} too small for {p}-bit tables")
    vals = (tables.astype(np.int64).view(np.uint64)) << np.uint64(64 - (p + 1))
    poly = np.repeat(vals, box, axis=1)                       # [T, N]
    half = box // 2
    rolled = np.concatenate([poly[:, half:], (np.uint64(0) - poly[:, :half])], axis=1)
    out = np.zeros((T, (k + 1) * N), dtype=np.uint64)
    out[:, k * N:] = rolled
    return out


def keyswitch_launch(ev: EvalKeys, out: torch.Tensor, ct: torch.Tensor, count: int) -> None:
    """Keyswitch `count` ciphertexts of `ct` into `out` on the current stream: int8-limb GEMM on the tensor
    cores when the parameters allow it exactly (bit-identical results), else the CUDA-core kernel."""
    p = ev.params
    if count <= 0:
        return
    if ev.ks_tc_ok and os.environ.get("CNP_B200_KS_TC", "1") not in ("", "0"):
        lib = _lib.load()
        ksk_tc = ev.ksk_tc()
        # the digit matrix of a chunk (rows x level*kN bytes) is kept under ~1 GiB
        rows = max(128, ((1 << 30) // max(lib.fhe_keyswitch_tc_scratch_bytes(128, p.big_dim, p.ks_level), 1)) * 128)
        out2 = out.view(-1, p.n + 1)
        ct2 = ct.view(-1, p.ct_words)
        for lo in range(0, count, rows):
            hi = min(count, lo + rows)
            scratch = torch.empty(lib.fhe_keyswitch_tc_scratch_bytes(hi - lo, p.big_dim, p.ks_level), dtype=torch.uint8,
                                  device=ct.device)
            _lib.call("fhe_keyswitch_batch_tc", out2[lo:hi].data_ptr(), ct2[lo:hi].data_ptr(), hi - lo, ksk_tc.data_ptr(),
                      scratch.data_ptr(), p.big_dim, p.n, p.ks_level, p.ks_base_log, _stream())
    else:
        _lib.call("fhe_keyswitch_batch", out.data_ptr(), ct.data_ptr(), count, ev.ksk.data_ptr(), p.big_dim,
                  p.n, p.ks_level, p.ks_base_log, _stream())


def keyswitch(ev: EvalKeys, ct: torch.Tensor) -> torch.Tensor:
    p = ev.params
    ct = ct.contiguous()
    E = ct.numel() // p.ct_words
    out = torch.empty((E, p.n + 1), dtype=torch.int64, device=ct.device)
    with torch.cuda.device(ct.device):
        keyswitch_launch(ev, out, ct, E)
    return out


def pbs_launch(ev: EvalKeys, out: torch.Tensor, ct_small: torch.Tensor, count: int, luts: torch.Tensor,
               lut_idx: Optional[torch.Tensor], variant: Optional[int] = None) -> None:
    """One bootstrap launch on the current stream: `count` ciphertexts from `ct_small` into `out`."""
    p = ev.params
    v = ev.variant_for_batch(count) if variant is None else int(variant)
    _lib.call("fhe_pbs_batch_variant", out.data_ptr(), ct_small.data_ptr(), count, ev.bsk_for_variant(v).data_ptr(),
              luts.data_ptr(), _ptr(lut_idx), p.n, p.k, p.N, p.pbs_level, p.pbs_base_log, v, _stream())


def bootstrap(ev: EvalKeys, ct_small: torch.Tensor, luts: torch.Tensor, lut_idx: Optional[torch.Tensor],
              variant: Optional[int] = None) -> torch.Tensor:
    p = ev.params
    E = ct_small.numel() // (p.n + 1)
    out = torch.empty((E, p.ct_words), dtype=torch.int64, device=ct_small.device)
    with torch.cuda.device(ct_small.device):
        if variant is None:
            pbs_launch(ev, out, ct_small, E, luts, lut_idx)
        else:
            pbs_launch(ev, out, ct_small, E, luts, lut_idx, variant)
    return out


_SIDE_STREAMS = {}
# Two-stream keyswitch/bootstrap pipelining is opt-in (CNP_B200_PIPELINE=1): measured on B200 it hides the
# keyswitch (share of the bootstrap in a step 0.93 -> 0.96) but the concurrent keyswitch slows the bootstrap
# by about as much (L2/HBM contention) and chunking adds partial waves, so the net effect is within the
# box-to-box noise (+0.7 % / -2 % in two runs).
PIPELINE_MIN_CHUNK = 480          # ciphertexts; below 2 chunks of this size the lookup is not pipelined
PIPELINE_MAX_CHUNKS = 4 if os.environ.get("CNP_B200_PIPELINE", "0") not in ("", "0") else 1


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def table_lookup(ev: EvalKeys, ct: torch.Tensor, luts: torch.Tensor, lut_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FHE.apply_lookup_table = keyswitch + programmable bootstrap per element.

    Large batches are cut into chunks and software-pipelined over two streams: the bootstrap of chunk i
    (thread-block clusters, which cannot cover every SM of the chip: 15 clusters of 8 on a B200) runs
    on the current stream while the keyswitch of chunk i+1 runs on a side stream on the SMs the clusters
    leave idle.  Results are identical to the unchunked call (ciphertexts are independent)."""
    p = ev.params
    ct = ct.contiguous()
    E = ct.numel() // p.ct_words
    n_chunks = min(PIPELINE_MAX_CHUNKS, E // PIPELINE_MIN_CHUNK)
    if n_chunks < 2:
        return bootstrap(ev, keyswitch(ev, ct), luts, lut_idx)
    ct = ct.view(E, p.ct_words)
    dev = ct.device
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    bounds = [(E * i) // n_chunks for i in range(n_chunks + 1)]
    small = torch.empty((E, p.n + 1), dtype=torch.int64, device=dev)
    out = torch.empty((E, p.ct_words), dtype=torch.int64, device=dev)

    def ks(i):
        lo, hi = bounds[i], bounds[i + 1]
        keyswitch_launch(ev, small[lo:hi], ct[lo:hi], hi - lo)

    def pbs(i):
        lo, hi = bounds[i], bounds[i + 1]
        idx = None if lut_idx is None else lut_idx[lo:hi]
        pbs_launch(ev, out[lo:hi], small[lo:hi], hi - lo, luts, idx)

    with torch.cuda.device(dev):
        ready = torch.cuda.Event()
        ready.record(main)                   # inputs (and the buffers above) are ready on the main stream
        side.wait_event(ready)
        ks(0)
        for i in range(n_chunks):
            pbs(i)                           # enqueued first: the clusters take their SMs ...
            if i + 1 < n_chunks:
                with torch.cuda.stream(side):
                    ks(i + 1)                # ... and the next keyswitch fills the SMs they cannot use
                    done = torch.cuda.Event()
                    done.record(side)
                main.wait_event(done)
    return out


def _rows(t: torch.Tensor, L: int) -> int:
    return t.numel() // L


def lin_add(a, b, L, ia=None, ib=None, E=None):
    E = _rows(a, L) if E is None else E
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_add", out.data_ptr(), a.data_ptr(), b.data_ptr(), _ptr(ia), _ptr(ib), E, L, _stream())
    return out


def lin_sub(a, b, L, ia=None, ib=None, E=None):
    E = _rows(a, L) if E is None else E
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_sub", out.data_ptr(), a.data_ptr(), b.data_ptr(), _ptr(ia), _ptr(ib), E, L, _stream())
    return out


def lin_neg(a, L, ia=None, E=None):
    E = _rows(a, L) if E is None else E
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_neg", out.data_ptr(), a.data_ptr(), _ptr(ia), E, L, _stream())
    return out


def lin_add_plain(a, pt, L, ia=None):
    """pt: int64 tensor [E] of already-encoded plaintexts."""
    E = pt.numel()
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_add_plain", out.data_ptr(), a.data_ptr(), _ptr(ia), pt.data_ptr(), E, L, _stream())
    return out


def lin_plain_sub(a, pt, L, ia=None):
    E = pt.numel()
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_plain_sub", out.data_ptr(), a.data_ptr(), _ptr(ia), pt.data_ptr(), E, L, _stream())
    return out


def lin_mul_clear(a, c, L, ia=None):
    """c: int64 tensor [E] of clear signed multipliers."""
    E = c.numel()
    out = torch.empty((E, L), dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("fhe_lin_mul_clear", out.data_ptr(), a.data_ptr(), _ptr(ia), c.data_ptr(), E, L, _stream())
    return out


def lin_gather_mac(x, idx, w, L, bias=None):
    """out[e] = sum_j w[e, j] * x[idx[e, j]] (+ bias[e] on the body); idx int32 [E, K], w int64 [E, K]."""
    E, K = idx.shape
    out = torch.empty((E, L), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("fhe_lin_gather_mac", out.data_ptr(), x.data_ptr(), idx.data_ptr(), w.data_ptr(), _ptr(bias),
                  E, K, L, _stream())
    return out


def matmul_ct_clear(x, W, Mr, K, Nc, L):
    """x: [Mr, K, L] ciphertexts, W: int64 [K, Nc] clear weights -> [Mr, Nc, L]."""
    out = torch.empty((Mr * Nc, L), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("fhe_matmul_ct_clear", out.data_ptr(), x.data_ptr(), W.data_ptr(), Mr, K, Nc, L, _stream())
    return out


def measure_fp64_peak(device) -> float:
    import ctypes as C

    v = C.c_double(0.0)
    with torch.cuda.device(device):
        _lib.call("fhe_measure_fp64_peak", C.byref(v), _stream())
    return float(v.value)
