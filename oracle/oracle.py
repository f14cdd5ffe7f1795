"""
ctypes front of the CPU oracle (oracle/tfhe_ref.c).

TEST INFRASTRUCTURE ONLY -- see the header of tfhe_ref.c.  Only tests/,
`__graft_entry__.smoke()` and bench.py's CPU legs may import this module.

PARITY UNPINNED at ciphertext level (no reference golden ciphertexts exist, the
arithmetic lives in the absent `concrete-compiler==0.24.0rc5` wheel); pinned at the
decrypted-output level by tests/golden/ (generated from the reference frontend).
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(native: bool = False) -> str:
    """Compile tfhe_ref.c (portable x86-64-v3 by default, -march=native on request)."""
    target = "native" if native else "liboracle.so"
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "_native" if native else "", "liboracle.so")


def _load(native: bool = False):
    path = os.path.join(_HERE, "_native" if native else "", "liboracle.so")
    src = os.path.join(_HERE, "tfhe_ref.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build(native)
    return C.CDLL(path)


class Oracle:
    """Thin numpy wrapper over the C restatement."""

    def __init__(self, native: bool = False):
        self.lib = _load(native)
        L = self.lib
        for name in [n for n in dir(self) if n.startswith("_sig_")]:
            getattr(self, name)(L)

    # -- helpers
    @staticmethod
    def _u64(a):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return a, a.ctypes.data_as(_u64p)

    # -- rng
    def rng_u64(self, seed, stream, idx0, count):
        out = np.empty(count, dtype=np.uint64)
        self.lib.ref_rng_u64(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(idx0),
                             C.c_uint64(count), out.ctypes.data_as(_u64p))
        return out

    def gaussian(self, seed, stream, j0, count, std):
        out = np.empty(count, dtype=np.uint64)
        self.lib.ref_gaussian(C.c_uint64(seed), C.c_uint64(stream), C.c_uint64(j0),
                              C.c_uint64(count), C.c_double(std), out.ctypes.data_as(_u64p))
        return out

    # -- key material: 16 words (secret ChaCha key, public-mask key); an int is a TEST seed expanded deterministically
    def key_material(self, seed):
        key = (C.c_uint32 * 16)()
        if isinstance(seed, (bytes, bytearray)):
            if len(seed) != 64:
                raise ValueError("key material is 64 bytes (secret key, public-mask key)")
            C.memmove(key, bytes(seed), 64)
        else:
            self.lib.ref_seed_expand(C.c_uint64(int(seed) & ((1 << 64) - 1)), key)
        return key

    # -- keys
    def keygen_lwe_secret(self, dim, seed, domain):
        sk = np.empty(dim, dtype=np.uint64)
        self.lib.ref_keygen_lwe_secret(sk.ctypes.data_as(_u64p), C.c_int(dim), self.key_material(seed),
                                       C.c_int(domain))
        return sk

    def keygen_bsk(self, sk_small, sk_glwe, k, N, level, base_log, std, seed):
        n = len(sk_small)
        bsk = np.empty((n, level, k + 1, k + 1, N), dtype=np.uint64)
        s1, p1 = self._u64(sk_small)
        s2, p2 = self._u64(sk_glwe)
        self.lib.ref_keygen_bsk(bsk.ctypes.data_as(_u64p), p1, C.c_int(n), p2, C.c_int(k),
                                C.c_int(N), C.c_int(level), C.c_int(base_log), C.c_double(std),
                                self.key_material(seed))
        return bsk

    def keygen_ksk(self, sk_big, sk_small, level, base_log, std, seed):
        kN, n = len(sk_big), len(sk_small)
        ksk = np.empty((kN, level, n + 1), dtype=np.uint64)
        s1, p1 = self._u64(sk_big)
        s2, p2 = self._u64(sk_small)
        self.lib.ref_keygen_ksk(ksk.ctypes.data_as(_u64p), p1, C.c_int(kN), p2, C.c_int(n),
                                C.c_int(level), C.c_int(base_log), C.c_double(std),
                                self.key_material(seed))
        return ksk

    # -- encrypt / decrypt
    def encrypt(self, pt, sk, std, seed, first_index=0):
        pt, pp = self._u64(np.asarray(pt).reshape(-1))
        sk, ps = self._u64(sk)
        dim = len(sk)
        ct = np.empty((len(pt), dim + 1), dtype=np.uint64)
        self.lib.ref_encrypt_lwe_batch(ct.ctypes.data_as(_u64p), pp, C.c_long(len(pt)), ps,
                                       C.c_int(dim), C.c_double(std), self.key_material(seed),
                                       C.c_uint64(first_index))
        return ct

    def trivial(self, pt, dim):
        pt, pp = self._u64(np.asarray(pt).reshape(-1))
        ct = np.empty((len(pt), dim + 1), dtype=np.uint64)
        self.lib.ref_trivial_lwe_batch(ct.ctypes.data_as(_u64p), pp, C.c_long(len(pt)),
                                       C.c_int(dim))
        return ct

    def decrypt_phase(self, ct, sk):
        ct, pc = self._u64(ct)
        sk, ps = self._u64(sk)
        dim = len(sk)
        B = ct.size // (dim + 1)
        ph = np.empty(B, dtype=np.uint64)
        self.lib.ref_decrypt_lwe_batch(ph.ctypes.data_as(_u64p), pc, C.c_long(B), ps,
                                       C.c_int(dim))
        return ph

    def decode(self, phase, p):
        phase, pp = self._u64(phase)
        out = np.empty(phase.size, dtype=np.uint64)
        self.lib.ref_decode_batch(out.ctypes.data_as(_u64p), pp, C.c_long(phase.size), C.c_int(p))
        return out

    def signed_decompose(self, a, base_log, level):
        a, pa = self._u64(a)
        out = np.empty((a.size, level), dtype=np.int64)
        self.lib.ref_signed_decompose(pa, C.c_long(a.size), C.c_int(base_log), C.c_int(level),
                                      out.ctypes.data_as(_i64p))
        return out

    # -- keyswitch / bootstrap
    def keyswitch(self, ct, ksk, level, base_log):
        kN, _, n1 = ksk.shape
        ct, pc = self._u64(ct)
        ksk, pk = self._u64(ksk)
        B = ct.size // (kN + 1)
        out = np.empty((B, n1), dtype=np.uint64)
        self.lib.ref_keyswitch_batch(out.ctypes.data_as(_u64p), pc, C.c_long(B), pk, C.c_int(kN),
                                     C.c_int(n1 - 1), C.c_int(level), C.c_int(base_log))
        return out

    def bsk_to_fourier(self, bsk):
        n, level, k1, _, N = bsk.shape
        bsk, pb = self._u64(bsk)
        out = np.empty((n, level, k1, k1, N // 2, 2), dtype=np.float64)
        self.lib.ref_bsk_to_fourier(out.ctypes.data_as(_f64p), pb, C.c_int(n), C.c_int(k1 - 1),
                                    C.c_int(N), C.c_int(level))
        return out

    def pbs(self, ct, bsk_f, luts, lut_idx, base_log):
        n, level, k1, _, M, _ = bsk_f.shape
        k, N = k1 - 1, 2 * M
        ct, pc = self._u64(ct)
        luts, pl = self._u64(luts)
        B = ct.size // (n + 1)
        out = np.empty((B, k * N + 1), dtype=np.uint64)
        if lut_idx is None:
            pi = None
        else:
            lut_idx = np.ascontiguousarray(lut_idx, dtype=np.int32)
            pi = lut_idx.ctypes.data_as(_i32p)
        bsk_f = np.ascontiguousarray(bsk_f)
        self.lib.ref_pbs_batch(out.ctypes.data_as(_u64p), pc, C.c_long(B),
                               bsk_f.ctypes.data_as(_f64p), pl, pi, C.c_int(n), C.c_int(k),
                               C.c_int(N), C.c_int(level), C.c_int(base_log))
        return out

    def pbs_ex(self, ct, bsk_f, luts, lut_idx, base_log, n, bsk_ct_stride, theta=0, nout=1):
        """Extended bootstrap.  bsk_ct_stride > 0: one list of n Fourier GGSWs per ciphertext (bsk_f: [..., level,
        k+1, k+1, M, 2]; stride in complex elements).  theta / nout: multi-output bootstrap, result [B, nout, kN+1]."""
        level, k1, _, M, _ = bsk_f.shape[-5:]
        k, N = k1 - 1, 2 * M
        ct, pc = self._u64(ct)
        luts, pl = self._u64(luts)
        B = ct.size // (n + 1)
        out = np.empty((B, nout, k * N + 1), dtype=np.uint64)
        lut_idx = np.ascontiguousarray(lut_idx, dtype=np.int32)
        bsk_f = np.ascontiguousarray(bsk_f)
        self.lib.ref_pbs_batch_ex(out.ctypes.data_as(_u64p), pc, C.c_long(B), bsk_f.ctypes.data_as(_f64p), pl,
                                  lut_idx.ctypes.data_as(_i32p), C.c_int(n), C.c_int(k), C.c_int(N),
                                  C.c_int(level), C.c_int(base_log), C.c_long(bsk_ct_stride), C.c_int(theta),
                                  C.c_int(nout))
        return out[:, 0] if nout == 1 and theta == 0 else out

    def cmux_batch(self, pool, i0, i1, ggsw_f, gi, base_log):
        """out[b] = pool[i0[b]] + ggsw[gi[b]] (x) (pool[i1[b]] - pool[i0[b]]); pool [P, (k+1)N]."""
        level, k1, _, M, _ = ggsw_f.shape[-5:]
        k, N = k1 - 1, 2 * M
        pool, pp = self._u64(pool)
        i0 = np.ascontiguousarray(i0, dtype=np.int32)
        i1 = np.ascontiguousarray(i1, dtype=np.int32)
        gi = np.ascontiguousarray(gi, dtype=np.int32)
        ggsw_f = np.ascontiguousarray(ggsw_f)
        out = np.empty((len(i0), (k + 1) * N), dtype=np.uint64)
        self.lib.ref_cmux_batch(out.ctypes.data_as(_u64p), pp, i0.ctypes.data_as(_i32p), i1.ctypes.data_as(_i32p),
                                ggsw_f.ctypes.data_as(_f64p), gi.ctypes.data_as(_i32p), C.c_long(len(i0)),
                                C.c_int(k), C.c_int(N), C.c_int(level), C.c_int(base_log))
        return out

    def build_lut_poly(self, table, p, p_out, N):
        table = np.ascontiguousarray(np.asarray(table, dtype=np.int64))
        assert table.size == (1 << p)
        poly = np.empty(N, dtype=np.uint64)
        self.lib.ref_build_lut_poly(poly.ctypes.data_as(_u64p), table.ctypes.data_as(_i64p),
                                    C.c_int(p), C.c_int(p_out), C.c_int(N))
        return poly

    # -- linear ops on [E, L] tensors
    def lin_add(self, a, b):
        a, pa = self._u64(a)
        b, pb = self._u64(b)
        out = np.empty_like(a)
        self.lib.ref_lin_add(out.ctypes.data_as(_u64p), pa, pb, C.c_long(a.size))
        return out

    def lin_sub(self, a, b):
        a, pa = self._u64(a)
        b, pb = self._u64(b)
        out = np.empty_like(a)
        self.lib.ref_lin_sub(out.ctypes.data_as(_u64p), pa, pb, C.c_long(a.size))
        return out

    def lin_neg(self, a):
        a, pa = self._u64(a)
        out = np.empty_like(a)
        self.lib.ref_lin_neg(out.ctypes.data_as(_u64p), pa, C.c_long(a.size))
        return out

    def lin_add_plain(self, a, pt):
        a, pa = self._u64(a)
        E, L = a.shape
        pt, pp = self._u64(np.asarray(pt).reshape(-1))
        out = np.empty_like(a)
        self.lib.ref_lin_add_plain(out.ctypes.data_as(_u64p), pa, pp, C.c_long(E), C.c_int(L))
        return out

    def lin_mul_clear(self, a, c):
        a, pa = self._u64(a)
        E, L = a.shape
        c = np.ascontiguousarray(np.asarray(c, dtype=np.int64).reshape(-1))
        out = np.empty_like(a)
        self.lib.ref_lin_mul_clear(out.ctypes.data_as(_u64p), pa, c.ctypes.data_as(_i64p),
                                   C.c_long(E), C.c_int(L))
        return out

    def lin_gather_mac(self, x, idx, w, bias=None):
        x, px = self._u64(x)
        L = x.shape[-1]
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.int64)
        E, K = idx.shape
        out = np.empty((E, L), dtype=np.uint64)
        if bias is None:
            pb = None
        else:
            bias, pb = self._u64(np.asarray(bias).reshape(-1))
        self.lib.ref_lin_gather_mac(out.ctypes.data_as(_u64p), px, idx.ctypes.data_as(_i32p),
                                    w.ctypes.data_as(_i64p), pb, C.c_long(E), C.c_int(K),
                                    C.c_int(L))
        return out


DOM_SK_SMALL, DOM_SK_BIG = 1, 2
