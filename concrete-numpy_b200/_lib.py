"""
ctypes binding of the C-ABI library `libfhe_b200.so` (declared in include/fhe_b200.h).

The product path has NO CPU fallback: if the CUDA library is missing or a call fails, a
`RuntimeError` carrying `fhe_last_error()` is raised.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfhe_b200.so")

_p = C.c_void_p
_SIGS = {
    # name: (restype, [argtypes])
    "fhe_last_error": (C.c_char_p, []),
    "fhe_abi_version": (C.c_int, []),
    "fhe_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long)]),
    "fhe_measure_fp64_peak": (C.c_int, [C.POINTER(C.c_double), _p]),
    "fhe_seed_expand": (C.c_int, [C.c_uint64, _p]),
    "fhe_keygen_lwe_secret": (C.c_int, [_p, C.c_int, _p, C.c_int, _p]),
    "fhe_keygen_ksk": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_double, _p, _p]),
    "fhe_keygen_bsk": (C.c_int, [_p, _p, C.c_int, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _p, _p, _p]),
    "fhe_encrypt_lwe_batch": (C.c_int, [_p, _p, C.c_long, _p, C.c_int, C.c_double, _p, C.c_uint64, _p]),
    "fhe_trivial_lwe_batch": (C.c_int, [_p, _p, C.c_long, C.c_int, _p]),
    "fhe_decrypt_lwe_batch": (C.c_int, [_p, _p, _p, C.c_long, _p, C.c_int, C.c_int, _p]),
    "fhe_keyswitch_batch": (C.c_int, [_p, _p, C.c_long, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_ks_tc_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_ksk_tc_bytes": (C.c_long, [C.c_int, C.c_int, C.c_int]),
    "fhe_keyswitch_tc_scratch_bytes": (C.c_long, [C.c_long, C.c_int, C.c_int]),
    "fhe_ksk_to_tc": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_keyswitch_batch_tc": (C.c_int, [_p, _p, C.c_long, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_cluster_size": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "fhe_bsk_fourier_elems": (C.c_long, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_bsk_to_fourier": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_batch": (C.c_int, [_p, _p, C.c_long, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_num_variants": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "fhe_pbs_variant_pipelined": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_pbs_variant_grouped": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_pbs_variant_cluster_size": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_pbs_variant_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_variant_capacity": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fhe_bsk_to_fourier_variant": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_bsk_relayout": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_batch_variant": (C.c_int, [_p, _p, C.c_long, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_batch_multi": (C.c_int, [_p, _p, C.c_long, _p, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_cmux_batch": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_blind_rotate_batch": (C.c_int, [_p, _p, C.c_long, _p, C.c_long, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_pbs_set_profile": (C.c_int, [_p]),
    "fhe_pbs_set_pipeline": (C.c_int, [C.c_int]),
    "fhe_pbs_set_group": (C.c_int, [C.c_int]),
    "fhe_debug_polymul": (C.c_int, [_p, _p, _p, C.c_long, C.c_int, C.c_int, C.c_int, _p]),
    "fhe_lin_add": (C.c_int, [_p, _p, _p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_sub": (C.c_int, [_p, _p, _p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_neg": (C.c_int, [_p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_add_plain": (C.c_int, [_p, _p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_plain_sub": (C.c_int, [_p, _p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_mul_clear": (C.c_int, [_p, _p, _p, _p, C.c_long, C.c_int, _p]),
    "fhe_lin_gather_mac": (C.c_int, [_p, _p, _p, _p, _p, C.c_long, C.c_int, C.c_int, _p]),
    "fhe_matmul_ct_clear": (C.c_int, [_p, _p, _p, C.c_long, C.c_int, C.c_int, C.c_int, _p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


class FheError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""


def load():
    """Load libfhe_b200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(there is no CPU fallback for the encrypted-execution path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


KEY_BYTES = 64       # key material: 256-bit secret ChaCha key + 256-bit public-mask key


def key_material(seed):
    """16-word key material for the keygen / encrypt entry points: 64 bytes are used as they are (production:
    os.urandom(64)); an int is a TEST seed, expanded deterministically by fhe_seed_expand."""
    key = (C.c_uint32 * 16)()
    if isinstance(seed, (bytes, bytearray)):
        if len(seed) != KEY_BYTES:
            raise ValueError(f"key material is {KEY_BYTES} bytes (secret key, public-mask key)")
        C.memmove(key, bytes(seed), KEY_BYTES)
    else:
        rc = load().fhe_seed_expand(C.c_uint64(int(seed) & ((1 << 64) - 1)), key)
        if rc != 0:
            raise FheError("fhe_seed_expand failed")
    return key


LAUNCH_COUNTS = {}   # entry point -> number of calls (each call launches >= 1 CUDA kernel)


def launches_total() -> int:
    return sum(LAUNCH_COUNTS.values())


def call(name, *args):
    """Invoke a status-returning entry point and convert failures into FheError."""
    lib = load()
    LAUNCH_COUNTS[name] = LAUNCH_COUNTS.get(name, 0) + 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FheError(f"{name} failed with status {rc}: {lib.fhe_last_error().decode()}")

