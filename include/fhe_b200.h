/*
 * fhe_b200.h -- C ABI of the B200-native encrypted-execution backend (libfhe_b200.so).
 *
 * This is the drop-in boundary for the hot path behind concrete-numpy's Circuit.run / Server.run:
 * the entry points a runtime for the reference's FHE / FHELinalg programs binds to.  In the
 * reference the same role is played by the C runtime of the (absent) concrete-compiler wheel
 * (`concretelang/Runtime/wrappers.cpp`: memref_keyswitch_lwe_u64, memref_bootstrap_lwe_u64,
 * memref_batched_*, memref_add_lwe_ciphertexts_u64, ... -- SURVEY.md section 2.2), reached from
 * concrete/numpy/compilation/server.py:286 (`self._support.server_call(...)`) and, for keys /
 * encryption / decryption, from concrete/numpy/compilation/client.py:106, 178-182, 201.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (cudaMalloc / torch .data_ptr()), never freed or
 *     reallocated by the library; `stream` is a cudaStream_t passed as void*;
 *   - torus words are uint64_t (q = 2^64); an LWE ciphertext of dimension d is d mask words
 *     followed by the body; tensors of ciphertexts are [E][d+1] row-major;
 *   - every function returns 0 on success, 2 for invalid arguments, 3/4 for unsupported
 *     shapes, 1000+cudaError_t for CUDA failures; fhe_last_error() returns the message of the
 *     last failure on the calling thread;
 *   - calls are asynchronous on `stream` (no device synchronisation inside), one host thread per
 *     device.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- housekeeping ------------------------------------------------------------------- */
const char* fhe_last_error(void);
int fhe_abi_version(void);
int fhe_device_info(int* sm_count, int* cc_major, int* cc_minor, long* total_mem);
/* DFMA throughput of the current device in TFLOP/s (roofline denominator of the bootstrap) */
int fhe_measure_fp64_peak(double* tflops, void* stream);

/* ---- key generation: replaces ClientSupport.key_set (client.py:96-106) ---------------
 * Keys are a pure function of (key material, parameters): counter-based ChaCha20 streams.
 * key16: HOST pointer to 16 words = a SECRET 256-bit ChaCha key (secret-key bits, noise) followed by a separate
 * 256-bit key for the PUBLIC masks of BSK / KSK / ciphertexts.  Production callers fill it with 64 bytes of OS
 * entropy (os.urandom); fhe_seed_expand() is the deterministic 64-bit-seed expansion for tests and for
 * regenerating identical keys on every rank of a benchmark.
 * domain: 1 = small LWE key (dimension n), 2 = big key (k*N, the flattened GLWE key). */
int fhe_seed_expand(uint64_t seed, uint32_t* key16);
int fhe_keygen_lwe_secret(uint64_t* sk, int dim, const uint32_t* key16, int domain, void* stream);
/* KSK [kN][level][n+1]: row (i, lv) encrypts sk_big[i] * 2^(64-(lv+1)*base_log) under sk_small */
int fhe_keygen_ksk(uint64_t* ksk, const uint64_t* sk_big, int kN, const uint64_t* sk_small, int n,
                   int level, int base_log, double std, const uint32_t* key16, void* stream);
/* standard-domain BSK [n][level][k+1][k+1][N]; scratch: >= k*N + k int32 words */
int fhe_keygen_bsk(uint64_t* bsk, const uint64_t* sk_small, int n, const uint64_t* sk_glwe, int k,
                   int N, int level, int base_log, double std, const uint32_t* key16, int* scratch,
                   void* stream);
/* Fourier-domain BSK conversion on device; layout owned by the bootstrap kernel.
 * fhe_bsk_fourier_elems() = number of (re, im) double pairs of the output buffer. */
int fhe_pbs_cluster_size(int N, int k, int level);
long fhe_bsk_fourier_elems(int n, int k, int N, int level);
int fhe_bsk_to_fourier(void* bsk_f /* double2* */, const uint64_t* bsk_std, int n, int k, int N,
                       int level, int base_log, void* stream);

/* ---- encrypt / decrypt: replace ClientSupport.encrypt_arguments / decrypt_result -----
 * (client.py:178-182, 201).  plaintexts are already encoded (m << (63 - p)). */
int fhe_encrypt_lwe_batch(uint64_t* ct /* [B][dim+1] */, const uint64_t* plaintexts /* [B] */,
                          long B, const uint64_t* sk, int dim, double std, const uint32_t* key16,
                          uint64_t first_index, void* stream);
int fhe_trivial_lwe_batch(uint64_t* ct, const uint64_t* plaintexts, long B, int dim, void* stream);
/* phase[e] = body - <a, s>; msg[e] = phase decoded at p bits + 1 padding bit (either may be NULL) */
int fhe_decrypt_lwe_batch(uint64_t* phase, uint64_t* msg, const uint64_t* ct, long B,
                          const uint64_t* sk, int dim, int p, void* stream);

/* ---- table lookup = keyswitch + programmable bootstrap -------------------------------
 * replace memref_batched_keyswitch_lwe_u64 / memref_batched_bootstrap_lwe_u64 for
 * FHE.apply_lookup_table, FHELinalg.apply_lookup_table, apply_mapped_lookup_table
 * (concrete/numpy/mlir/node_converter.py:1190-1206). */
int fhe_keyswitch_batch(uint64_t* out /* [B][n+1] */, const uint64_t* in /* [B][kN+1] */, long B,
                        const uint64_t* ksk, int kN, int n, int level, int base_log, void* stream);
/* Tensor-core keyswitch (csrc/keyswitch_tc.cu): the same contraction as an int8 x uint8 -> int32 GEMM on
 * tcgen05 (8 byte limbs per key word), bit-exact; usable when fhe_ks_tc_supported() (digits fit int8 and no
 * int32 limb sum can overflow).  The key is re-laid out once (fhe_ksk_to_tc), the digits of a batch go to a
 * caller-provided scratch buffer of fhe_keyswitch_tc_scratch_bytes(). */
int fhe_ks_tc_supported(int kN, int n, int level, int base_log);
long fhe_ksk_tc_bytes(int kN, int n, int level);
long fhe_keyswitch_tc_scratch_bytes(long B, int kN, int level);
int fhe_ksk_to_tc(void* ksk_tc, const uint64_t* ksk, int kN, int n, int level, void* stream);
int fhe_keyswitch_batch_tc(uint64_t* out, const uint64_t* in, long B, const void* ksk_tc, void* scratch,
                           int kN, int n, int level, int base_log, void* stream);
int fhe_pbs_batch(uint64_t* out /* [B][kN+1] */, const uint64_t* in /* [B][n+1] */, long B,
                  const void* bsk_f, const uint64_t* luts /* [T][(k+1)N] accumulators */,
                  const int* lut_idx /* [B] or NULL */, int n, int k, int N, int level, int base_log,
                  void* stream);

/* Plan variants of the bootstrap: variant 0 is the throughput plan (smallest cluster, what
 * fhe_bsk_to_fourier / fhe_pbs_batch use); higher variants spread one ciphertext over larger clusters and
 * are meant for small batches (sequential circuits).  The Fourier BSK layout depends on the variant:
 * fhe_bsk_relayout permutes a key from one variant's layout into another's (same spectrum values). */
int fhe_pbs_num_variants(int N, int k, int level);
/* 1 if two-level k = 1 bootstraps of the variant run the pipelined kernel, 0 if not, < 0 if there is no such variant */
int fhe_pbs_variant_pipelined(int N, int k, int level, int variant);
/* 1 if k = 1 bootstraps of the variant run the grouped kernel (one CTA per SM, several ciphertext slots), 0 if not */
int fhe_pbs_variant_grouped(int N, int k, int level, int variant);
int fhe_pbs_variant_cluster_size(int N, int k, int level, int variant);
/* plan of a variant: meta[0..7] = cluster size, log2 of the stage-1 / mid-pass A / mid-pass B / fused radices,
 * threads per CTA, minimum CTAs per SM, largest GLWE dimension (documentation / tests) */
int fhe_pbs_variant_plan(int N, int k, int level, int variant, int* meta);
/* ciphertexts the variant bootstraps concurrently on the current device (one wave of clusters) */
int fhe_pbs_variant_capacity(int N, int k, int level, int variant);
int fhe_bsk_to_fourier_variant(void* bsk_f, const uint64_t* bsk_std, int n, int k, int N, int level,
                               int base_log, int variant, void* stream);
int fhe_bsk_relayout(void* dst_bsk_f, const void* src_bsk_f, int n, int k, int N, int level,
                     int src_variant, int dst_variant, void* stream);
int fhe_pbs_batch_variant(uint64_t* out, const uint64_t* in, long B, const void* bsk_f,
                          const uint64_t* luts, const int* lut_idx, int n, int k, int N, int level,
                          int base_log, int variant, void* stream);

/* ---- wide integers (9..16 bits): building blocks of the CRT / without-padding lookup -----------------
 * The reference announces this path through the `crt` list of its client parameters
 * (concrete/numpy/compilation/client.py:213-235) and allows table lookups up to 16 bits
 * (concrete/numpy/mlir/utils.py:16); the arithmetic lives in the absent wheel.  Host composition:
 * concrete-numpy_b200/wide.py (bit extraction = keyswitch + sign bootstraps, circuit bootstrapping = bootstraps +
 * packing keyswitch = fhe_keyswitch_batch[_tc] with (k+1)^2 N columns, Fourier conversion = fhe_bsk_to_fourier).
 *
 * fhe_cmux_batch: out[b] = pool[i0[b]] + GGSW[gi[b]] (x) (pool[i1[b]] - pool[i0[b]]) on GLWE ciphertexts of
 * (k+1)N words (CMUX tree of vertical packing).
 * fhe_blind_rotate_batch: the bootstrap's blind rotation + sample extraction with one list of n GGSWs per
 * ciphertext (ct_stride double2 elements apart) instead of a bootstrap key. */
/* fhe_pbs_batch_multi: one blind rotation, `nout` <= 2^theta results on the same input (PBSmanyLUT): the
 * accumulators interleave the functions (coefficient i belongs to function i mod 2^theta), the modulus switch
 * keeps theta zero low bits; out [B][nout][kN+1].  Used for the sign bootstraps of bit extraction and circuit
 * bootstrapping, which all act on the same extracted-bit ciphertext. */
int fhe_pbs_batch_multi(uint64_t* out, const uint64_t* in, long B, const void* bsk_f, const uint64_t* luts,
                        const int* lut_idx, int n, int k, int N, int level, int base_log, int theta, int nout,
                        int variant, void* stream);
int fhe_cmux_batch(uint64_t* out, const uint64_t* pool, const int* i0, const int* i1, const void* ggsw_f,
                   const int* gi, long B, int k, int N, int level, int base_log,
                   int plain /* pool entries are plaintext polynomials (zero masks) */, void* stream);
int fhe_blind_rotate_batch(uint64_t* out, const uint64_t* in, long B, const void* ggsw_f, long ct_stride,
                           const uint64_t* luts, const int* lut_idx, int n, int k, int N, int level,
                           int base_log, void* stream);

/* development aid: k_pbs accumulates per-phase clock64() deltas of CTA b < 16 into dev_buf[16*b .. 16*b+15]
 * (NULL switches it off); not part of the reference-facing surface */
int fhe_pbs_set_profile(long long* dev_buf);
/* test / measurement hook: 0 makes the pipelined bootstrap plans (k_pbs_pipe) run their generic kernel; returns the
 * previous setting.  Both kernels are bit-identical by construction; not part of the reference-facing surface */
int fhe_pbs_set_pipeline(int enabled);
/* the same switch for the grouped plans (k_pbs_group: two ciphertexts per CTA at N = 2048, twiddles in tensor memory):
 * 0 = k_pbs, 1 = grouped with GGSW_i streamed from L2 (default), 2 = grouped with GGSW_i staged once per SM in shared memory */
int fhe_pbs_set_group(int enabled);

/* ---- torus linear algebra on [E][L] tensors (L = kN+1) --------------------------------
 * ia / ib: optional int32 row maps (numpy broadcasting resolved by the host), NULL = identity.
 * FHE.add_eint / FHELinalg.add_eint (node_converter.py:216-242) ... */
int fhe_lin_add(uint64_t* out, const uint64_t* a, const uint64_t* b, const int* ia, const int* ib,
                long E, int L, void* stream);
int fhe_lin_sub(uint64_t* out, const uint64_t* a, const uint64_t* b, const int* ia, const int* ib,
                long E, int L, void* stream);
int fhe_lin_neg(uint64_t* out, const uint64_t* a, const int* ia, long E, int L, void* stream);
/* add_eint_int / sub_eint_int: out[e] = a[ia[e]], body += pt[e] (encoded plaintext) */
int fhe_lin_add_plain(uint64_t* out, const uint64_t* a, const int* ia, const uint64_t* pt, long E,
                      int L, void* stream);
/* sub_int_eint: out[e] = pt[e] - a[ia[e]] */
int fhe_lin_plain_sub(uint64_t* out, const uint64_t* a, const int* ia, const uint64_t* pt, long E,
                      int L, void* stream);
/* mul_eint_int (node_converter.py:630-650): every word times the signed clear c[e] */
int fhe_lin_mul_clear(uint64_t* out, const uint64_t* a, const int* ia, const int64_t* c, long E,
                      int L, void* stream);
/* general clear-weighted gather (dot / matmul / conv2d / sum, node_converter.py:436-524, 596-616,
 * 1094-1127): out[e] = sum_j w[e][j] * x[idx[e][j]] (+ bias[e] on the body); idx < 0 = no term */
int fhe_lin_gather_mac(uint64_t* out, const uint64_t* x, const int* idx, const int64_t* w,
                       const uint64_t* bias, long E, int K, int L, void* stream);
/* FHELinalg.matmul_eint_int on a plain 2-D problem: x [Mr][K][L], W [K][Nc] -> out [Mr][Nc][L] */
int fhe_matmul_ct_clear(uint64_t* out, const uint64_t* x, const int64_t* W, long Mr, int K, int Nc,
                        int L, void* stream);

/* ---- test hook: negacyclic product through the fp64 FFT path --------------------------- */
int fhe_debug_polymul(uint64_t* out, const int64_t* a_small, const uint64_t* b_torus, long npoly,
                      int N, int k, int level, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
