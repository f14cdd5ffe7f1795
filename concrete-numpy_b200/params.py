"""
Crypto-parameter selection for one circuit: stands in for `concrete-optimizer` (inside the absent
concrete-compiler wheel), which the reference drives through
`CompilationOptions.set_p_error / set_global_p_error` (concrete/numpy/compilation/server.py:102-126).

Given the circuit-wide precision `p` (the reference forces ONE width on every encrypted value,
concrete/numpy/mlir/graph_converter.py:315-322), the largest squared 2-norm of the clear linear
maps feeding a table lookup, and the per-lookup error budget, search (n, k, N, pbs level/base,
ks level/base) minimising an fp64/int cost model subject to

    erfc( 2^-(p+2) / (sqrt(2) * sigma_total) ) <= p_error
    sigma_total^2 = norm2 * V_pbs_out + V_keyswitch + V_modswitch

with the variance formulas below (torus units, binary keys) and 128-bit security curves
(noise std as a function of LWE dimension) fitted to the published TFHE-rs `PARAM_MESSAGE_x_CARRY_y`
sets, which were produced by the same concrete-optimizer.

The formulas are validated empirically on device by tests/test_noise.py.
"""

import math
from functools import lru_cache

import numpy as np

from .engine import CryptoParams

SECURITY_SLOPE = -0.0264
SECURITY_BIAS = 2.48
MIN_LOG2_STD = -62.0          # ~2.17e-19, the floor used for 64-bit torus
SMEM_MAX = 232448             # must match PBS_SMEM_MAX in csrc/pbs.cu


def secure_log2_std(dim: int) -> float:
    """log2 of the minimal noise std (relative to q) giving 128-bit security at LWE dimension `dim`."""
    return max(SECURITY_SLOPE * dim + SECURITY_BIAS, MIN_LOG2_STD)


# (C, logM, lra, lrb, lrf, kmax, minb) of every FFT plan instantiated in csrc/pbs_inst_{a,b,c}.cu
_PBS_PLANS = [
    (1, 9, 3, 0, 3, 1, 2), (1, 8, 3, 0, 3, 1, 2), (1, 7, 3, 0, 3, 1, 1),
    (1, 7, 3, 0, 2, 3, 1), (1, 8, 3, 0, 2, 3, 1), (1, 9, 3, 0, 2, 3, 1), (1, 10, 4, 0, 2, 3, 1),
    (2, 12, 4, 4, 3, 1, 1), (2, 11, 4, 3, 3, 1, 2), (1, 11, 4, 0, 3, 1, 1), (1, 10, 4, 0, 3, 1, 2), (1, 10, 4, 0, 3, 1, 1),
    (16, 14, 4, 3, 3, 1, 2), (8, 14, 4, 4, 3, 1, 1), (8, 13, 4, 3, 3, 1, 2), (4, 13, 4, 4, 3, 1, 1), (4, 12, 4, 3, 3, 1, 2),
]


def pbs_cluster_size(N: int, k: int, level: int) -> int:
    """CTAs per ciphertext of the bootstrap plan for this shape, -1 if no on-chip layout fits.  Asked from the
    library itself (fhe_pbs_cluster_size: pure host logic, no GPU needed) so that the parameter search can never
    pick a shape the kernels reject; the Python mirror below only serves when the library is not built, and
    tests/test_abi_cpu.py checks that the two agree on every shape."""
    try:
        from . import _lib

        return int(_lib.load().fhe_pbs_cluster_size(int(N), int(k), int(level)))
    except Exception:                                                   # noqa: BLE001  (library absent: CPU-only tooling)
        return pbs_cluster_size_mirror(N, k, level)


def pbs_cluster_size_mirror(N: int, k: int, level: int) -> int:
    """Mirror of fitting_plans() in csrc/pbs.cu."""
    logM = N.bit_length() - 2
    if N != (2 << logM) or not 8 <= logM + 1 <= 15 or not 1 <= k <= 3 or not 1 <= level <= 4:
        return -1
    for C in (1, 2, 4, 8, 16):
        for (pc, plogM, lra, lrb, lrf, kmax, minb) in _PBS_PLANS:
            if pc != C or plogM != logM or k > kmax or (k <= 1 and kmax > 1):
                continue
            logL1 = lra + lrb + lrf
            tw = 0          # twiddles live in tensor memory for the one-CTA-per-SM plans (minb == 1)
            if minb != 1:
                if lra:
                    tw += ((1 << lra) - 1) << (logL1 - lra)
                if lrb:
                    tw += ((1 << lrb) - 1) << lrf
            buf = (1 << logM) // C if C > 1 else (1 << logM)
            smem = (k + 1) * (N // C) * 8 + (k + 1) * level * buf * 16 + tw * 16
            if smem <= SMEM_MAX:
                return C
    return -1


# ---------------------------------------------------------------- noise model ----

def variance_pbs_out(n, k, N, level, base_log, glwe_var):
    """Output variance of a programmable bootstrap (blind rotation of n CMUXes)."""
    B2 = 4.0 ** base_log
    ext = level * (k + 1) * N * (B2 + 2.0) / 12.0 * glwe_var
    rounding = (1.0 + k * N / 2.0) / (12.0 * 4.0 ** (base_log * level))
    # fp64 FFT precision loss of one external product (53-bit mantissa on a 64-bit torus): the
    # empirical law published with concrete-optimizer, 2^-2.577 * 2^22 * l * B^2 * N^2 * (k+1) / 2^128;
    # measured on device by scripts/noise_probe.py (within 1 bit, on the conservative side)
    fft = level * (k + 1) * B2 * float(N) ** 2 * 2.0 ** (-2.577 + 22 - 128)
    return n * (ext + rounding + fft)


def variance_keyswitch(kN, level, base_log, lwe_var):
    B2 = 4.0 ** base_log
    return kN * level * (B2 + 2.0) / 12.0 * lwe_var + kN / 2.0 / (12.0 * 4.0 ** (base_log * level))


def variance_modswitch(n, N, theta=0):
    """theta > 0: the multi-output bootstrap switches to 2N / 2^theta levels (fhe_pbs_batch_multi)."""
    return (1.0 / 12.0 + n / 24.0) / (4.0 * N * N) * 4.0 ** theta


def p_error_of(p, variance):
    """Probability that noise of the given variance leaves the half-box 2^-(p+2)."""
    if variance <= 0:
        return 0.0
    z = 2.0 ** -(p + 2) / math.sqrt(2.0 * variance)
    return math.erfc(z)


def input_variance(params: CryptoParams, norm2: float) -> float:
    """Variance at the input of a table lookup (after clear linear maps of squared 2-norm norm2)."""
    v_out = variance_pbs_out(params.n, params.k, params.N, params.pbs_level, params.pbs_base_log,
                             params.glwe_std ** 2)
    v_in = max(v_out, params.glwe_std ** 2) * max(norm2, 1.0)
    return (v_in + variance_keyswitch(params.big_dim, params.ks_level, params.ks_base_log, params.lwe_std ** 2)
            + variance_modswitch(params.n, params.N))


def estimate_p_error(params: CryptoParams, norm2: float) -> float:
    return p_error_of(params.p, input_variance(params, norm2))


# ---------------------------------------------------------------- cost model -----

def pbs_flops(n, k, N, level):
    """Real flops of one bootstrap (BASELINE.md section 4 / SURVEY 8d)."""
    M = N / 2
    return n * ((k + 1) * (level + 1) * 5 * M * math.log2(M) + 8 * (k + 1) ** 2 * level * M)


def ks_macs(kN, n, level):
    return kN * level * (n + 1)


KS_MAC_COST = 7.0   # one u64 MAC on the int pipe costs about this many fp64 flops of device time

WIDE_SHAPES = ((1, 2048), (1, 4096), (2, 1024), (3, 512))      # (k, N) candidates of the wide search
# fraction of the k = 1, N = 2048 kernels' flop rate the bootstrap reaches at each shape (16-bit lookups on B200,
# scripts/wide_bench.py: 7.0 / 3.5 / 2.5 TFLOP/s; profiles/r01_wide_lookup_bench.log): the search minimises time
WIDE_SHAPE_EFFICIENCY = {(1, 2048): 1.0, (1, 4096): 0.9, (2, 1024): 0.5, (3, 512): 0.36}


def wide_efficiency(k: int, N: int, pbs_level: int) -> float:
    """... and by decomposition depth: up to two levels the N = 2048 bootstrap runs the grouped kernel (two ciphertext slots
    per SM, 12.1 TFLOP/s); deeper decompositions leave room for one ciphertext per SM only and run the generic kernel
    (8.3 TFLOP/s at three levels; scripts/group_ablation.py, profiles/r02_group_ablation.log)."""
    eff = WIDE_SHAPE_EFFICIENCY.get((k, N), 0.5)
    if (k, N) == (1, 2048) and pbs_level >= 3:
        eff *= 0.69
    return eff


@lru_cache(maxsize=None)
def select_parameters(p: int, norm2: float = 1.0, p_error: float = 1e-9) -> CryptoParams:
    """Cheapest parameter set meeting `p_error` per table lookup at precision p."""
    if p < 1 or p > 8:
        raise ValueError(
            f"precision {p} bits is outside the native PBS range 1..8 of this backend "
            "(the reference would switch to CRT / WoP-PBS; see DESIGN.md 'out of scope')"
        )
    target_z = _z_for_p_error(p_error)
    max_var = (2.0 ** -(p + 2) / target_z) ** 2        # erfc(z/sqrt2) <= p_error  <=> sigma <= delta/z
    best = None
    shapes = [(3, 512), (2, 1024), (1, 2048), (1, 4096), (1, 8192), (1, 16384), (1, 32768)]
    n_grid = np.arange(450, 1201, 2)
    lwe_var_grid = 4.0 ** np.maximum(SECURITY_SLOPE * n_grid + SECURITY_BIAS, MIN_LOG2_STD)
    for k, N in shapes:
        if N < (1 << (p + 1)):
            continue
        glwe_var = 4.0 ** secure_log2_std(k * N)
        v_ms = (1.0 / 12.0 + n_grid / 24.0) / (4.0 * N * N)
        for level in (1, 2, 3):
            if pbs_cluster_size(N, k, level) < 0:
                continue
            for base_log in range(4, 36):
                if level * base_log > 52:
                    break
                v_out_per_n = variance_pbs_out(1, k, N, level, base_log, glwe_var)
                v_in = np.maximum(v_out_per_n * n_grid, glwe_var) * max(norm2, 1.0)
                budget = max_var - v_in - v_ms                     # what is left for the keyswitch
                if not np.any(budget > 0):
                    continue
                pbs_cost = np.array([pbs_flops(int(n), k, N, level) for n in (n_grid[0], n_grid[-1])])
                pbs_cost = pbs_cost[0] + (pbs_cost[1] - pbs_cost[0]) * (n_grid - n_grid[0]) / (n_grid[-1] - n_grid[0])
                for ks_level in range(1, 10):
                    for ks_base_log in range(1, 12):
                        if ks_level * ks_base_log > 40:
                            break
                        B2 = 4.0 ** ks_base_log
                        v_ks = (k * N * ks_level * (B2 + 2.0) / 12.0 * lwe_var_grid
                                + k * N / 24.0 / 4.0 ** (ks_base_log * ks_level))
                        ok = v_ks <= budget
                        if not np.any(ok):
                            continue
                        cost = pbs_cost + KS_MAC_COST * k * N * ks_level * (n_grid + 1)
                        # modelled work -> time: the k > 1 plans reach a fraction of the k = 1 kernels' flop rate
                        cost = np.where(ok, cost, np.inf) / WIDE_SHAPE_EFFICIENCY.get((k, N), 1.0)
                        j = int(np.argmin(cost))
                        if best is None or cost[j] < best[0]:
                            best = (float(cost[j]), int(n_grid[j]), k, N, level, base_log, ks_level, ks_base_log)
    if best is None:
        raise ValueError(f"no parameter set reaches p_error={p_error:g} at {p} bits with 2-norm^2={norm2:g}")
    _, n, k, N, level, base_log, ks_level, ks_base_log = best
    return CryptoParams(p=p, n=n, k=k, N=N, pbs_level=level, pbs_base_log=base_log, ks_level=ks_level,
                        ks_base_log=ks_base_log, lwe_std=2.0 ** secure_log2_std(n),
                        glwe_std=2.0 ** secure_log2_std(k * N))


# ---------------------------------------------------------------- circuits without table lookups ----

MAX_LINEAR_BITS = 55


@lru_cache(maxsize=None)
def select_linear_parameters(p: int, norm2: float = 1.0, p_error: float = 1e-9) -> CryptoParams:
    """Parameters of a circuit WITHOUT table lookups (the reference accepts any width there, only lookups are limited
    to 16 bits: concrete/numpy/mlir/graph_converter.py:288-306, e.g. tests/execution/test_add.py with values up to
    10^6): one LWE ciphertext per value in the native encoding, no bootstrap / keyswitch keys at all.  The LWE
    dimension is the smallest whose 128-bit-secure noise leaves the half-step 2^-(p+2) at the requested error rate."""
    if p < 1 or p > MAX_LINEAR_BITS:
        raise ValueError(f"{p}-bit values do not fit one 64-bit torus ciphertext (up to {MAX_LINEAR_BITS} bits)")
    z = _z_for_p_error(p_error)
    need = -(p + 2) - math.log2(z) - 0.5 * math.log2(max(norm2, 1.0))       # log2 of the largest admissible noise std
    if need < MIN_LOG2_STD:
        raise ValueError(f"no noise level supports {p} bits with 2-norm^2={norm2:g} at p_error={p_error:g}")
    dim = int(math.ceil((need - SECURITY_BIAS) / SECURITY_SLOPE))
    dim = max(512, (dim + 15) // 16 * 16)
    return CryptoParams(p=p, n=0, k=1, N=dim, pbs_level=0, pbs_base_log=0, ks_level=0, ks_base_log=0,
                        lwe_std=0.0, glwe_std=2.0 ** secure_log2_std(dim))


# ---------------------------------------------------------------- wide integers (CRT) ----

def variance_packing_keyswitch(kN, level, base_log, glwe_var):
    """Noise the packing keyswitch of circuit bootstrapping adds to a GGSW row (wide.py)."""
    B2 = 4.0 ** base_log
    return (kN + 1) * level * (B2 + 2.0) / 12.0 * glwe_var + (kN / 2.0) / (12.0 * 4.0 ** (base_log * level))


def variance_cmux(k, N, level, base_log, ggsw_var):
    """One external product of vertical packing with a circuit-bootstrapped GGSW of row variance ggsw_var."""
    B2 = 4.0 ** base_log
    ext = level * (k + 1) * N * (B2 + 2.0) / 12.0 * ggsw_var
    rounding = (1.0 + k * N / 2.0) / (12.0 * 4.0 ** (base_log * level))
    fft = level * (k + 1) * B2 * float(N) ** 2 * 2.0 ** (-2.577 + 22 - 128)
    return ext + rounding + fft


def wide_output_variance(params: CryptoParams) -> float:
    """Variance of a residue ciphertext coming out of a wide lookup (depth = number of extracted bits)."""
    from .wide import total_bits

    g2 = params.glwe_std ** 2
    v_pbs = variance_pbs_out(params.n, params.k, params.N, params.pbs_level, params.pbs_base_log, g2)
    v_ggsw = v_pbs + variance_packing_keyswitch(params.big_dim, params.pf_level, params.pf_base_log, g2)
    return total_bits(params.moduli) * variance_cmux(params.k, params.N, params.cbs_level, params.cbs_base_log, v_ggsw)


def wide_sign_tests(params: CryptoParams, norm2: float):
    """(margin, variance) of every sign test of bit extraction: residue m with b bits, bit t."""
    from .wide import block_layout

    g2 = params.glwe_std ** 2
    v_pbs = variance_pbs_out(params.n, params.k, params.N, params.pbs_level, params.pbs_base_log, g2)
    v_in = max(wide_output_variance(params), g2) * max(norm2, 1.0)
    from .wide import bootstrap_theta

    v_small = (variance_keyswitch(params.big_dim, params.ks_level, params.ks_base_log, params.lwe_std ** 2)
               + variance_modswitch(params.n, params.N, bootstrap_theta(params)))
    tests = []
    for m in params.moduli:
        b, _, _, cells = block_layout(int(m))
        for t in range(b):
            margin = cells / 2.0 if t == 0 else 0.25 * (1.0 - 2.0 ** -t)
            tests.append((margin, 4.0 ** (b - 1 - t) * (v_in + t * v_pbs) + v_small))
    return tests


def estimate_wide_p_error(params: CryptoParams, norm2: float) -> float:
    """Probability that one wide lookup takes a wrong sign decision (union bound over its extracted bits)."""
    return min(1.0, sum(math.erfc(m / math.sqrt(2.0 * v)) for m, v in wide_sign_tests(params, norm2)))


def wide_lookup_flops(params: CryptoParams) -> float:
    """fp64-equivalent work of one wide lookup (all residues)."""
    from .wide import PF_PAD, total_bits

    T = total_bits(params.moduli)
    n_boot = T                     # one multi-output bootstrap per extracted bit
    M = params.N / 2
    ext = ((params.k + 1) * (params.cbs_level + 1) * 5 * M * math.log2(M) + 8 * (params.k + 1) ** 2 * params.cbs_level * M)
    logN = params.N.bit_length() - 1
    vp = len(params.moduli) * ((1 << max(T - logN, 0)) + min(T, logN)) * ext
    packing = 0.1 * KS_MAC_COST * T * params.cbs_level * (params.big_dim + PF_PAD) * params.pf_level * (params.k + 1) ** 2 * params.N
    # decompositions that keep more than 30 bits run stage 1 on a 64-bit state (measured: +15 % per bootstrap)
    wide_state = 1.15 if params.pbs_level * params.pbs_base_log > 30 else 1.0
    return (n_boot * wide_state * pbs_flops(params.n, params.k, params.N, params.pbs_level)
            + T * KS_MAC_COST * ks_macs(params.big_dim, params.n, params.ks_level) + packing + vp)


@lru_cache(maxsize=None)
def select_wide_parameters(p: int, norm2: float = 1.0, p_error: float = 1e-9,
                           shapes: tuple = WIDE_SHAPES) -> CryptoParams:
    """Cheapest parameter set for p-bit values (9..16) carried as CRT residues: every sign test of a lookup must
    hold with probability 1 - p_error / (number of extracted bits)."""
    from .wide import block_layout, crt_moduli, total_bits

    if p < 9 or p > 16:
        raise ValueError(f"wide (CRT) encoding covers 9..16 bits, not {p}")
    moduli = crt_moduli(p)
    T = total_bits(moduli)
    z = _z_for_p_error(p_error / T)
    lay = [block_layout(m) for m in moduli]
    best = None
    n_grid = np.arange(500, 1101, 12)
    lwe_var_grid = 4.0 ** np.maximum(SECURITY_SLOPE * n_grid + SECURITY_BIAS, MIN_LOG2_STD)
    ks_options = [(lv, b) for lv in range(1, 10) for b in range(1, 8) if lv * b <= 40]
    for k, N in shapes:
        kN = k * N
        g2 = 4.0 ** secure_log2_std(kN)
        v_ms = (1.0 / 12.0 + n_grid / 24.0) / (4.0 * N * N)
        # packing keyswitch on the tensor cores (base <= 2^7): smallest level whose noise hides under the bootstrap's
        for pbs_level in (1, 2, 3, 4):
            if pbs_cluster_size(N, k, pbs_level) < 0:
                continue
            for pbs_base_log in range(4, 27):
                if pbs_level * pbs_base_log > 52:
                    break
                v_pbs = variance_pbs_out(1, k, N, pbs_level, pbs_base_log, g2) * n_grid
                pf = None
                for pf_level in range(2, 9):
                    v_pf = variance_packing_keyswitch(kN, pf_level, 7, g2)
                    if v_pf <= 0.25 * float(v_pbs.min()) or pf_level == 8:
                        pf = (pf_level, 7, v_pf)
                        break
                for cbs_level in (1, 2, 3, 4):
                    if pbs_cluster_size(N, k, cbs_level) < 0:
                        continue
                    v_ms_multi = v_ms * 4.0 ** max(1, cbs_level.bit_length())      # coarse modulus switch
                    for cbs_base_log in range(2, 16):
                        if cbs_level * cbs_base_log > 40:
                            break
                        v_out = T * variance_cmux(k, N, cbs_level, cbs_base_log, v_pbs + pf[2])
                        v_in = np.maximum(v_out, g2) * max(norm2, 1.0)
                        # what the keyswitch may add: the tightest sign test decides
                        budget = np.full(n_grid.shape, np.inf)
                        for b, _, _, cells in lay:
                            for t in range(b):
                                margin = cells / 2.0 if t == 0 else 0.25 * (1.0 - 2.0 ** -t)
                                budget = np.minimum(budget, (margin / z) ** 2 - 4.0 ** (b - 1 - t) * (v_in + t * v_pbs) - v_ms_multi)
                        if not np.any(budget > 0):
                            continue
                        probe = CryptoParams(p=p, n=int(n_grid[0]), k=k, N=N, pbs_level=pbs_level, pbs_base_log=pbs_base_log,
                                             ks_level=1, ks_base_log=1, lwe_std=1.0, glwe_std=1.0, moduli=moduli,
                                             cbs_level=cbs_level, cbs_base_log=cbs_base_log, pf_level=pf[0], pf_base_log=pf[1])
                        c0 = wide_lookup_flops(probe)
                        c1 = wide_lookup_flops(CryptoParams(**{**probe.to_dict(), "n": int(n_grid[-1])}))
                        base_cost = c0 + (c1 - c0) * (n_grid - n_grid[0]) / (n_grid[-1] - n_grid[0])
                        for ks_level, ks_base_log in ks_options:
                            B2 = 4.0 ** ks_base_log
                            v_ks = kN * ks_level * (B2 + 2.0) / 12.0 * lwe_var_grid + kN / 24.0 / 4.0 ** (ks_base_log * ks_level)
                            ok = v_ks <= budget
                            if not np.any(ok):
                                continue
                            cost = np.where(ok, base_cost + T * KS_MAC_COST * kN * (ks_level - 1) * (n_grid + 1), np.inf)
                            cost = cost / wide_efficiency(k, N, pbs_level)
                            j = int(np.argmin(cost))
                            if best is None or cost[j] < best[0]:
                                best = (float(cost[j]), int(n_grid[j]), k, N, pbs_level, pbs_base_log, ks_level, ks_base_log,
                                        cbs_level, cbs_base_log, pf[0], pf[1])
    if best is None:
        raise ValueError(f"no wide parameter set reaches p_error={p_error:g} at {p} bits with 2-norm^2={norm2:g}")
    _, n, k, N, pl, pb, kl, kb, cl, cb, fl, fb = best
    return CryptoParams(p=p, n=n, k=k, N=N, pbs_level=pl, pbs_base_log=pb, ks_level=kl, ks_base_log=kb,
                        lwe_std=2.0 ** secure_log2_std(n), glwe_std=2.0 ** secure_log2_std(k * N), moduli=moduli,
                        cbs_level=cl, cbs_base_log=cb, pf_level=fl, pf_base_log=fb)


def _z_for_p_error(p_error: float) -> float:
    """z such that erfc(z / sqrt(2)) == p_error (bisection; p_error in (0, 1))."""
    p_error = min(max(p_error, 1e-300), 0.999999)
    lo, hi = 0.0, 40.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if math.erfc(mid / math.sqrt(2.0)) > p_error:
            lo = mid
        else:
            hi = mid
    return hi


def complexity(params: CryptoParams, n_pbs: int) -> float:
    """CompilationFeedback.complexity stand-in: modelled fp64-equivalent operations of a run."""
    return float(n_pbs) * (pbs_flops(params.n, params.k, params.N, params.pbs_level)
                           + KS_MAC_COST * ks_macs(params.big_dim, params.n, params.ks_level))


# Published TFHE-rs sets the curves were fitted to (provenance only; used by tests and the bench
# so that numbers are comparable with other TFHE implementations).
REFERENCE_SETS = {
    4: CryptoParams(p=4, n=742, k=1, N=2048, pbs_level=1, pbs_base_log=23, ks_level=5, ks_base_log=3,
                    lwe_std=7.069849454709433e-06, glwe_std=2.9403601535432533e-16),
    6: CryptoParams(p=6, n=864, k=1, N=8192, pbs_level=2, pbs_base_log=15, ks_level=6, ks_base_log=3,
                    lwe_std=7.57998020150446e-07, glwe_std=2.168404344971009e-19),
    8: CryptoParams(p=8, n=996, k=1, N=32768, pbs_level=2, pbs_base_log=15, ks_level=7, ks_base_log=3,
                    lwe_std=6.767666038309478e-08, glwe_std=2.168404344971009e-19),
}
