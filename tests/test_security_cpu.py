"""128-bit security of every parameter set the backend can pick (the reference drives the absent
concrete-optimizer through `set_p_error / set_global_p_error`, concrete/numpy/compilation/server.py:102-126).

`params.secure_log2_std(n)` is the minimal noise (log2 of the std relative to q) the search allows at LWE dimension n.
It is checked against PUBLISHED 128-bit parameter sets: the `PARAM_MESSAGE_x_CARRY_y` constants of TFHE-rs 0.2
(tfhe/src/shortint/parameters/mod.rs), produced by the same concrete-optimizer and its lattice-estimator security
oracle [UPSTREAM-MEMORY: values restated from that file; three of them are this repo's REFERENCE_SETS].  Each pair is
(dimension, noise std); LWE dimensions n and GLWE dimensions k*N obey the same curve (binary secrets, q = 2^64)."""
import math
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from concrete_numpy_b200 import params as P

PUBLISHED_128_BIT = [
    # (dimension, std)                          source constant
    (656, 3.4119201269311964e-5),             # PARAM_MESSAGE_2_CARRY_0 lwe
    (678, 2.2810107419132102e-5),             # PARAM_MESSAGE_1_CARRY_0 lwe
    (684, 2.043357291315440e-5),              # PARAM_MESSAGE_1_CARRY_1 lwe
    (742, 7.069849454709433e-6),              # PARAM_MESSAGE_2_CARRY_2 lwe   (REFERENCE_SETS[4])
    (864, 7.57998020150446e-7),               # PARAM_MESSAGE_3_CARRY_3 lwe   (REFERENCE_SETS[6])
    (996, 6.767666038309478e-8),              # PARAM_MESSAGE_4_CARRY_4 lwe   (REFERENCE_SETS[8])
    (1024, 4.053919869756513e-8),             # PARAM_MESSAGE_2_CARRY_0 glwe  k = 2, N = 512
    (1280, 3.7411618952047216e-10),           # PARAM_MESSAGE_1_CARRY_0 glwe  k = 5, N = 256
    (1536, 3.4525330484572114e-12),           # PARAM_MESSAGE_1_CARRY_1 glwe  k = 3, N = 512
    (2048, 2.9403601535432533e-16),           # PARAM_MESSAGE_2_CARRY_2 glwe  k = 1, N = 2048
    (8192, 2.168404344971009e-19),            # PARAM_MESSAGE_3_CARRY_3 glwe  (the 2^-62 floor of a 64-bit torus)
    (32768, 2.168404344971009e-19),           # PARAM_MESSAGE_4_CARRY_4 glwe
]


def test_security_curve_is_never_below_a_published_128_bit_point():
    for dim, std in PUBLISHED_128_BIT:
        ours, published = P.secure_log2_std(dim), math.log2(std)
        assert ours >= published - 1e-9, (dim, ours, published)          # at least as much noise: at least as secure
        assert ours <= published + 0.05, (dim, ours, published)          # ... and no margin wasted
    # monotone: a larger dimension never demands more noise
    vals = [P.secure_log2_std(n) for n in range(400, 40000, 97)]
    assert all(a >= b for a, b in zip(vals, vals[1:]))


def _on_or_above_curve(p):
    tol = 1e-9
    assert math.log2(p.glwe_std) >= P.secure_log2_std(p.big_dim) - tol, ("glwe", p)
    if not p.linear_only:
        assert math.log2(p.lwe_std) >= P.secure_log2_std(p.n) - tol, ("lwe", p)


@pytest.mark.parametrize("bits", list(range(1, 9)))
@pytest.mark.parametrize("norm2,p_error", [(1.0, 1e-6), (64.0, 1e-9), (4096.0, 1e-4)])
def test_every_selected_native_parameter_set_is_on_or_above_the_curve(bits, norm2, p_error):
    try:
        p = P.select_parameters(bits, norm2, p_error)
    except ValueError:
        pytest.skip("no feasible set for this corner (the reference's optimizer reports the same way)")
    _on_or_above_curve(p)
    assert P.estimate_p_error(p, norm2) <= p_error


@pytest.mark.parametrize("bits", list(range(9, 17)))
def test_every_selected_wide_parameter_set_is_on_or_above_the_curve(bits):
    p = P.select_wide_parameters(bits, 1.0, 1e-6)
    _on_or_above_curve(p)
    assert P.estimate_wide_p_error(p, 1.0) <= 1e-6


def test_linear_only_and_reference_sets_are_on_or_above_the_curve():
    for bits in (12, 20, 31, 40):
        _on_or_above_curve(P.select_linear_parameters(bits, 16.0, 1e-6))
    # the pinned reference sets ARE published points (the curve sits a hair, < 0.01 bit, above them)
    published = dict(PUBLISHED_128_BIT)
    for p in P.REFERENCE_SETS.values():
        assert published[p.n] == p.lwe_std and published[p.big_dim] == p.glwe_std
