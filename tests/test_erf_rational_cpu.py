"""CPU: the two rational erf approximations the GEMM epilogues use for the exact (erf) GELU (csrc/common.cuh erf_rational, erf_rational3),
restated in numpy fp32 with the kernels' coefficients: absolute error against math.erf, and of the resulting GELU, as the header claims
(3.3e-7 / 1e-6 for the full form; 3.4e-6 on erf and 1.5e-5 on gelu(x) for the cheaper GEGLU form) - far below the bf16 resolution (4e-3
relative) of the stored activations."""
import math

import numpy as np

F = np.float32


def _horner(cs, z2):
    acc = np.full_like(z2, F(cs[0]))
    for c in cs[1:]:
        acc = (acc * z2 + F(c)).astype(np.float32)
    return acc


def erf_rational(z):
    z = np.clip(z.astype(np.float32), F(-4), F(4))
    z2 = (z * z).astype(np.float32)
    pn = _horner([2.0269792457838776e-06, 0.0002861879765987396, 0.003845315193757415, 0.05298357829451561, 0.1923242062330246, 1.128378987312317], z2)
    qn = _horner([3.7925383367110044e-05, 0.0011811800068244338, 0.015125652775168419, 0.11488588154315948, 0.5037747621536255, 1.0], z2)
    return (z * pn / qn).astype(np.float32)


def erf_rational3(z):
    z = np.clip(z.astype(np.float32), F(-3.2), F(3.2))
    z2 = (z * z).astype(np.float32)
    pn = _horner([0.0007654472137801349, 0.04346451908349991, 0.15304264426231384, 1.1283873319625854], z2)
    qn = _horner([0.009417801164090633, 0.09465143829584122, 0.4690375328063965, 1.0], z2)
    return (z * pn / qn).astype(np.float32)


def test_erf_and_gelu_errors():
    x = np.linspace(-12, 12, 480001).astype(np.float32)
    ref_erf = np.array([math.erf(v) for v in x.astype(np.float64)])
    z = (x * F(0.70710678118654752)).astype(np.float32)
    ref_erf_z = np.array([math.erf(v / math.sqrt(2.0)) for v in x.astype(np.float64)])
    gelu_ref = 0.5 * x.astype(np.float64) * (1.0 + ref_erf_z)
    assert np.abs(erf_rational(x) - ref_erf).max() < 5e-7
    assert np.abs(erf_rational3(x) - ref_erf).max() < 7e-6                      # the clamp at 3.2 costs 1 - erf(3.2) = 6e-6
    g_full = 0.5 * x * (1.0 + erf_rational(z))
    g_cheap = 0.5 * x * (1.0 + erf_rational3(z))
    assert np.abs(g_full - gelu_ref).max() < 2e-6
    assert np.abs(g_cheap - gelu_ref).max() < 4e-5
