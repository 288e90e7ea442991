"""CPU checks of two arithmetic building blocks of the CUDA path, restated in numpy from the constants in the sources:
the multiply-high division used for the window index arithmetic (loaders.cuh FastDiv) and the single-polynomial exact
GELU of the tcgen05 epilogue (gemm.cuh gelu_erf2)."""
import math
import os
import random
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "efficient-speech-codec_b200", "csrc")


def fastdiv_make(d):
    if d <= 1:
        return 0, 0
    s = 0
    while (1 << s) < d:
        s += 1
    return ((1 << (31 + s)) + d - 1) // d, s - 1


def test_fastdiv_is_exact_below_2_31():
    rnd = random.Random(0)
    divisors = list(range(1, 2000)) + [2 ** k for k in range(1, 31)] + [2 ** k + 1 for k in range(1, 30)] + \
        [2 ** k - 1 for k in range(2, 31)] + [rnd.randrange(1, 2 ** 31) for _ in range(500)]
    for d in divisors:
        magic, shift = fastdiv_make(d)
        assert magic < 2 ** 32
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1] + [rnd.randrange(0, 2 ** 31) for _ in range(50)] + \
            [k * d - 1 for k in range(1, 20)] + [k * d for k in range(1, 20)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                q = (((n * magic) >> 32) >> shift) if magic else n
                assert q == n // d, (d, n)


def test_fastdiv_source_matches_this_restatement():
    src = open(os.path.join(CSRC, "loaders.cuh")).read()
    assert "(1ull << (31 + s)) + d - 1) / d" in src and "f.shift = s - 1" in src and "__umulhi(n, magic) >> shift" in src


def gelu_coefficients():
    src = open(os.path.join(CSRC, "gemm.cuh")).read()
    body = src[src.index("void gelu_erf2("):]
    body = body[:body.index("#undef ESCB_C2")]
    vals = [float(v.rstrip("f")) for v in re.findall(r"ESCB_C2\((-?[0-9.e+-]+f)\)", body)]
    assert len(vals) == 11 and vals[-1] == -1.0        # c9 .. c0 of R, then the -1 of Q = -1 + a R(a)
    return vals


def test_gelu_polynomial_accuracy():
    """x * (x < 0 ? e : 1 - e), e = 2^Q(min(|x|, 6)), evaluated with fp32 FMA semantics, against float64 erf GELU."""
    from scipy.special import erf
    co = [np.float32(v) for v in gelu_coefficients()]
    x = np.concatenate([np.linspace(-9, 9, 1_200_001), [0.0, -0.0, 1e-20, -1e-20, 30.0, -30.0]]).astype(np.float32)
    a = np.minimum(np.abs(x), np.float32(6.0))

    def fma(p, q, r):
        return (p.astype(np.float64) * q.astype(np.float64) + np.float64(r)).astype(np.float32)

    acc = np.full_like(a, co[0])
    for c in co[1:]:
        acc = fma(acc, a, c)
    e = np.exp2(acc.astype(np.float64)).astype(np.float32)
    y = (x * np.where(x < 0, e, (np.float32(1) - e).astype(np.float32))).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1.0 + erf(x.astype(np.float64) / math.sqrt(2.0)))
    err = np.abs(y - ref)
    assert err.max() <= 5e-7, (err.max(), x[err.argmax()])
    assert err[np.abs(x) < 2].max() <= 2e-7
    # the erff form it replaced, same comparison (documents that the two are equivalent at fp32 level)
    y_erf = (np.float32(0.5) * x * (np.float32(1) + erf((x * np.float32(0.70710678)).astype(np.float64)).astype(np.float32))).astype(np.float32)
    assert err.max() <= 1.2 * np.abs(y_erf - ref).max() + 1e-9
