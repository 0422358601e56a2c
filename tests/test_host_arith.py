"""The 128-bit fixed-point arithmetic and the comb-tooth stepping of csrc/common.cuh, compiled
for the host (tests/host/host_arith.cu, nvcc) and checked against Python's exact integers:
floor(v 2^s), 128/32 division, teeth counting, and the survivor scan's per-tile / per-thread
tooth seek against a direct evaluation of sample_stratified! (src/fearnhead.jl:97-114) in
rational arithmetic.  No GPU needed: the functions are __host__ __device__."""
import ctypes as C
import os
import shutil
import subprocess
from fractions import Fraction

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "host_arith.cu")
OUT = os.path.join(HERE, "host", "_build", "libhost_arith.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
pytestmark = pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")

U64 = C.c_uint64
M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def H():
    dep = os.path.join(HERE, "..", "particles.jl_b200", "csrc", "common.cuh")
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(dep)):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([NVCC if os.path.exists(NVCC) else "nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.h_teeth_below.restype = U64
    L.h_to_double128.restype = C.c_double
    L.h_comb.restype = C.c_longlong
    return L


def _bits(v):
    return int(np.float64(v).view(np.uint64))


def _pair(x):
    return (U64 * 2)(x & M64, x >> 64)


def _val(a):
    return int(a[0]) | (int(a[1]) << 64)


def _exact_fx(v, s):
    return int(Fraction(float(v)) * Fraction(2) ** s // 1)


def test_fixed_point_conversion_is_floor(H):
    rng = np.random.default_rng(0)
    vals = list(np.exp(rng.standard_normal(300) * 30)) + [0.0, 5e-324, 2.2250738585072014e-308, 1.0, 0.5, 3.0, 1e-300, 1e300]
    out = (U64 * 2)()
    for v in vals:
        e = H.h_exp_of_bits(U64(_bits(v)))
        assert (e == -1022 and v < 2.2250738585072014e-308) or 2.0 ** e <= v < 2.0 ** (e + 1)
        for s in (69 - e, 60 - e, 30 - e, -e, -e - 60, 52 - e, 99 - e):
            if v * 2.0 ** min(s, 900) >= 2.0 ** 126:
                continue
            H.h_fx128(U64(_bits(v)), s, out)
            assert _val(out) == _exact_fx(v, s), (v, s)
            if s <= 69 - e:                     # fx70's contract: v 2^s < 2^71
                H.h_fx70(U64(_bits(v)), s, out)
                assert _val(out) == _exact_fx(v, s), (v, s)


def test_shifts_multiply_divide(H):
    rng = np.random.default_rng(1)
    out = (U64 * 4)()
    for _ in range(500):
        a = int(rng.integers(0, 1 << 62)) << int(rng.integers(0, 60)) | int(rng.integers(0, 1 << 30))
        a &= (1 << 120) - 1
        n = int(rng.integers(1, 1 << 31))
        H.h_divmod128_32(U64(a & M64), U64(a >> 64), C.c_uint(n), out)
        assert (_val(out), int(out[2])) == divmod(a, n)
        b = int(rng.integers(0, 1 << 30))
        if a * b < 1 << 128:
            H.h_mul128_64(U64(a & M64), U64(a >> 64), U64(b), out)
            assert _val(out) == a * b
        s = int(rng.integers(0, 130))
        H.h_shift128(U64(a & M64), U64(a >> 64), s, out)
        assert _val(out) == (a << s) & ((1 << 128) - 1) and (int(out[2]) | int(out[3]) << 64) == a >> s
        d = H.h_to_double128(U64(a & M64), U64(a >> 64))
        assert abs(Fraction(d) - a) <= Fraction(a, 1 << 52)


def test_256_bit_product_and_compare(H):
    # Chen-Liu rejuvenation: cum[mid] * (N 2^53) > (i 2^53 + mu) * Qtot decided in 256 bits
    import random
    rnd = random.Random(3)
    prod = (U64 * 4)()
    for _ in range(2000):
        a, b, c, d = (rnd.getrandbits(rnd.choice((20, 64, 100, 128))) for _ in range(4))
        got = H.h_mul_gt256(_pair(a), _pair(b), _pair(c), _pair(d), prod)
        assert sum(int(prod[k]) << (64 * k) for k in range(4)) == a * b
        assert got == (1 if a * b > c * d else 0)
    assert H.h_mul_gt256(_pair(7), _pair(9), _pair(9), _pair(7), prod) == 0      # equal products are not greater


def _teeth_below(X, U, T):
    return 0 if X <= U else (X - U - 1) // T + 1


def test_teeth_below(H):
    import random
    rnd = random.Random(2)
    for _ in range(2000):
        T = rnd.randrange(1, 1 << 62) << rnd.randrange(0, 36)
        U = rnd.randrange(0, T)
        k = rnd.randrange(0, 1 << 24)                   # (the quotient is < 2^40 by construction in the kernels)
        for X in (U + k * T, U + k * T + 1, U + k * T - 1 if U + k * T > 0 else 0, rnd.randrange(0, 1 << 62), U, 0):
            if X >= 1 << 127:
                continue
            assert H.h_teeth_below(_pair(X), _pair(U), _pair(T)) == _teeth_below(X, U, T)


@pytest.mark.parametrize("inv_scale", [1.0, 1.37, 0.61, 0.0, 40.0])
@pytest.mark.parametrize("seed,m,n,tile,per,spread", [(0, 5000, 700, 256, 1, 3.0), (1, 4000, 1500, 256, 7, 1.0), (2, 3000, 40, 64, 17, 8.0),
                                                      (3, 6000, 5999, 512, 3, 0.3), (4, 2000, 1, 256, 10, 2.0), (5, 1, 1, 256, 1, 1.0)])
def test_comb_tiles_and_threads_equal_direct_comb(H, seed, m, n, tile, per, spread, inv_scale):
    """Picks of the tiled tooth seek == teeth of {U + k T} falling strictly below n S_i but not below
    n S_{i-1} (sample_stratified! in exact arithmetic), for every item, with the right tooth index.
    inv_scale != 1 spoils the floating-point estimate of the seek: the exact correction must absorb it."""
    rng = np.random.default_rng(seed)
    w = np.exp(rng.standard_normal(m) * spread)
    kept = (rng.random(m) < 0.3).astype(np.uint8)
    if kept.all():
        kept[0] = 0
    scale = 69 - int(np.floor(np.log2(w.max())))
    q = [_exact_fx(v, scale) for v in w]
    T = sum(qi for qi, k in zip(q, kept) if not k)
    U = int(Fraction(float(rng.random())) * T)          # floor(u T)
    qlo = np.array([x & M64 for x in q], dtype=np.uint64)
    qhi = np.array([x >> 64 for x in q], dtype=np.uint64)
    copies = np.zeros(m, dtype=np.uint8)
    first = np.zeros(m, dtype=np.int64)
    tot = H.h_comb(qlo.ctypes.data_as(C.POINTER(U64)), qhi.ctypes.data_as(C.POINTER(U64)), kept.ctypes.data_as(C.POINTER(C.c_ubyte)),
                   C.c_longlong(m), U64(n), _pair(U), _pair(T), C.c_longlong(tile), C.c_longlong(per),
                   copies.ctypes.data_as(C.POINTER(C.c_ubyte)), first.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_double(inv_scale))
    S = 0
    exp_copies = np.zeros(m, dtype=np.int64)
    exp_first = np.full(m, -1, dtype=np.int64)
    for i in range(m):
        if kept[i]:
            continue
        before = _teeth_below(n * S, U, T)
        S += q[i]
        after = _teeth_below(n * S, U, T)
        exp_copies[i] = after - before
        if after > before:
            exp_first[i] = before
    assert tot == exp_copies.sum() == n                 # every tooth lands on exactly one item
    assert np.array_equal(copies, exp_copies)
    assert np.array_equal(first, exp_first)


def _first_tooth(S, n, U, T):
    """(k, Q_k, R_k) of the first comb tooth not passed by the running mass S: tooth k is passed iff floor((U + k T) / n) < S."""
    k = _teeth_below(n * S, U, T)
    return (k,) + divmod(U + k * T, n)


@pytest.mark.parametrize("seed,near", [(0, False), (1, False), (2, True), (3, True)])
def test_fast_walk_decides_like_the_exact_walk_or_hands_over(H, seed, near):
    """The survivor scan's fast walk (double precision, walk_compare / walk_reject_below of csrc/common.cuh) against the
    exact integer walk: whenever it answers, its teeth land in the same slots; with the running mass steered to within a
    few units of a tooth (`near`) it may only answer correctly or hand the particle over -- and the one-comparison
    rejection bound never changes an answer."""
    import random
    rnd = random.Random(100 + seed)
    H.h_fast_walk.restype = C.c_int
    answered = handed = 0
    for _ in range(1500):
        c = rnd.randrange(1, 26)
        top = rnd.uniform(-30.0, 30.0)
        w = [2.0 ** (top - rnd.expovariate(0.25)) * rnd.uniform(1.0, 2.0) for _ in range(c)]
        scale = 69 - int(np.floor(np.log2(max(w)))) - rnd.randrange(0, 12)      # (the bracket's top is at or above the items)
        q = [_exact_fx(v, scale) for v in w]
        n = rnd.randrange(1, 1 << 24)
        tot_p = sum(q)
        # the comb: spacing c = T / n around the particle's own mass, so that 0 .. 3 teeth fall inside it
        T = max(n, int(tot_p * n * rnd.uniform(0.3, 4.0)))
        S = rnd.randrange(0, 1 << 80)
        U = rnd.randrange(0, T)
        if near:
            # put a tooth within a few units of the running mass after a random item
            j = rnd.randrange(0, c)
            target = S + sum(q[:j + 1]) + rnd.choice([-1, 1]) * rnd.choice([rnd.randrange(0, 4), rnd.randrange(40, 300), rnd.randrange(10 ** 3, 10 ** 6)])
            k0 = max(0, (n * target - U) // T)
            U = (n * target - k0 * T) % T if n * target >= k0 * T else U
        k, Q, R = _first_tooth(S, n, U, T)
        Tq, Tr = divmod(T, n)
        bits = (U64 * c)(*[_bits(v) for v in w])
        res = []
        for use_bound in (1, 0):
            slots = (C.c_int * 64)()
            nt = H.h_fast_walk(bits, c, scale, _pair(S), _pair(Q), C.c_uint(R), _pair(Tq), C.c_uint(Tr), C.c_uint(n), slots, use_bound)
            res.append((nt, list(slots[:max(nt, 0)])))
        assert res[0] == res[1]                          # the rejection bound is only a shortcut
        nt, slots = res[0]
        # exact walk
        exp, run, kk = [], S, k
        for j in range(c):
            run += q[j]
            while (U + kk * T) // n < run:
                exp.append(j); kk += 1
        if nt < 0:
            handed += 1
        else:
            answered += 1
            assert slots == exp
    assert answered > 0
    if not near:
        assert handed < 0.02 * (answered + handed)       # the exact path is the rare case
