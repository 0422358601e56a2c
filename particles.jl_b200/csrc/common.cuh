// common.cuh -- shared device helpers: 128-bit fixed point, block reductions, layouts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PF_MAX_D 8
#define PF_MAX_SLOTS 64          // k_max + 1 <= 64
#define PF_INF_BITS 0x7FF0000000000000ull

typedef unsigned long long u64;
typedef unsigned int u32;

// ------------------------------------------------------------------ 128-bit unsigned
struct u128 {
    u64 lo, hi;
};
__host__ __device__ __forceinline__ u128 mk128(u64 lo, u64 hi = 0) { u128 r; r.lo = lo; r.hi = hi; return r; }
__host__ __device__ __forceinline__ u128 add128(u128 a, u128 b)
{
    u128 r;
    r.lo = a.lo + b.lo;
    r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
    return r;
}
__host__ __device__ __forceinline__ u128 sub128(u128 a, u128 b)
{
    u128 r;
    r.lo = a.lo - b.lo;
    r.hi = a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull);
    return r;
}
__host__ __device__ __forceinline__ bool lt128(u128 a, u128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__host__ __device__ __forceinline__ bool le128(u128 a, u128 b) { return !lt128(b, a); }
__host__ __device__ __forceinline__ bool isz128(u128 a) { return (a.lo | a.hi) == 0; }
__host__ __device__ __forceinline__ u64 mulhi64(u64 a, u64 b)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
// a (128 bit) * b (64 bit), low 128 bits (callers guarantee no overflow)
__host__ __device__ __forceinline__ u128 mul128_64(u128 a, u64 b)
{
    u128 r;
    r.lo = a.lo * b;
    r.hi = mulhi64(a.lo, b) + a.hi * b;
    return r;
}
__host__ __device__ __forceinline__ u128 shr128(u128 a, int s)
{
    if (s <= 0) return a;
    if (s >= 128) return mk128(0, 0);
    if (s >= 64) return mk128(a.hi >> (s - 64), 0);
    return mk128((a.lo >> s) | (a.hi << (64 - s)), a.hi >> s);
}
__host__ __device__ __forceinline__ u128 shl128(u128 a, int s)
{
    if (s <= 0) return a;
    if (s >= 128) return mk128(0, 0);
    if (s >= 64) return mk128(0, a.lo << (s - 64));
    return mk128(a.lo << s, (a.hi << s) | (a.lo >> (64 - s)));
}
// round-to-nearest conversion (hi*2^64 is exact for hi < 2^53, one rounding in the add)
__host__ __device__ __forceinline__ double to_double128(u128 a)
{
    return (double)a.hi * 18446744073709551616.0 + (double)a.lo;
}

// Fixed point: q = floor(v * 2^s) for a finite v >= 0 given by its bit pattern.
// Callers guarantee v * 2^s < 2^127.
__host__ __device__ __forceinline__ u128 fx128(u64 bits, int s)
{
    int e = (int)(bits >> 52) & 0x7FF;
    u64 m = bits & 0x000FFFFFFFFFFFFFull;
    if (e == 0) e = 1; else m |= 0x0010000000000000ull;
    int sh = e - 1075 + s;                 // v = m * 2^(e-1075)
    if (sh >= 0) return shl128(mk128(m, 0), sh);
    if (sh <= -64) return mk128(0, 0);
    return mk128(m >> (-sh), 0);
}
// Fast path of fx128 for callers that guarantee v * 2^s < 2^71 (shift of the 53-bit mantissa
// by at most 18 to the left): the hot loops of the threshold search and the survivor scan.
__host__ __device__ __forceinline__ u128 fx70(u64 bits, int s)
{
    int e = (int)(bits >> 52);                      // sign bit is 0 for weights
    u64 m = bits & 0x000FFFFFFFFFFFFFull;
    if (e == 0) e = 1; else m |= 0x0010000000000000ull;
    const int sh = e - 1075 + s;
    u128 r;
    if (sh > 0) { r.lo = m << sh; r.hi = m >> (64 - sh); }
    else { r.lo = sh > -64 ? m >> (-sh) : 0ull; r.hi = 0ull; }
    return r;
}

// a / n and a % n for a 128-bit a and a 32-bit n > 0 (schoolbook, 32-bit digits)
__host__ __device__ __forceinline__ u128 divmod128_32(u128 a, unsigned int n, unsigned int *rem)
{
    unsigned int d[4] = {(unsigned int)(a.hi >> 32), (unsigned int)a.hi, (unsigned int)(a.lo >> 32), (unsigned int)a.lo};
    unsigned int q[4];
    u64 r = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u64 cur = (r << 32) | d[k];
        q[k] = (unsigned int)(cur / n);
        r = cur % n;
    }
    *rem = (unsigned int)r;
    return mk128(((u64)q[2] << 32) | q[3], ((u64)q[0] << 32) | q[1]);
}

// unbiased exponent E with 2^E <= v < 2^(E+1) (subnormals: -1022)
__host__ __device__ __forceinline__ int exp_of_bits(u64 bits)
{
    int e = (int)(bits >> 52) & 0x7FF;
    return (e == 0 ? 1 : e) - 1023;
}

// number of comb teeth strictly below X:  #{k >= 0 : U + k T < X}  (T > 0)
__host__ __device__ inline u64 teeth_below(u128 X, u128 U, u128 T)
{
    if (!lt128(U, X)) return 0;
    u128 R = sub128(sub128(X, U), mk128(1, 0));      // X - U - 1 >= 0
    // floor(R / T) + 1 ; quotient < 2^40 by construction, FP estimate + exact correction
    double est = to_double128(R) / to_double128(T);
    u64 q = est >= 1.0 ? (u64)est : 0;
    if (q > 0) q -= 1;
    u128 qt = mul128_64(T, q);
    while (lt128(R, qt)) { q--; qt = sub128(qt, T); }           // q*T <= R
    u128 rem = sub128(R, qt);
    while (le128(T, rem)) { q++; rem = sub128(rem, T); }        // rem < T
    return q + 1;
}

// 128 x 128 -> 256 bit product, little-endian limbs
__host__ __device__ __forceinline__ void mul128x128(u128 a, u128 b, u64 o[4])
{
    u64 p00l = a.lo * b.lo, p00h = mulhi64(a.lo, b.lo);
    u64 p01l = a.lo * b.hi, p01h = mulhi64(a.lo, b.hi);
    u64 p10l = a.hi * b.lo, p10h = mulhi64(a.hi, b.lo);
    u64 p11l = a.hi * b.hi, p11h = mulhi64(a.hi, b.hi);
    o[0] = p00l;
    u64 c = 0, s = p00h + p01l; c += s < p00h; u64 s2 = s + p10l; c += s2 < s; o[1] = s2;
    u64 c2 = 0; s = p01h + p10h; c2 += s < p01h; s2 = s + p11l; c2 += s2 < s; u64 s3 = s2 + c; c2 += s3 < s2; o[2] = s3;
    o[3] = p11h + c2;
}
__host__ __device__ __forceinline__ bool gt256(const u64 a[4], const u64 b[4])
{
    for (int k = 3; k >= 0; k--) { if (a[k] != b[k]) return a[k] > b[k]; }
    return false;
}

// ------------------------------------------------------------------ comb teeth in running-mass space
// Tooth k of the systematic comb sits at U + k T (in units of 1/n of the fixed-point mass); it is
// passed by a running mass S iff U + k T < n S iff Q_k < S with (Q_k, R_k) = divmod(U + k T, n).
// From tooth to tooth (Q, R) steps by (Tq, Tr) = divmod(T, n).
struct ToothBase {
    u128 Q0;                 // quotient of the first tooth not passed by the tile's starting mass S0
    unsigned int R0;         // its remainder
    u64 t0;                  // its index = number of teeth passed by S0
    double inv;              // n / T: teeth per unit of running mass (estimate only)
};
// once per tile (the one 128-bit division)
__host__ __device__ inline ToothBase tooth_base(u128 S0, u64 n, u128 U, u128 T)
{
    ToothBase b;
    b.t0 = teeth_below(mul128_64(S0, n), U, T);
    b.Q0 = divmod128_32(add128(U, mul128_64(T, b.t0)), (unsigned int)n, &b.R0);
    b.inv = (double)n / to_double128(T);
    return b;
}
__host__ __device__ __forceinline__ void tooth_step(u128 &Q, unsigned int &R, u128 Tq, unsigned int Tr, unsigned int n32)
{
    Q = add128(Q, Tq);
    R += Tr;
    if (R >= n32) { R -= n32; Q = add128(Q, mk128(1, 0)); }
}
// number k of teeth of the tile passed by S >= S0 (so t0 + k teeth are passed in all), with (Q, R) of
// tooth t0 + k: k from a floating-point estimate, corrected exactly in both directions
// x / n and x % n for x < 2^63 and a 32-bit n > 0 whose quotient is small (< 2^20): a double-precision
// estimate corrected by one step either way (two 64-bit divisions are ~200 instructions on the GPU)
__host__ __device__ __forceinline__ u64 divmod_small(u64 x, unsigned int n32, unsigned int *rem)
{
    u64 q = (u64)((double)x / (double)n32);
    long long r = (long long)(x - q * (u64)n32);
    if (r < 0) { q--; r += n32; }
    else if (r >= (long long)n32) { q++; r -= n32; }
    *rem = (unsigned int)r;
    return q;
}
__host__ __device__ __forceinline__ u64 tooth_seek(u128 Q0, unsigned int R0, double inv, u128 S, u128 Tq, unsigned int Tr,
                                                   unsigned int n32, u128 &Q, unsigned int &R)
{
    u64 k = 0; Q = Q0; R = R0;
    if (lt128(Q0, S)) {
        k = (u64)(to_double128(sub128(S, Q0)) * inv);
        u64 x = (u64)R0 + k * (u64)Tr;
        if (k < (1ull << 20)) { unsigned int rr; const u64 qq = divmod_small(x, n32, &rr); Q = add128(add128(Q0, mul128_64(Tq, k)), mk128(qq, 0)); R = rr; }
        else { Q = add128(add128(Q0, mul128_64(Tq, k)), mk128(x / n32, 0)); R = (unsigned int)(x % n32); }
        if (lt128(Q, S)) {
            do { k++; tooth_step(Q, R, Tq, Tr, n32); } while (lt128(Q, S));
        } else {
            while (k > 0) {
                x = (u64)R0 + (k - 1) * (u64)Tr;
                unsigned int rp;
                const u64 qp = (k - 1 < (1ull << 20)) ? divmod_small(x, n32, &rp) : (rp = (unsigned int)(x % n32), x / n32);
                u128 Qp = add128(add128(Q0, mul128_64(Tq, k - 1)), mk128(qp, 0));
                if (lt128(Qp, S)) break;
                k--; Q = Qp; R = rp;
            }
        }
    }
    return k;
}

// The survivor scan's fast walk asks "has the running low mass of this particle passed the next tooth?" in double
// precision: cum = sum of the items' values scaled by 2^scale (exact doubles; the integer masses floor(v 2^scale) differ
// from them by < 1 each, at most 64 items), Dd = (double)(Q - S) (one rounding).  The answer is taken only when it
// clears the error bound by a wide margin; otherwise the particle is redone in exact 128-bit arithmetic.
//   walk_compare: +1 passed, -1 not passed, 0 too close to tell.
//   walk_reject_below(Dd): a bound below which cum is certainly "not passed" (cum < Dd there, so the margin
//   64 + (cum + Dd) 2^-45 is < 64 + Dd 2^-44; the bound uses twice that) -- one comparison per item in the common case.
__host__ __device__ __forceinline__ int walk_compare(double cum, double Dd)
{
    const double diff = cum - Dd;
    const double tol = 64.0 + (cum + Dd) * 2.8421709430404007e-14;          // 2^-45 relative + the floors
    return diff > tol ? 1 : (diff >= -tol ? 0 : -1);
}
__host__ __device__ __forceinline__ double walk_reject_below(double Dd)
{
    return Dd - (128.0 + Dd * 1.1368683772161603e-13);                       // 2^-43
}

// ------------------------------------------------------------------ warp / block helpers
__device__ __forceinline__ u128 shfl_up128(u128 v, int d)
{
    u128 r;
    r.lo = __shfl_up_sync(0xffffffffu, v.lo, d);
    r.hi = __shfl_up_sync(0xffffffffu, v.hi, d);
    return r;
}
__device__ __forceinline__ u128 shfl_xor128(u128 v, int d)
{
    u128 r;
    r.lo = __shfl_xor_sync(0xffffffffu, v.lo, d);
    r.hi = __shfl_xor_sync(0xffffffffu, v.hi, d);
    return r;
}
__device__ __forceinline__ u128 shfl_xor_bcast128(u128 v, int src_lane)
{
    u128 r;
    r.lo = __shfl_sync(0xffffffffu, v.lo, src_lane);
    r.hi = __shfl_sync(0xffffffffu, v.hi, src_lane);
    return r;
}
__device__ __forceinline__ u128 warp_sum128(u128 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = add128(v, shfl_xor128(v, d));
    return v;
}
__device__ __forceinline__ u64 warp_sum64(u64 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ u64 warp_max64(u64 v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        u64 o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o > v ? o : v;
    }
    return v;
}

// 128-bit global accumulation through two 64-bit atomics on 3 limbs would lose carries;
// instead accumulate four 32-bit limbs in four u64 cells (exact for < 2^32 contributions).
struct acc128 {
    u64 l[4];
};
__device__ __forceinline__ void atomic_add_acc(acc128 *a, u128 v)
{
    if (v.lo & 0xffffffffull) atomicAdd(&a->l[0], v.lo & 0xffffffffull);
    if (v.lo >> 32) atomicAdd(&a->l[1], v.lo >> 32);
    if (v.hi & 0xffffffffull) atomicAdd(&a->l[2], v.hi & 0xffffffffull);
    if (v.hi >> 32) atomicAdd(&a->l[3], v.hi >> 32);
}
__host__ __device__ __forceinline__ u128 acc_value(const volatile acc128 *a)
{
    u128 r = mk128(a->l[0], 0);
    r = add128(r, shl128(mk128(a->l[1], 0), 32));
    r = add128(r, mk128(0, a->l[2]));
    r = add128(r, mk128(0, a->l[3] << 32));
    return r;
}

// Checked build (-DPF_CHECKED, `PF_CHECKED=1 python build.py`): index invariants of the gathers / scatters are
// asserted on the device; a violation prints its site and traps (every later CUDA call then fails, so the tests fail
// loudly).  compute-sanitizer is closed on the GPU pool this was developed on; the parity suite run against the checked
// build (profiles/r2/checked_build.log) is the substitute.
#ifdef PF_CHECKED
#include <cstdio>
#define PF_ASSERT(cond)                                                                                          \
    do {                                                                                                         \
        if (!(cond)) {                                                                                           \
            printf("PF_ASSERT failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                                            \
        }                                                                                                        \
    } while (0)
#else
#define PF_ASSERT(cond) ((void)0)
#endif

#define CUDA_TRY(x)                                                                      \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) return set_cuda_error(h, e_, #x, __LINE__);               \
    } while (0)
