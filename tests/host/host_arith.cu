// Host-side harness for the __host__ __device__ integer arithmetic of csrc/common.cuh: the same
// source the kernels compile, built for the CPU so that tests/test_host_arith.py can check it
// against Python's exact integers without a GPU.  Test infrastructure only.
#include <cstring>
#include "../../particles.jl_b200/csrc/common.cuh"

extern "C" {

void h_fx128(u64 bits, int s, u64 *out) { u128 r = fx128(bits, s); out[0] = r.lo; out[1] = r.hi; }
void h_fx70(u64 bits, int s, u64 *out) { u128 r = fx70(bits, s); out[0] = r.lo; out[1] = r.hi; }
void h_divmod128_32(u64 lo, u64 hi, unsigned int n, u64 *out)
{
    unsigned int rem;
    u128 q = divmod128_32(mk128(lo, hi), n, &rem);
    out[0] = q.lo; out[1] = q.hi; out[2] = rem;
}
u64 h_teeth_below(const u64 *X, const u64 *U, const u64 *T) { return teeth_below(mk128(X[0], X[1]), mk128(U[0], U[1]), mk128(T[0], T[1])); }
void h_mul128_64(u64 lo, u64 hi, u64 b, u64 *out) { u128 r = mul128_64(mk128(lo, hi), b); out[0] = r.lo; out[1] = r.hi; }
double h_to_double128(u64 lo, u64 hi) { return to_double128(mk128(lo, hi)); }
int h_exp_of_bits(u64 bits) { return exp_of_bits(bits); }
void h_shift128(u64 lo, u64 hi, int s, u64 *out)
{
    u128 l = shl128(mk128(lo, hi), s), r = shr128(mk128(lo, hi), s);
    out[0] = l.lo; out[1] = l.hi; out[2] = r.lo; out[3] = r.hi;
}

int h_mul_gt256(const u64 *a, const u64 *b, const u64 *c, const u64 *d, u64 *prod)
{
    u64 x[4], y[4];
    mul128x128(mk128(a[0], a[1]), mk128(b[0], b[1]), x);
    mul128x128(mk128(c[0], c[1]), mk128(d[0], d[1]), y);
    for (int k = 0; k < 4; k++) prod[k] = x[k];
    return gt256(x, y) ? 1 : 0;
}

// The fast walk of one particle (walk_particle of csrc/fearnhead.cuh, the double-precision path) on the host, with the
// kernel's own decision helpers walk_compare / walk_reject_below / tooth_step: `c` low items given by their bit patterns,
// S = running fixed-point mass before the particle, (Q, R) = the first tooth not passed by S, (Tq, Tr, n32) the comb step.
// Returns -1 when the walk would hand the particle to the exact path (a comparison too close to tell, a subnormal or an
// item whose scaled value is not a normal double), else the number of teeth found; tooth_slot[k] = slot of tooth k.
int h_fast_walk(const u64 *bits, int c, int scale, const u64 *Sv, const u64 *Qv, unsigned int R, const u64 *Tqv, unsigned int Tr,
                unsigned int n32, int *tooth_slot, int use_reject_bound)
{
    const u128 S = mk128(Sv[0], Sv[1]), Tq = mk128(Tqv[0], Tqv[1]);
    u128 Q = mk128(Qv[0], Qv[1]);
    double cum = 0.0;
    double Dd = to_double128(sub128(Q, S));
    double Dlo = walk_reject_below(Dd);
    const long long sadd = (long long)scale << 52;
    int nt = 0;
    for (int j = 0; j < c; j++) {
        const u64 b = bits[j];
        if (b == 0) continue;
        const int eb = (int)(b >> 52) + scale;
        if ((b >> 52) == 0 || (unsigned int)(eb - 1) >= 2045u) return -1;
        long long sb = (long long)b + sadd;
        double v; memcpy(&v, &sb, 8);
        cum += v;
        if (use_reject_bound && cum < Dlo) continue;
        for (;;) {
            const int passed = walk_compare(cum, Dd);
            if (passed > 0) {
                tooth_slot[nt++] = j;
                tooth_step(Q, R, Tq, Tr, n32);
                Dd = to_double128(sub128(Q, S));
                Dlo = walk_reject_below(Dd);
                continue;
            }
            if (passed == 0) return -1;
            break;
        }
    }
    return nt;
}

// The survivor scan's comb (k_scan_apply) on the host: `m` items with fixed-point masses q (lo, hi)
// and a kept flag, traversed in order; tiles of `tile` items, threads of `per` items each.  Every
// tile starts from its exclusive prefix through tooth_base, every thread through tooth_seek, items
// step with tooth_step -- the kernel's code path.  Writes for each item the number of comb teeth it
// receives (0/1 for a valid comb) and the index t of its first tooth; returns the total.
long long h_comb(const u64 *qlo, const u64 *qhi, const unsigned char *kept, long long m, u64 n, const u64 *Uv, const u64 *Tv,
                 long long tile, long long per, unsigned char *copies, long long *first_tooth, double inv_scale)
{
    const u128 U = mk128(Uv[0], Uv[1]), T = mk128(Tv[0], Tv[1]);
    unsigned int Tr;
    const u128 Tq = divmod128_32(T, (unsigned int)n, &Tr);
    const unsigned int n32 = (unsigned int)n;
    u128 run = mk128(0, 0);                       // exclusive prefix of the low masses
    long long total = 0;
    for (long long t0 = 0; t0 < m; t0 += tile) {
        ToothBase b = tooth_base(run, n, U, T);
        b.inv *= inv_scale;            // tests: a wrong estimate must be corrected exactly, in either direction
        u128 S_thread = run;
        for (long long a = t0; a < t0 + tile && a < m; a += per) {
            u128 Q; unsigned int R;
            u128 S = S_thread;
            u64 t = b.t0 + tooth_seek(b.Q0, b.R0, b.inv, S, Tq, Tr, n32, Q, R);
            for (long long i = a; i < a + per && i < t0 + tile && i < m; i++) {
                copies[i] = 0; first_tooth[i] = -1;
                if (kept[i]) continue;
                S = add128(S, mk128(qlo[i], qhi[i]));
                while (lt128(Q, S)) {
                    if (!copies[i]) first_tooth[i] = (long long)t;
                    copies[i]++; total++; t++;
                    tooth_step(Q, R, Tq, Tr, n32);
                }
            }
            S_thread = S;
        }
        run = S_thread;
    }
    return total;
}

}
