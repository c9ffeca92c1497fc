// Exact skip-ahead of castRay's tMax recurrence  t = fl(t + d)  -- a measured dead end kept here
// with its proof-by-test: advance_n must equal the literal loop  for (k < n) t = t + d  bit for bit.
//
// While t stays inside one binade [2^e, 2^(e+1)) every addition rounds on the same grid u = 2^(e-52): writing
// d = q*u + rho, the rounded increment is the constant D = q*u or (q+1)*u when rho != u/2; in the tie case rho == u/2
// round-half-even makes every result an even multiple of u, so the increment is constant from the SECOND in-binade
// addition on (the first may start from an odd multiple).  Hence after three real additions t1, t2, t3 inside one
// binade, D = t3 - t2 (exact) and t3 + m*D is exactly the value after m further additions as long as it stays below
// 2^(e+1); fma(m, D, t3) evaluates it without rounding because the result lies on the binade's grid.  m is a
// conservative estimate ((2^(e+1) - t3) / d, shaved by 1e-6 and 1), so a jump never reaches the binade's top; the
// additions that cross into the next binade are executed for real by the next round.
//
// On the B200 this bought 2 % on C2 (~130 additions per axis) and cost 7 % on C3 (~65 per axis), so the march kernel
// keeps its plain counted loops.  Build: g++ -O2 -ffp-contract=off.  Usage: test_advance [iterations]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static inline int hi_word(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
}
static inline double from_hi_word(int hi) {
    const uint64_t b = (uint64_t)(uint32_t)hi << 32;
    double x;
    memcpy(&x, &b, 8);
    return x;
}

static void advance_n(double& t, double d, int n) {
    if (n >= 24) {
        const double rcp = (double)(1.0f / (float)d) * 0.999999;  // ~1/d, 1e-7 accurate, biased low
        while (n >= 12) {
            const double t1 = t + d, t2 = t1 + d, t3 = t2 + d;
            n -= 3;
            t = t3;
            const int e1 = hi_word(t1) >> 20, e3 = hi_word(t3) >> 20;
            if (e1 == e3) {
                const double D = t3 - t2;                          // exact: same binade
                const double top = from_hi_word((e3 + 1) << 20);   // 2^(e+1)
                int m = (int)((top - t3) * rcp) - 1;               // additions that certainly stay below top
                m = m < n ? m : n;
                if (m > 0) {
                    t = fma((double)m, D, t3);                     // exact: the result is on the binade's grid
                    n -= m;
                }
            }
        }
    }
    for (int k = 0; k < n; k++) t = t + d;
}

int main(int argc, char** argv) {
    const long iters = argc > 1 ? atol(argv[1]) : 2000000;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0, 1);
    long bad = 0, tot = 0;
    for (long it = 0; it < iters; it++) {
        const double dir = (double)(float)(U(rng) * 2 - 1);
        if (dir == 0) continue;
        const double res = (it & 1) ? 0.001 : 0.002;
        double d = res / fabs(dir), t0;
        const int mode = (it % 3 == 0) ? 4 + (int)(it % 2) : (int)(it % 7);
        if (mode == 0) t0 = d * 0.5;
        else if (mode == 1) t0 = (res * 0.5 + 1e-9 * U(rng)) / fabs(dir);
        else if (mode == 2) t0 = d * U(rng);
        else if (mode == 3) { d = ldexp(1.0, -(int)(U(rng) * 12) - 1); t0 = d * 0.5; }
        else if (mode == 4) { const uint64_t m = (uint64_t)(U(rng) * 16) + 16; d = ldexp((double)m, -14); t0 = d * (0.5 + (int)(U(rng) * 4) * 0.25); }
        else if (mode == 5) { uint64_t b; memcpy(&b, &d, 8); const int z = (int)(U(rng) * 12); b = ((b >> (z + 1)) << (z + 1)) | (1ull << z); memcpy(&d, &b, 8); t0 = d * U(rng); }
        else t0 = (res * 0.5) / fabs(dir);
        const int n = (int)(U(rng) * (it % 3 == 0 ? 5000 : 400));
        double a = t0;
        for (int k = 0; k < n; k++) a = a + d;
        double b = t0;
        advance_n(b, d, n);
        tot++;
        if (memcmp(&a, &b, 8)) {
            if (bad++ < 10) printf("MISMATCH t0=%a d=%a n=%d loop=%a skip=%a\n", t0, d, n, a, b);
        }
    }
    printf("tests %ld mismatches %ld\n", tot, bad);
    return bad != 0;
}
