// CPU check of csrc/powtab.h + csrc/pqfast.h (test infrastructure): builds the tables exactly as the library does, then
//   1. every table: the stored bound really bounds the error against powl on random points,
//   2. every chain: |fast - exact| <= rel * |value| against the float64 chain with libm pow (the exact path's arithmetic),
//   3. the keep-or-recompute rule: whenever round_is_safe accepts, (float)fast == (float)exact; and how often it refuses.
// usage: powtab_check [samples per test, default 2000000]     prints one line per test and exits 1 on any violation
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../adaptive-edge-aware-jpeg_b200/csrc/pqfast.h"
#include "../../adaptive-edge-aware-jpeg_b200/csrc/pqtabs_build.h"

static uint64_t rng = 88172645463325252ull;
static double urand() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (double)(rng >> 11) / 9007199254740992.0; }
static double lograd(double lo, double hi) { return exp(log(lo) + (log(hi) - log(lo)) * urand()); }

static double exact_inv_eotf(double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0, m1 = 2610.0 / 16384.0;
    double t = pow(c / 10000.0, m1);
    return pow((c1 + c2 * t) / (1.0 + c3 * t), m2);
}
static double exact_eotf(double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0, m1 = 2610.0 / 16384.0;
    double t = pow(c, 1.0 / m2);
    double num = t - c1, den = c2 - c3 * t;
    if (num < 0.0) num = 0.0;
    if (den <= 0.0) den = 1e-12;
    return 10000.0 * pow(num / den, 1.0 / m1);
}

int main(int argc, char** argv) {
    const long N = argc > 1 ? atol(argv[1]) : 2000000;
    PqTabsHost H;
    pqtabs_build(H);
    const PqTabs& Q = H.view;
    int bad = 0;
    const struct { const char* name; const PowTabHost* t; double m; } tabs[] = {
        {"m1", &H.m1, 2610.0 / 16384.0}, {"m2[pq]", &H.m2[0], 2523.0 / 32.0}, {"m2[jz]", &H.m2[1], 1.7 * 2523.0 / 32.0},
        {"1/m2[pq]", &H.im2[0], 32.0 / 2523.0}, {"1/m2[jz]", &H.im2[1], 1.0 / (1.7 * 2523.0 / 32.0)}, {"1/m1", &H.im1, 16384.0 / 2610.0},
        {"1/2.4", &H.isrgb, 1.0 / 2.4}, {"f32(1/3)", &H.cbrt32, (double)(float)(1.0 / 3.0)}, {"cube", &H.cube, 3.0}};
    for (auto& tb : tabs) {
        const PowTabView& v = tb.t->v;
        double worst = 0.0;
        long inside = 0;
        for (long i = 0; i < N; i++) {
            const double lo = ldexp(1.0 + (double)v.jmin / (1 << v.lg_nseg), v.emin), hi = ldexp(1.0, v.emin + v.nexp) * (1 - 1e-16);
            const double x = (v.nexp == 1) ? lo + (hi - lo) * urand() : lograd(lo, hi);
            bool ok = true;
            const double got = (v.deg == 12) ? powtab_eval<12>(v, x, ok) : powtab_eval<8>(v, x, ok);
            if (!ok) continue;
            inside++;
            const long double want = powl((long double)x, (long double)tb.m);
            const double rel = (double)fabsl(((long double)got - want) / want);
            if (rel > worst) worst = rel;
        }
        const bool okb = worst <= v.eps && inside > N / 2;
        printf("table %-9s segs %5d deg %d  %6.1f KB  measured at build %.2e  random max %.2e  bound %.2e  %s\n", tb.name,
               (1 << v.lg_nseg) - v.jmin, v.deg, ((size_t)((1 << v.lg_nseg) - v.jmin) * v.stride + v.nexp) * 8 / 1024.0, tb.t->measured, worst, v.eps,
               okb ? "ok" : "VIOLATED");
        bad += !okb;
    }
    for (int w = 0; w < 2; w++)
        printf("curve enc[%d] %6.1f KB measured %.2e bound %.2e (+ exact-path noise -> %.2e);  dec[%d] %6.1f KB measured %.2e bound %.2e (-> %.2e)\n", w,
               H.enc[w].coef.size() * 8 / 1024.0, H.enc[w].measured, H.enc[w].v.eps, Q.enc_rel[w], w, H.dec[w].coef.size() * 8 / 1024.0, H.dec[w].measured,
               H.dec[w].v.eps, Q.dec_rel[w]);
    // forward chain, float inputs (ICtCp / ICaCb: l is a float32) and double inputs (JzAzBz)
    for (int which = 0; which < 2; which++) {
        const double m2 = which ? 1.7 * 2523.0 / 32.0 : 2523.0 / 32.0;
        long viol = 0, rejected = 0, unsafe = 0, flips = 0;
        double worst_ratio = 0.0, worst_rel = 0.0;
        for (long i = 0; i < N; i++) {
            double c = lograd(1e-6, 1.2);
            if (!which) c = (double)(float)c;
            if (i % 1000 == 0) c = 0.0;
            double rel; bool ok = true;
            const double f = pqf_inv_eotf(Q, which, m2, c, rel, ok);
            if (!ok) { rejected++; continue; }
            const double e = exact_inv_eotf(c, m2);
            const double d = fabs(f - e) / e;
            if (d > rel) viol++;
            if (rel > worst_rel) worst_rel = rel;
            if (d / rel > worst_ratio) worst_ratio = d / rel;
            if (round_is_safe(f, rel * f)) flips += ((float)f != (float)e); else unsafe++;
        }
        printf("pq_inv_eotf[%s]: rejected %ld (out of domain), bound violated %ld, worst |d|/bound %.3f, worst bound %.2e, unsafe roundings %ld (%.2e), accepted-but-different %ld\n",
               which ? "jz" : "pq", rejected, viol, worst_ratio, worst_rel, unsafe, (double)unsafe / N, flips);
        bad += (viol != 0) + (flips != 0) + (rejected > N / 100 + N / 500);
    }
    // inverse chain
    for (int which = 0; which < 2; which++) {
        const double m2 = which ? 1.7 * 2523.0 / 32.0 : 2523.0 / 32.0;
        const double kink = pow(3424.0 / 4096.0, m2);
        long viol = 0, rejected = 0, unsafe = 0, flips = 0, zeros = 0, unsafe_bulk = 0;
        double worst_ratio = 0.0;
        for (long i = 0; i < N; i++) {
            const bool near_kink = (i % 4 == 0);
            double y = near_kink ? kink * (1.0 + (urand() - 0.5) * 1e-3) : lograd(kink * 0.5, which ? 0.02 : 0.2);
            if (i % 7 == 0) y = (double)(float)y;
            double rel; bool ok = true;
            const double f = pqf_eotf(Q, which, y, rel, ok);
            if (!ok) { rejected++; continue; }
            const double e = exact_eotf(y, m2);
            if (f == 0.0) { zeros++; if (e != 0.0) viol++; continue; }
            const double d = fabs(f - e) / e;
            if (d > rel) viol++;
            if (d / rel > worst_ratio) worst_ratio = d / rel;
            if (round_is_safe(f, rel * f)) flips += ((float)f != (float)e); else { unsafe++; unsafe_bulk += !near_kink; }
        }
        printf("pq_eotf[%s]: rejected %ld, exact zeros %ld, bound violated %ld, worst |d|/bound %.3f, unsafe roundings %ld (%.2e; away from the kink %.2e), accepted-but-different %ld\n",
               which ? "jz" : "pq", rejected, zeros, viol, worst_ratio, unsafe, (double)unsafe / N, (double)unsafe_bulk / (0.75 * N), flips);
        bad += (viol != 0) + (flips != 0);
    }
    // sRGB encode
    {
        long viol = 0, unsafe = 0, flips = 0, rejected = 0;
        for (long i = 0; i < N; i++) {
            const double d = (double)(float)lograd(1e-4, 1.5);
            double E; bool ok = true;
            const double f = pqf_linear_to_srgb(Q, d, E, ok);
            if (!ok) { rejected++; continue; }
            const double e = d <= 0.0031308 ? d * 12.92 : 1.055 * pow(d, 1.0 / 2.4) - 0.055;
            if (fabs(f - e) > E) viol++;
            if (round_is_safe(f, E)) flips += ((float)f != (float)e); else unsafe++;
        }
        printf("linear_to_srgb: rejected %ld, bound violated %ld, unsafe roundings %ld (%.2e), accepted-but-different %ld\n", rejected, viol, unsafe, (double)unsafe / N, flips);
        bad += (viol != 0) + (flips != 0);
    }
    return bad ? 1 : 0;
}
