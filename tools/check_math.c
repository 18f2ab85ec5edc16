// Exhaustive accuracy check of sb_math.h against float64 libm.  gcc -O2 -ffp-contract=off -mfma -fopenmp
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <float.h>
#include "../sionna_b200/csrc/sb_math.h"
static double ulp_of(double ref) { float f = (float)ref; if (f == 0) return FLT_MIN; int e; frexpf(fabsf(f), &e); return ldexp(1.0, e - 24); }
// usage: check_math [stride]   stride 1 = every float (about 100 s on 8 cores); tests/test_sb_math.py runs a stride of
// 997 (prime: every exponent and a spread of mantissas) and asserts the bounds sb_math.h states.
int main(int argc, char** argv) {
    const long long stride = argc > 1 ? atoll(argv[1]) : 1;
    double worst_exp = 0, worst_log = 0, worst_tanh = 0, worst_atanh = 0; float wx = 0, wl = 0, wt = 0, wa = 0;
    #pragma omp parallel
    {
        double we = 0, wlg = 0, wth = 0, wat = 0; float xe = 0, xl = 0, xt = 0, xa = 0;
        #pragma omp for schedule(static)
        for (long long i = 0; i < (1LL << 32); i += stride) {
            uint32_t u = (uint32_t)i; float x; memcpy(&x, &u, 4);
            if (!(x == x) || isinf(x)) continue;
            if (x >= -87.3f && x <= 88.7f) {
                double ref = exp((double)x); double err = fabs((double)sb_expf(x) - ref) / ulp_of(ref);
                if (err > we) { we = err; xe = x; }
            }
            if (x >= FLT_MIN) {
                double ref = log((double)x); double err = fabs((double)sb_logf(x) - ref) / ulp_of(ref);
                if (err > wlg) { wlg = err; xl = x; }
            }
            if (fabsf(x) <= 20.f && fabsf(x) >= 1e-30f) {
                double ref = tanh((double)x); double err = fabs((double)sb_tanhf(x) - ref) / ulp_of(ref);
                if (err > wth) { wth = err; xt = x; }
            }
            if (fabsf(x) < 1.0f && fabsf(x) >= 1e-30f) {
                double ref = atanh((double)x); double err = fabs((double)sb_atanhf(x) - ref) / ulp_of(ref);
                if (err > wat) { wat = err; xa = x; }
            }
        }
        #pragma omp critical
        { if (we > worst_exp) { worst_exp = we; wx = xe; } if (wlg > worst_log) { worst_log = wlg; wl = xl; }
          if (wth > worst_tanh) { worst_tanh = wth; wt = xt; } if (wat > worst_atanh) { worst_atanh = wat; wa = xa; } }
    }
    printf("expf max ulp err %.4f at %a\nlogf max ulp err %.4f at %a\ntanhf max ulp err %.4f at %a\natanhf max ulp err %.4f at %a\n",
           worst_exp, wx, worst_log, wl, worst_tanh, wt, worst_atanh, wa);
    float x = 8.5e-8f; printf("exp(8.5e-8)=%a  phi(0)=%.9g phi(8.5e-8)=%.9g phi(1)=%.9g phi(10)=%.9g phi(16)=%.9g phi(16.635532)=%.9g phi(40)=%.9g\n",
        sb_expf(x), sb_phif(0.f), sb_phif(x), sb_phif(1.f), sb_phif(10.f), sb_phif(16.f), sb_phif(16.635532f), sb_phif(40.f));
    {   // table log used inside phi: accuracy on all positive normals, and phi >= 0 / monotone on its whole domain
        double wl2 = 0; float xl2 = 0; long long neg = 0, nonmono = 0; float xneg = 0;
        #pragma omp parallel
        {
            double w_ = 0; float x_ = 0; long long ng = 0, nm = 0; float xn = 0;
            #pragma omp for schedule(static)
            for (long long i = 0x00800000LL; i < 0x7f800000LL; i += stride) {
                uint32_t u = (uint32_t)i; float y; memcpy(&y, &u, 4);
                double ref = log((double)y);
                // y >= 1: error in ulps of the result; y < 1: absolute error in units of 2^-24 (what phi needs)
                double err = fabs((double)sb_logf_tab(y) - ref) / (y >= 1.0f ? ulp_of(ref) : fmax(ulp_of(ref), ldexp(1.0, -24)));
                if (err > w_) { w_ = err; x_ = y; }
                if (y >= 8.5e-8f && y <= 16.635532f) {
                    float ph = sb_phif(y);
                    if (ph < 0.f || (ph == 0.f && signbit(ph))) { ++ng; xn = y; }
                    uint32_t u2 = u + 1; float y2; memcpy(&y2, &u2, 4);
                    if (y2 <= 16.635532f && sb_phif(y2) > ph) ++nm;
                }
            }
            #pragma omp critical
            { if (w_ > wl2) { wl2 = w_; xl2 = x_; } neg += ng; nonmono += nm; if (ng) xneg = xn; }
        }
        printf("logf_tab max err %.4f (ulp for y>=1, max(ulp, 2^-24) for y<1) at %a; phi<0 count %lld (last at %a); phi non-monotone steps %lld\n", wl2, xl2, neg, xneg, nonmono);
    }
    printf("log(2^24)=%a log(2^24-1)=%a phi(2*phi(0))=%g\n", sb_logf(16777216.f), sb_logf(16777215.f), sb_phif(2*sb_phif(0.f)));
    return 0;
}
