// sk_plan.h -- host-side plan data shared by the CUDA translation unit and the host planner.
#pragma once
#include <cstdint>

#define SK_WMAX 16        // widest exp-of-semicircle kernel (eps = 1e-15 -> w = 16)
#define SK_NC 16          // monomial coefficients per tap polynomial (8 even + 8 odd)
#define SK_NQMAX 24       // Chebyshev terms of the deconvolution factor

// Exp-of-semicircle kernel phi(z) = exp(beta (sqrt(1-z^2) - 1)), |z| <= 1, width w grid cells.
// Tap i (0 <= i < w) at offset x in [-1/2, 1/2) from the window centre is phi((x + i - w/2 + 1/2) 2/w).
// With s = 2x:  tap_i(s) = E_i(s^2) + s O_i(s^2),  tap_{w-1-i}(s) = E_i(s^2) - s O_i(s^2)   (phi is even).
struct SkEsPlan {
  int32_t w;                       // even, 4..16
  int32_t nq;                      // Chebyshev terms used
  double beta;
  double ximax;                    // deconvolution factor valid for |xi| <= ximax = pi w / 4 (sigma = 2), with margin
  double E[SK_WMAX / 2][SK_NC / 2];  // even-part coefficients, ascending powers of s^2
  double O[SK_WMAX / 2][SK_NC / 2];  // odd-part coefficients
  double qc[SK_NQMAX];             // (2/w)/phihat(xi) = sum_j qc[j] T_j(2 (xi/ximax)^2 - 1)
};

// Host planner (sk_plan_host.cpp, compiled by g++ with libquadmath)
int sk_plan_make_es(int w, SkEsPlan *out);
int sk_plan_gauss_rule(int n, double p, double *no, double *wt);
int sk_plan_jacobi_coeffs(int n, double p, double *A, double *B, double *C);   // (hi, lo) pairs, 2n doubles each
// piecewise-polynomial table of J_0 .. J_numax on [0, 2 nint] (sk_hankel.h): tab[numax+1][nint][nc]
int sk_plan_bessel_table(int numax, int nint, int nc, double *tab);
