// Device-side Poisson upper tail and abundance p-value (product code, sm_100a).
//
// Replaces, for the B200 path, the reference's
//     calc_pA()  /root/reference/src/pval.cpp:44-64   (Rcpp::ppois -> R nmath, pval.cpp:50)
//     get_pA()   /root/reference/src/pval.cpp:67-89
// R's ppois(x, lambda, lower=FALSE) == pgamma(lambda, floor(x+1e-7)+1, 1, lower=TRUE); the
// branch structure of pgamma_raw is the one the reference author transcribed at
// pval.cpp:255-316 (upstream wch/r-source@af7f52f7): x<1 -> small-x series; x<=alph-1 ->
// upper series x dpois; alph-1<x -> lower series x dpois; otherwise the asymptotic normal
// expansion; results below DBL_MIN/DBL_EPSILON are redone in log space.
// All arithmetic is fp64 and this translation unit is compiled with -fmad=false so that
// a*b+c is not contracted (the CPU reference on x86-64 does not contract either).
// Contract: <= 1e-10 relative vs the reference (tests: <= 1e-12 vs exact summation).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <math_constants.h>

namespace dd2 {

#define DD_LN_SQRT_2PI 0.918938533204672741780329736406
#define DD_1_SQRT_2PI 0.398942280401432677939946059934
#define DD_2PI 6.283185307179586476925286766559
#define DD_SQRT_32 5.656854249492380195206754896838
#define DD_LN2 0.693147180559945309417232121458
#define DD_SCALEF 1.157920892373162e+77 /* 2^256 */

// Stirling-series error delta(n) = lgamma(n+1) - (n+1/2)log n + n - log sqrt(2pi) at n = k/2
__constant__ double c_sferr_halves[31] = {
    0.0,
    0.1534264097200273452913839393, 0.08106146679532725821967026359, 0.05481412105191765389613870235,
    0.04134069595540929409382208141, 0.03316287351993628748511050974, 0.02767792568499833914878929275,
    0.02374616365629749597133027909, 0.02079067210376509311152277177, 0.01848845053267318523077935748,
    0.01664469118982119216319486537, 0.01513497322191737887351383688, 0.01387612882307074799874572702,
    0.01281046524292022692425065528, 0.01189670994589177009505572412, 0.0111045597582069173266307552,
    0.01041126526197209649747856713, 0.009799416126158803298390373402, 0.009255462182712732917728636633,
    0.008768700134139385462955047269, 0.00833056343336287125646931866, 0.007934114564314020547249562491,
    0.007573675487951840794972024212, 0.007244554301320383179546196602, 0.006942840107209529865664152663,
    0.006665247032707682442356180895, 0.006408994188004207068439631083, 0.006171712263039457647534604798,
    0.005951370112758847735624416046, 0.005746216513010115682026102477, 0.00555473355196280137103868996};

__device__ __forceinline__ double d_ninf() { return -CUDART_INF; }

__device__ inline double log1_exp(double x) {  // log(1 - e^x), x < 0
  return (x > -DD_LN2) ? log(-expm1(x)) : log1p(-exp(x));
}

__device__ inline double logcf(double x, double i, double d, double eps) {
  double c1 = 2 * d, c2 = i + d, c4 = c2 + d, a1 = c2;
  double b1 = i * (c2 - i * x), b2 = d * d * x, a2 = c4 * c2 - b2;
  b2 = c4 * b1 - i * b2;
  while (fabs(a2 * b1 - a1 * b2) > fabs(eps * b1 * b2)) {
    double c3 = c2 * c2 * x;
    c2 += d; c4 += d;
    a1 = c4 * a2 - c3 * a1;
    b1 = c4 * b2 - c3 * b1;
    c3 = c1 * c1 * x;
    c1 += d; c4 += d;
    a2 = c4 * a1 - c3 * a2;
    b2 = c4 * b1 - c3 * b2;
    if (fabs(b2) > DD_SCALEF) {
      a1 /= DD_SCALEF; b1 /= DD_SCALEF; a2 /= DD_SCALEF; b2 /= DD_SCALEF;
    } else if (fabs(b2) < 1 / DD_SCALEF) {
      a1 *= DD_SCALEF; b1 *= DD_SCALEF; a2 *= DD_SCALEF; b2 *= DD_SCALEF;
    }
  }
  return a2 / b2;
}

__device__ inline double log1pmx(double x) {
  if (x > 1 || x < -0.79149064) return log1p(x) - x;
  double r = x / (2 + x), y = r * r;
  if (fabs(x) < 1e-2) {
    const double two = 2;
    return r * ((((two / 9 * y + two / 7) * y + two / 5) * y + two / 3) * y - x);
  }
  return r * (2 * y * logcf(y, 3, 2, 1e-14) - x);
}

__device__ inline double stirlerr(double n) {
  const double S0 = 0.083333333333333333333, S1 = 0.00277777777777777777778,
               S2 = 0.00079365079365079365079365, S3 = 0.000595238095238095238095238,
               S4 = 0.0008417508417508417508417508;
  double nn;
  if (n <= 15.0) {
    nn = n + n;
    if (nn == (int)nn) return c_sferr_halves[(int)nn];
    return lgamma(n + 1.) - (n + 0.5) * log(n) + n - DD_LN_SQRT_2PI;
  }
  nn = n * n;
  if (n > 500) return (S0 - S1 / nn) / n;
  if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
  if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
  return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

__device__ inline double bd0(double x, double np) {
  if (fabs(x - np) < 0.1 * (x + np)) {
    double v = (x - np) / (x + np);
    double s = (x - np) * v;
    if (fabs(s) < DBL_MIN) return s;
    double ej = 2 * x * v;
    v = v * v;
    for (int j = 1; j < 1000; j++) {
      ej *= v;
      double s1 = s + ej / ((j << 1) + 1);
      if (s1 == s) return s1;
      s = s1;
    }
  }
  return x * log(x / np) + np - x;
}

__device__ inline double dpois_raw(double x, double lambda, int give_log) {
  if (lambda == 0) return (x == 0) ? (give_log ? 0. : 1.) : (give_log ? d_ninf() : 0.);
  if (!isfinite(lambda)) return give_log ? d_ninf() : 0.;
  if (x < 0) return give_log ? d_ninf() : 0.;
  if (x <= lambda * DBL_MIN) return give_log ? -lambda : exp(-lambda);
  if (lambda < x * DBL_MIN) {
    if (!isfinite(x)) return give_log ? d_ninf() : 0.;
    double e = -lambda + x * log(lambda) - lgamma(x + 1);
    return give_log ? e : exp(e);
  }
  double f = DD_2PI * x, e = -stirlerr(x) - bd0(x, lambda);
  return give_log ? -0.5 * log(f) + e : exp(e) / sqrt(f);
}

// Cody (1969) rational approximations, as laid out in R's pnorm_both.
__device__ inline void pnorm_both(double x, double *cum, double *ccum, int i_tail, int log_p) {
  const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                       18154.981253343561249, 0.065682337918207449113};
  const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                       45507.789335026729956};
  const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                       597.27027639480026226,  2494.5375852903726711, 6848.1904505362823326,
                       11602.651437647350124,  9842.7148383839780218, 1.0765576773720192317e-8};
  const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                       6485.558298266760755,  18615.571640885098091, 34900.952721145977266,
                       38912.003286093271411, 19685.429676859990727};
  const double p[6] = {0.21589853405795699,     0.1274011611602473639, 0.022235277870649807,
                       0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
  const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                       0.00378239633202758244, 7.29751555083966205e-5};
  double xden, xnum, temp, del, xsq, y;
  const double eps = DBL_EPSILON * 0.5;
  const int lower = i_tail != 1, upper = i_tail != 0;
  if (isnan(x)) { *cum = *ccum = x; return; }
  y = fabs(x);
  if (y <= 0.67448975) {
    if (y > eps) {
      xsq = x * x;
      xnum = a[4] * xsq;
      xden = xsq;
#pragma unroll
      for (int i = 0; i < 3; ++i) { xnum = (xnum + a[i]) * xsq; xden = (xden + b[i]) * xsq; }
    } else xnum = xden = 0.0;
    temp = x * (xnum + a[3]) / (xden + b[3]);
    if (lower) *cum = 0.5 + temp;
    if (upper) *ccum = 0.5 - temp;
    if (log_p) { if (lower) *cum = log(*cum); if (upper) *ccum = log(*ccum); }
    return;
  }
  bool mid = y <= DD_SQRT_32;
  if (mid) {
    xnum = c[8] * y;
    xden = y;
#pragma unroll
    for (int i = 0; i < 7; ++i) { xnum = (xnum + c[i]) * y; xden = (xden + d[i]) * y; }
    temp = (xnum + c[7]) / (xden + d[7]);
  } else if ((log_p && y < 1e170) || (lower && -37.5193 < x && x < 8.2924) ||
             (upper && -8.2924 < x && x < 37.5193)) {
    xsq = 1.0 / (x * x);
    xnum = p[5] * xsq;
    xden = xsq;
#pragma unroll
    for (int i = 0; i < 4; ++i) { xnum = (xnum + p[i]) * xsq; xden = (xden + q[i]) * xsq; }
    temp = xsq * (xnum + p[4]) / (xden + q[4]);
    temp = (DD_1_SQRT_2PI - temp) / y;
  } else {
    if (x > 0) { *cum = log_p ? 0. : 1.; *ccum = log_p ? d_ninf() : 0.; }
    else { *cum = log_p ? d_ninf() : 0.; *ccum = log_p ? 0. : 1.; }
    return;
  }
  double X = mid ? y : x;
  xsq = trunc(X * 16) / 16;
  del = (X - xsq) * (X + xsq);
  if (log_p) {
    *cum = (-xsq * xsq * 0.5) + (-del * 0.5) + log(temp);
    if ((lower && x > 0.) || (upper && x <= 0.))
      *ccum = log1p(-exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp);
  } else {
    *cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
    *ccum = 1.0 - *cum;
  }
  if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
}

__device__ inline double pnorm_std(double x, int lower_tail, int log_p) {
  double p = 0, cp = 0;
  if (!isfinite(x)) {
    if (isnan(x)) return x;
    if (x < 0) return lower_tail ? (log_p ? d_ninf() : 0.) : (log_p ? 0. : 1.);
    return lower_tail ? (log_p ? 0. : 1.) : (log_p ? d_ninf() : 0.);
  }
  pnorm_both(x, &p, &cp, lower_tail ? 0 : 1, log_p);
  return lower_tail ? p : cp;
}

__device__ inline double dnorm_std(double x) {
  x = fabs(x);
  if (x >= 2 * sqrt(DBL_MAX)) return 0.;
  if (x < 5) return DD_1_SQRT_2PI * exp(-0.5 * x * x);
  if (x > sqrt(-2 * DD_LN2 * (DBL_MIN_EXP + 1 - DBL_MANT_DIG))) return 0.;
  double x1 = ldexp(nearbyint(ldexp(x, 16)), -16);
  double x2 = x - x1;
  return DD_1_SQRT_2PI * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
}

__device__ inline double dpois_wrap(double x_plus_1, double lambda, int give_log) {
  const double M_cutoff = DD_LN2 * DBL_MAX_EXP / DBL_EPSILON;
  if (!isfinite(lambda)) return give_log ? d_ninf() : 0.;
  if (x_plus_1 > 1) return dpois_raw(x_plus_1 - 1, lambda, give_log);
  if (lambda > fabs(x_plus_1 - 1) * M_cutoff) {
    double e = -lambda - lgamma(x_plus_1);
    return give_log ? e : exp(e);
  }
  double d = dpois_raw(x_plus_1, lambda, give_log);
  return give_log ? d + log(x_plus_1 / lambda) : d * (x_plus_1 / lambda);
}

__device__ inline double pgamma_smallx(double x, double alph, int lower_tail, int log_p) {
  double sum = 0, c = alph, n = 0, term;
  do {
    n++;
    c *= -x / n;
    term = c / (alph + n);
    sum += term;
  } while (fabs(term) > DBL_EPSILON * fabs(sum));
  if (lower_tail) {
    double f1 = log_p ? log1p(sum) : 1 + sum;
    double f2;
    if (alph > 1) {
      f2 = dpois_raw(alph, x, log_p);
      f2 = log_p ? f2 + x : f2 * exp(x);
    } else if (log_p)
      f2 = alph * log(x) - lgamma(alph + 1);
    else
      f2 = pow(x, alph) / exp(lgamma(alph + 1));
    return log_p ? f1 + f2 : f1 * f2;
  }
  double lf2 = alph * log(x) - lgamma(alph + 1);
  if (log_p) return log1_exp(log1p(sum) + lf2);
  double f1m1 = sum, f2m1 = expm1(lf2);
  return -(f1m1 + f2m1 + f1m1 * f2m1);
}

__device__ inline double pd_upper(double x, double y, int log_p) {
  double term = x / y, sum = term;
  do {
    y++;
    term *= x / y;
    sum += term;
  } while (term > sum * DBL_EPSILON);
  return log_p ? log(sum) : sum;
}

__device__ inline double pd_lower_cf(double y, double d) {
  double f = 0.0, of, f0, i, c2, c3, c4, a1, b1, a2, b2;
  if (y == 0) return 0;
  f0 = y / d;
  if (fabs(y - 1) < fabs(d) * DBL_EPSILON) return f0;
  if (f0 > 1.) f0 = 1.;
  c2 = y; c4 = d;
  a1 = 0; b1 = 1; a2 = y; b2 = d;
  while (b2 > DD_SCALEF) { a1 /= DD_SCALEF; b1 /= DD_SCALEF; a2 /= DD_SCALEF; b2 /= DD_SCALEF; }
  i = 0; of = -1.;
  while (i < 200000) {
    i++; c2--; c3 = i * c2; c4 += 2;
    a1 = c4 * a2 + c3 * a1;
    b1 = c4 * b2 + c3 * b1;
    i++; c2--; c3 = i * c2; c4 += 2;
    a2 = c4 * a1 + c3 * a2;
    b2 = c4 * b1 + c3 * b2;
    if (b2 > DD_SCALEF) { a1 /= DD_SCALEF; b1 /= DD_SCALEF; a2 /= DD_SCALEF; b2 /= DD_SCALEF; }
    if (b2 != 0) {
      f = a2 / b2;
      if (fabs(f - of) <= DBL_EPSILON * fmax(f0, fabs(f))) return f;
      of = f;
    }
  }
  return f;
}

__device__ inline double pd_lower_series(double lambda, double y) {
  double term = 1, sum = 0;
  while (y >= 1 && term > sum * DBL_EPSILON) {
    term *= y / lambda;
    sum += term;
    y--;
  }
  if (y != floor(y)) sum += term * pd_lower_cf(y, lambda + 1 - y);
  return sum;
}

__device__ inline double dpnorm(double x, int lower_tail, double lp) {
  if (x < 0) { x = -x; lower_tail = !lower_tail; }
  if (x > 10 && !lower_tail) {
    double term = 1 / x, sum = term, x2 = x * x, i = 1;
    do {
      term *= -i / x2;
      sum += term;
      i += 2;
    } while (fabs(term) > DBL_EPSILON * sum);
    return 1 / sum;
  }
  return dnorm_std(x) / exp(lp);
}

__device__ inline double ppois_asymp(double x, double lambda, int lower_tail, int log_p) {
  const double ca[8] = {-1e99, 2 / 3., -4 / 135., 8 / 2835., 16 / 8505., -8992 / 12629925.,
                        -334144 / 492567075., 698752 / 1477701225.};
  const double cb[8] = {-1e99, 1 / 12., 1 / 288., -139 / 51840., -571 / 2488320., 163879 / 209018880.,
                        5246819 / 75246796800., -534703531 / 902961561600.};
  double dfm = lambda - x;
  double pt_ = -log1pmx(dfm / x);
  double s2pt = sqrt(2 * x * pt_);
  if (dfm < 0) s2pt = -s2pt;
  double res12 = 0, res1_term, res1_ig, res2_term, res2_ig;
  res1_ig = res1_term = sqrt(x);
  res2_ig = res2_term = s2pt;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    res12 += res1_ig * ca[i];
    res12 += res2_ig * cb[i];
    res1_term *= pt_ / i;
    res2_term *= 2 * pt_ / (2 * i + 1);
    res1_ig = res1_ig / x + res1_term;
    res2_ig = res2_ig / x + res2_term;
  }
  double elfb = x, elfb_term = 1;
#pragma unroll
  for (int i = 1; i < 8; i++) {
    elfb += elfb_term * cb[i];
    elfb_term /= x;
  }
  if (!lower_tail) elfb = -elfb;
  double f = res12 / elfb;
  double np = pnorm_std(s2pt, !lower_tail, log_p);
  if (log_p) return np + log1p(f * dpnorm(s2pt, !lower_tail, np));
  return np + f * dnorm_std(s2pt);
}

__device__ inline double pgamma_raw_1(double x, double alph, int lower_tail, int log_p) {
  double res;
  if (x <= 0) return lower_tail ? (log_p ? d_ninf() : 0.) : (log_p ? 0. : 1.);
  if (isinf(x)) return lower_tail ? (log_p ? 0. : 1.) : (log_p ? d_ninf() : 0.);
  if (x < 1) {
    res = pgamma_smallx(x, alph, lower_tail, log_p);
  } else if (x <= alph - 1 && x < 0.8 * (alph + 50)) {
    double sum = pd_upper(x, alph, log_p);
    double d = dpois_wrap(alph, x, log_p);
    if (!lower_tail) res = log_p ? log1_exp(d + sum) : 1 - d * sum;
    else res = log_p ? sum + d : sum * d;
  } else if (alph - 1 < x && alph < 0.8 * (x + 50)) {
    double sum, d = dpois_wrap(alph, x, log_p);
    if (alph < 1) {
      if (x * DBL_EPSILON > 1 - alph) sum = log_p ? 0. : 1.;
      else {
        double f = pd_lower_cf(alph, x - (alph - 1)) * x / alph;
        sum = log_p ? log(f) : f;
      }
    } else {
      sum = pd_lower_series(x, alph - 1);
      sum = log_p ? log1p(sum) : 1 + sum;
    }
    if (!lower_tail) res = log_p ? sum + d : sum * d;
    else res = log_p ? log1_exp(d + sum) : 1 - d * sum;
  } else {
    res = ppois_asymp(alph - 1, x, !lower_tail, log_p);
  }
  return res;
}

__device__ inline double pgamma_raw(double x, double alph, int lower_tail) {
  double res = pgamma_raw_1(x, alph, lower_tail, 0);
  // accuracy is lost to underflow close to DBL_MIN: redo in log space (pval.cpp:310-316)
  if (res < DBL_MIN / DBL_EPSILON) return exp(pgamma_raw_1(x, alph, lower_tail, 1));
  return res;
}

// P(X > x) for X ~ Poisson(lambda)   ==   Rcpp::ppois(x, lambda, lower=false)
__device__ inline double ppois_upper(double x, double lambda) {
  if (isnan(x) || isnan(lambda)) return x + lambda;
  if (lambda < 0.) return CUDART_NAN;
  if (x < 0) return 1.;
  if (lambda == 0.) return 0.;
  if (!isfinite(x)) return 0.;
  x = floor(x + 1e-7);
  return pgamma_raw(lambda, x + 1, 1);
}

// pval.cpp:44-64
__device__ inline double calc_pA(int reads, double E_reads, bool prior) {
  double pval = ppois_upper((double)(reads - 1), E_reads);
  if (!prior) {
    double norm = (1.0 - exp(-E_reads));
    if (norm < 1e-7) norm = E_reads - 0.5 * E_reads * E_reads;  // TAIL_APPROX_CUTOFF dada.h:25
    pval = pval / norm;
  }
  return pval;
}

}  // namespace dd2
