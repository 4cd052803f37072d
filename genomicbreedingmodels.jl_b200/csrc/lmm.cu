// Per-marker REML delta search on rotated data -- the "per-marker solve on rotated data ...
// batched per-SNP delta searches rather than a P3D shortcut" of BASELINE.json's north_star;
// the reference model is gwasreml / loglikreml (/root/reference/src/gwas.jl:450-483, :549-613:
// V = s2u*GRM + s2e*I re-estimated per marker, then b[end]/sqrt(inv(X'V^-1X)[end])), restated
// as the standard REML on the symmetric GRM (oracle/lmm_oracle.py explains the differences).
//
// After K = U S U', y~ = U'y, C~ = U'[1,C], x~ = U'x (gemm_tn.cu), for lam = log(delta):
//   w_i = 1/(s_i + e^lam);  G_k = Z' W^k Z for Z = [C~, x~, y~], k = 1,2,3
//   f(lam)  = dLL/dlam  = -1/2 [ (n-q) R'/R + L' + D' ]
//   f'(lam) = d2LL/dlam2 (closed form from G1, G2, G3)
// One warp per marker: each evaluation is one pass over the marker's rotated column (and
// the shared s, y~, C~ vectors, which stay L1/L2 resident) accumulating the 3*NP + 2 sums
// in FP64 registers, a butterfly all-reduce, then the q x q algebra redundantly per lane.
// The root of f is bracketed by marching from the null-model estimate lam0 in steps of 0.5
// and polished by safeguarded Newton (bisection fallback) to |dlam| < 1e-13.
//
// objective = 1 ("reference objective"): the reference's own function and box instead of the standard REML --
// minimise 0.5 log det V + y'Py + log det(X'V^-1X) (gwas.jl:478) over theta = [s2e, s2u] in [eps, 1]^2 (:588),
// statistic b[end] / sqrt(inv(X'V^-1X)[end]) without a sigma^2 factor (:596-599).  On rotated data, with
// c = s2u, delta = s2e / s2u: the objective is (n/2 - q) log c + 1/2 sum log(s_i + delta) + R(delta) / c +
// log det A(delta); for a fixed delta it is unimodal in c with minimiser c* = R / (n/2 - q), and the box is
// c <= min(1, 1/delta), so the constrained minimiser is c* clamped and the search stays ONE-dimensional in
// lam = log(delta): F'(lam) below has three regimes (interior / s2u = 1 / s2e = 1), continuous except at lam = 0.
// oracle/lmm_oracle.py: refobj_terms / refobj_fit restate it; its dense 2-D brute force agrees to 1e-12.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "pvalue.cuh"

namespace gbm {

constexpr double kLamMinDefault = -11.512925464970229;  // log(1e-5)
constexpr double kLamMaxDefault = 11.512925464970229;   // log(1e+5)
constexpr double kMarch = 0.5;

// q x q symmetric positive-definite inverse by Gauss-Jordan (q <= 4, registers only)
template <int Q>
__host__ __device__ inline void spd_inverse(const double (&A)[Q][Q], double (&Ai)[Q][Q]) {
  double M[Q][2 * Q];
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      M[i][j] = A[i][j];
      M[i][Q + j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int c = 0; c < Q; ++c) {
    const double inv = 1.0 / M[c][c];
#pragma unroll
    for (int j = 0; j < 2 * Q; ++j) M[c][j] *= inv;
#pragma unroll
    for (int r = 0; r < Q; ++r) {
      if (r == c) continue;
      const double f = M[r][c];
#pragma unroll
      for (int j = 0; j < 2 * Q; ++j) M[r][j] -= f * M[c][j];
    }
  }
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) Ai[i][j] = M[i][Q + j];
}

// Everything derived from the three weighted Gram matrices at one lam.
// MM = number of Z columns (fixed effects + y), Q = MM - 1 fixed effects.
template <int MM>
struct RemlEval {
  double f, fp;        // dLL/dlam, d2LL/dlam2
  double beta_last;    // GLS coefficient of the last fixed effect
  double var_last;     // [ (X'WX)^-1 ]_last,last
  double R;            // residual quadratic form
  double c;            // reference objective: the fitted s2u (clamped c*); 0 for REML
};

template <int MM>
__host__ __device__ inline RemlEval<MM> reml_eval(const double* g1, const double* g2, const double* g3, double sw,
                                                  double sw2, double delta, double n, int objective = 0) {
  constexpr int Q = MM - 1;
  // unpack packed upper triangles (index of (a,b), a <= b: a*MM - a(a-1)/2 + (b-a))
  double G1[MM][MM], H1[MM][MM], H2[MM][MM];
  int idx = 0;
#pragma unroll
  for (int a = 0; a < MM; ++a)
#pragma unroll
    for (int b = a; b < MM; ++b) {
      const double x1 = g1[idx], x2 = g2[idx], x3 = g3[idx];
      ++idx;
      G1[a][b] = G1[b][a] = x1;
      H1[a][b] = H1[b][a] = -delta * x2;
      H2[a][b] = H2[b][a] = -delta * x2 + 2.0 * delta * delta * x3;
    }
  double A[Q][Q], Ai[Q][Q];
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) A[i][j] = G1[i][j];
  spd_inverse<Q>(A, Ai);
  double v[MM];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < Q; ++j) s += Ai[i][j] * G1[j][Q];
    v[i] = -s;
  }
  v[Q] = 1.0;
  auto quad = [&](const double (&Mx)[MM][MM]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MM; ++i)
#pragma unroll
      for (int j = 0; j < MM; ++j) s += v[i] * Mx[i][j] * v[j];
    return s;
  };
  const double R = quad(G1), R1 = quad(H1);
  double u[Q];
#pragma unroll
  for (int i = 0; i < Q; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < MM; ++j) s += H1[i][j] * v[j];
    u[i] = s;
  }
  double uAu = 0.0;
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) uAu += u[i] * Ai[i][j] * u[j];
  const double R2 = quad(H2) - 2.0 * uAu;
  // D' = tr(Ai H1xx), D'' = tr(Ai H2xx) - tr((Ai H1xx)^2)
  double M1[Q][Q];
  double D1 = 0.0, D2 = 0.0;
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        s1 += Ai[i][k] * H1[k][j];
        s2 += Ai[i][k] * H2[k][j];
      }
      M1[i][j] = s1;
      if (i == j) {
        D1 += s1;
        D2 += s2;
      }
    }
#pragma unroll
  for (int i = 0; i < Q; ++i)
#pragma unroll
    for (int j = 0; j < Q; ++j) D2 -= M1[i][j] * M1[j][i];
  const double L1 = delta * sw, L2 = delta * sw - delta * delta * sw2;
  const double rr = R1 / R;
  RemlEval<MM> e;
  if (objective == 0) {
    e.f = -0.5 * ((n - Q) * rr + L1 + D1);
    e.fp = -0.5 * ((n - Q) * (R2 / R - rr * rr) + L2 + D2);
    e.c = 0.0;
  } else {
    const double nq = 0.5 * n - Q, cstar = R / nq, chi = delta <= 1.0 ? 1.0 : 1.0 / delta;
    double F1, F2;
    if (cstar < chi) {          // interior: both variance components inside the box
      F1 = nq * rr + 0.5 * L1 + D1;
      F2 = nq * (R2 / R - rr * rr) + 0.5 * L2 + D2;
      e.c = cstar;
    } else if (delta <= 1.0) {  // s2u = 1
      F1 = 0.5 * L1 + R1 + D1;
      F2 = 0.5 * L2 + R2 + D2;
      e.c = 1.0;
    } else {                    // s2e = 1, s2u = 1 / delta
      F1 = -nq + 0.5 * L1 + delta * (R1 + R) + D1;
      F2 = 0.5 * L2 + delta * (R2 + 2.0 * R1 + R) + D2;
      e.c = chi;
    }
    e.f = -F1;  // the search below looks for a maximum of -F
    e.fp = -F2;
  }
  e.beta_last = -v[Q - 1];
  e.var_last = Ai[Q - 1][Q - 1];
  e.R = R;
  return e;
}

struct LmmParams {
  int64_t n, pb, ld;      // rotated block: n x pb, leading dimension ld
  const double* Ar;       // rotated marker columns
  const double* S;        // eigenvalues (n)
  const double* Yr;       // rotated phenotype (n)
  const double* Cr;       // rotated fixed covariates incl. intercept, n x Q0, leading dimension ldcr
  int64_t ldcr;
  double lam0;
  double lam_min, lam_max;  // search interval in lam = log(delta)
  int objective;            // 0: standard REML, 1: the reference's objective and box (see the header comment)
  const double* col_sd;   // raw-column sd (beta / se are reported for the standardised column)
  const uint8_t* keep;    // fixed-locus filter
  double* beta;
  double* se;
  double* stat;
  double* nlp;
  double* log_delta;
  int flags;
};

template <int Q0>
struct WarpGram {
  static constexpr int MM = Q0 + 2;
  static constexpr int NP = MM * (MM + 1) / 2;
};

// One pass over the marker's rotated column at lam: the 3*NP + 2 weighted sums, all-reduced.
template <int Q0>
__device__ __forceinline__ RemlEval<Q0 + 2> warp_eval(const LmmParams& prm, const double* __restrict__ x, double lam,
                                                      int lane) {
  constexpr int MM = Q0 + 2, NP = MM * (MM + 1) / 2;
  const double delta = exp(lam);
  double g1[NP], g2[NP], g3[NP], sw = 0.0, sw2 = 0.0;
#pragma unroll
  for (int i = 0; i < NP; ++i) g1[i] = g2[i] = g3[i] = 0.0;
  for (int64_t i = lane; i < prm.n; i += 32) {
    double z[MM];
#pragma unroll
    for (int c = 0; c < Q0; ++c) z[c] = __ldg(prm.Cr + c * prm.ldcr + i);
    z[Q0] = x[i];
    z[Q0 + 1] = __ldg(prm.Yr + i);
    const double w1 = 1.0 / (__ldg(prm.S + i) + delta);
    const double w2 = w1 * w1, w3 = w2 * w1;
    sw += w1;
    sw2 += w2;
    int idx = 0;
#pragma unroll
    for (int a = 0; a < MM; ++a)
#pragma unroll
      for (int b = a; b < MM; ++b) {
        const double pr = z[a] * z[b];
        g1[idx] = fma(w1, pr, g1[idx]);
        g2[idx] = fma(w2, pr, g2[idx]);
        g3[idx] = fma(w3, pr, g3[idx]);
        ++idx;
      }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      g1[i] += __shfl_xor_sync(0xffffffffu, g1[i], o);
      g2[i] += __shfl_xor_sync(0xffffffffu, g2[i], o);
      g3[i] += __shfl_xor_sync(0xffffffffu, g3[i], o);
    }
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    sw2 += __shfl_xor_sync(0xffffffffu, sw2, o);
  }
  return reml_eval<MM>(g1, g2, g3, sw, sw2, delta, static_cast<double>(prm.n), prm.objective);
}

template <int Q0>
__global__ void __launch_bounds__(256, Q0 == 1 ? 3 : 1) lmm_delta_kernel(const LmmParams prm) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  constexpr int Q = Q0 + 1;
  for (int64_t j = warp_global; j < prm.pb; j += nwarps) {
    if (prm.keep && !prm.keep[j]) {
      if (lane == 0) {
        if (prm.beta) prm.beta[j] = NAN;
        if (prm.se) prm.se[j] = NAN;
        if (prm.stat) prm.stat[j] = NAN;
        if (prm.nlp) prm.nlp[j] = NAN;
        if (prm.log_delta) prm.log_delta[j] = NAN;
      }
      continue;
    }
    const double* x = prm.Ar + j * prm.ld;
    // bracket the stationary point by marching from lam0 in the ascent direction
    const double kLamMin = prm.lam_min, kLamMax = prm.lam_max;
    double a = fmin(fmax(prm.lam0, kLamMin), kLamMax);
    auto ea = warp_eval<Q0>(prm, x, a, lane);
    double lam = a;
    bool done = (ea.f == 0.0);
    double lo = a, hi = a, flo = ea.f, fhi = ea.f;
    auto ex = ea;  // evaluation at the current best point
    if (!done) {
      const double dir = ea.f > 0.0 ? 1.0 : -1.0;
      double fa = ea.f;
      for (;;) {
        const double b = fmin(fmax(a + dir * kMarch, kLamMin), kLamMax);
        auto eb = warp_eval<Q0>(prm, x, b, lane);
        if (fa * eb.f <= 0.0) {
          if (a < b) { lo = a; flo = fa; hi = b; fhi = eb.f; }
          else       { lo = b; flo = eb.f; hi = a; fhi = fa; }
          break;
        }
        if (b == kLamMin || b == kLamMax) {  // monotone up to the bound: boundary estimate
          lam = b;
          ex = eb;
          done = true;
          break;
        }
        a = b;
        fa = eb.f;
      }
    }
    if (!done) {
      // safeguarded Newton on f inside [lo, hi] (f(lo) > 0 > f(hi) for a maximum of LL)
      double xk = lo - flo * (hi - lo) / (fhi - flo);  // secant start
      if (!(xk > lo && xk < hi)) xk = 0.5 * (lo + hi);
      for (int it = 0; it < 60; ++it) {
        ex = warp_eval<Q0>(prm, x, xk, lane);
        lam = xk;
        if (ex.f == 0.0) break;
        if ((ex.f > 0.0) == (flo > 0.0)) { lo = xk; flo = ex.f; }
        else                             { hi = xk; fhi = ex.f; }
        double xn = xk - ex.f / ex.fp;
        if (!(xn > lo && xn < hi) || !isfinite(xn)) xn = 0.5 * (lo + hi);
        if (fabs(xn - xk) < 1e-13 * fmax(1.0, fabs(xk)) || (hi - lo) < 1e-14) {
          // final evaluation at the converged point so beta / se belong to it
          ex = warp_eval<Q0>(prm, x, xn, lane);
          lam = xn;
          break;
        }
        xk = xn;
      }
    }
    if (lane == 0) {
      const double n = static_cast<double>(prm.n);
      // REML: sigma2_g profiled out, R / (n - q); reference objective: the fitted s2u itself (no sigma^2 factor in
      // b[end] / sqrt(inv(X'V^-1X)[end]), gwas.jl:596-599)
      const double sg2 = prm.objective == 0 ? ex.R / (n - Q) : ex.c;
      const double se_raw = sqrt(sg2 * ex.var_last);
      const double z = ex.beta_last / se_raw;
      const double sd = prm.col_sd ? prm.col_sd[j] : 1.0;
      if (prm.beta) prm.beta[j] = ex.beta_last * sd;
      if (prm.se) prm.se[j] = se_raw * sd;
      if (prm.stat) prm.stat[j] = z;
      if (prm.nlp) {
        double v = -log_sf_normal(z) * 0.4342944819032518;  // Normal(), gwas.jl:606
        if (prm.flags & 1) v -= 0.3010299956639812;
        prm.nlp[j] = v;
      }
      if (prm.log_delta) prm.log_delta[j] = lam;
    }
  }
}

void launch_lmm_delta(int Q0, const double* Ar, int64_t n, int64_t pb, int64_t ld, const double* S, const double* Yr,
                      const double* Cr, int64_t ldcr, double lam0, const double* col_sd, const uint8_t* keep,
                      double* beta, double* se, double* stat, double* nlp, double* log_delta, int flags,
                      int sm_count, cudaStream_t stream, int objective, double s_min) {
  if (pb <= 0) return;
  // s_i + delta must stay positive: an indefinite K (the symmetrised standardised GRM) raises the lower end
  double lam_min = kLamMinDefault;
  if (s_min < 0.0) lam_min = fmax(lam_min, log(-s_min * (1.0 + 1e-6) + 1e-300) + 1e-3);
  LmmParams prm{n, pb, ld, Ar, S, Yr, Cr, ldcr, lam0, lam_min, kLamMaxDefault, objective, col_sd, keep, beta, se, stat, nlp,
                log_delta, flags};
  const int64_t warps_needed = pb;
  int64_t blocks = (warps_needed + 7) / 8;
  const int64_t max_blocks = static_cast<int64_t>(sm_count) * 6;
  if (blocks > max_blocks) blocks = max_blocks;
  switch (Q0) {
    case 1: lmm_delta_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(prm); break;
    case 2: lmm_delta_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(prm); break;
    case 3: lmm_delta_kernel<3><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(prm); break;
    default: GBM_THROW(1, "lmm scan: at most 2 covariates besides the intercept are supported");
  }
  GBM_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------
// host side: null model lam0 (fixed effects C~ only) -- n-vector work, done once
// ------------------------------------------------------------------------------------
template <int Q0>
static double null_f(const double* S, const double* Cr, int64_t ldcr, const double* Yr, int64_t n, double lam,
                     double* ll_out) {
  constexpr int MM = Q0 + 1, NP = MM * (MM + 1) / 2;
  const double delta = exp(lam);
  double g1[NP] = {0}, g2[NP] = {0}, g3[NP] = {0}, sw = 0, sw2 = 0, slog = 0;
  for (int64_t i = 0; i < n; ++i) {
    double z[MM];
    for (int c = 0; c < Q0; ++c) z[c] = Cr[c * ldcr + i];
    z[Q0] = Yr[i];
    const double w1 = 1.0 / (S[i] + delta), w2 = w1 * w1, w3 = w2 * w1;
    sw += w1;
    sw2 += w2;
    slog += log(S[i] + delta);
    int idx = 0;
    for (int a = 0; a < MM; ++a)
      for (int b = a; b < MM; ++b) {
        const double pr = z[a] * z[b];
        g1[idx] += w1 * pr;
        g2[idx] += w2 * pr;
        g3[idx] += w3 * pr;
        ++idx;
      }
  }
  RemlEval<MM> e = reml_eval<MM>(g1, g2, g3, sw, sw2, delta, static_cast<double>(n));
  if (ll_out) {
    // log det A from the packed Gram (Q0 x Q0 leading block) via Gaussian elimination
    double A[Q0][Q0];
    int idx = 0;
    for (int a = 0; a < MM; ++a)
      for (int b = a; b < MM; ++b) {
        if (a < Q0 && b < Q0) A[a][b] = A[b][a] = g1[idx];
        ++idx;
      }
    double logdet = 0.0;
    for (int c = 0; c < Q0; ++c) {
      logdet += log(A[c][c]);
      for (int r = c + 1; r < Q0; ++r) {
        const double f = A[r][c] / A[c][c];
        for (int k = c; k < Q0; ++k) A[r][k] -= f * A[c][k];
      }
    }
    *ll_out = -0.5 * ((static_cast<double>(n) - Q0) * log(e.R) + slog + logdet);
  }
  return e.f;
}

template <int Q0>
static double null_lam0_t(const double* S, const double* Cr, int64_t ldcr, const double* Yr, int64_t n) {
  const int grid = 101;
  double kLamMin = kLamMinDefault;
  const double kLamMax = kLamMaxDefault;
  if (S[0] < 0.0) kLamMin = fmax(kLamMin, log(-S[0] * (1.0 + 1e-6) + 1e-300) + 1e-3);
  double best = -INFINITY;
  int gbest = 0;
  for (int g = 0; g < grid; ++g) {
    const double lam = kLamMin + (kLamMax - kLamMin) * g / (grid - 1);
    double ll;
    null_f<Q0>(S, Cr, ldcr, Yr, n, lam, &ll);
    if (ll > best) {
      best = ll;
      gbest = g;
    }
  }
  auto f = [&](double lam) { return null_f<Q0>(S, Cr, ldcr, Yr, n, lam, nullptr); };
  const double step = (kLamMax - kLamMin) / (grid - 1);
  double lo = kLamMin + step * (gbest > 0 ? gbest - 1 : 0);
  double hi = kLamMin + step * (gbest < grid - 1 ? gbest + 1 : grid - 1);
  double flo = f(lo), fhi = f(hi);
  if (!(flo > 0.0 && fhi < 0.0)) {
    // maximum at a boundary cell without an interior stationary point
    if (flo <= 0.0) return lo;
    if (fhi >= 0.0) return hi;
  }
  for (int it = 0; it < 200 && (hi - lo) > 1e-14; ++it) {  // bisection: robust, n-vector cost only
    const double mid = 0.5 * (lo + hi), fm = f(mid);
    if (fm == 0.0) return mid;
    if (fm > 0.0) lo = mid;
    else hi = mid;
  }
  return 0.5 * (lo + hi);
}

double lmm_null_lam0(int Q0, const double* S, const double* Cr, int64_t ldcr, const double* Yr, int64_t n) {
  switch (Q0) {
    case 1: return null_lam0_t<1>(S, Cr, ldcr, Yr, n);
    case 2: return null_lam0_t<2>(S, Cr, ldcr, Yr, n);
    case 3: return null_lam0_t<3>(S, Cr, ldcr, Yr, n);
    default: GBM_THROW(1, "lmm scan: at most 2 covariates besides the intercept are supported");
  }
}

}  // namespace gbm
