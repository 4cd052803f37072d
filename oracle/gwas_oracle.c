/* CPU oracle, C twin  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates the per-marker arithmetic of GenomicBreedingModels.jl v0.3.0 for timing the
 * reference's algorithm on host cores (bench.py cpu_baseline / --impl reference) and for
 * cross-checking oracle/gwas_oracle.py.  PARITY UNPINNED where gwas_oracle.py says so
 * (no Julia runtime in this image; the reference holds no golden vectors for this path).
 *
 * Followed lines:
 *   std(G, dims=1), idx_cols                /root/reference/src/gwas.jl:112-113
 *   G = (G .- mean) ./ v                    /root/reference/src/gwas.jl:129
 *   X = hcat(ones, pc, G[:,j]); pinv(X'X)   /root/reference/src/gwas.jl:241-243
 *   b[end] / sqrt(Vinv[end,end])            /root/reference/src/gwas.jl:245
 *   Threads.@threads over markers           /root/reference/src/gwas.jl:239   (-> OpenMP)
 *   gwaslmm z, closed form                  /root/reference/src/gwas.jl:358-385
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_set_num_threads(int t) {
#ifdef _OPENMP
  if (t > 0) omp_set_num_threads(t);
#else
  (void)t;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Symmetric 3x3 pseudo-inverse through a cyclic Jacobi eigen-decomposition; eigenvalues
 * <= 3*eps*max|lambda| are dropped = Julia's pinv default rtol (eps*min(size)). */
static void pinv_sym3(const double S[3][3], double P[3][3]) {
  double a[3][3], v[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { a[i][j] = S[i][j]; v[i][j] = (i == j); }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq; a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk; a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq; v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  double lmax = fmax(fabs(a[0][0]), fmax(fabs(a[1][1]), fabs(a[2][2])));
  double tol = 3.0 * 2.220446049250313e-16 * lmax;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++)
        if (fabs(a[k][k]) > tol) s += v[i][k] * v[j][k] / a[k][k];
      P[i][j] = s;
    }
}

/* Two-pass column mean / corrected sd (Julia std(G, dims=1)). */
void oracle_colstats(const double* A, int64_t n, int64_t p, int64_t lda, double* mean, double* sd) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < p; j++) {
    const double* a = A + j * lda;
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i];
    double m = s / (double)n, ss = 0.0;
    for (int64_t i = 0; i < n; i++) { double d = a[i] - m; ss += d * d; }
    mean[j] = m;
    sd[j] = sqrt(ss / (double)(n - 1));
  }
}

/* The reference's marker loop on raw columns: standardise the column (gwas.jl:129), build
 * X'X and X'y for X = [1, pc, g] (gwas.jl:241), pinv, statistic (gwas.jl:242-245).
 * Columns failing the fixed-locus filter get stat = NaN and keep = 0.
 * stat_lmm (nullable) receives the gwaslmm z closed form from the same sums. */
void oracle_gwasols_raw(const double* A, int64_t n, int64_t p, int64_t lda, const double* ys,
                        const double* pc, double* stat_ols, double* stat_lmm, uint8_t* keep) {
  double s1p = 0, spp = 0, t1 = 0, tp = 0, tyy = 0;
  for (int64_t i = 0; i < n; i++) { s1p += pc[i]; spp += pc[i] * pc[i]; t1 += ys[i]; tp += pc[i] * ys[i]; tyy += ys[i] * ys[i]; }
  /* y'My for M = I - Q(Q'Q)^-1Q', Q = [1, pc] */
  double det = (double)n * spp - s1p * s1p;
  double yMy = tyy - (spp * t1 * t1 - 2.0 * s1p * t1 * tp + (double)n * tp * tp) / det;
#pragma omp parallel
  {
    double* g = (double*)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(static)
    for (int64_t j = 0; j < p; j++) {
      const double* a = A + j * lda;
      double s = 0.0;
      for (int64_t i = 0; i < n; i++) s += a[i];
      double m = s / (double)n, ss = 0.0;
      for (int64_t i = 0; i < n; i++) { double d = a[i] - m; ss += d * d; }
      double v = sqrt(ss / (double)(n - 1));
      if (!(v > 2.220446049250313e-16) || !isfinite(v)) {
        stat_ols[j] = NAN; if (stat_lmm) stat_lmm[j] = NAN; if (keep) keep[j] = 0;
        continue;
      }
      if (keep) keep[j] = 1;
      for (int64_t i = 0; i < n; i++) g[i] = (a[i] - m) / v;
      double s1g = 0, spg = 0, sgg = 0, tg = 0;
      for (int64_t i = 0; i < n; i++) { s1g += g[i]; spg += pc[i] * g[i]; sgg += g[i] * g[i]; tg += g[i] * ys[i]; }
      double XtX[3][3] = {{(double)n, s1p, s1g}, {s1p, spp, spg}, {s1g, spg, sgg}};
      double Vinv[3][3];
      pinv_sym3(XtX, Vinv);
      double b3 = Vinv[2][0] * t1 + Vinv[2][1] * tp + Vinv[2][2] * tg;
      double st = b3 / sqrt(Vinv[2][2]);
      stat_ols[j] = st;
      if (stat_lmm) stat_lmm[j] = st / sqrt((yMy - st * st) / (double)(n - 3));
    }
    free(g);
  }
}
