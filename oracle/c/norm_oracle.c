/*
 * norm_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into libb200norm.so).
 *
 * An independent fp64 restatement, in plain C, of the math of the hot path as
 * the reference's op builders define it (csrc/ops/normalization.cpp):
 *   softmax :169-189, softmax_backward :191-210, log_softmax :212-231,
 *   log_softmax_backward :233-249, layer_norm :304-342, rms_norm :352-380,
 *   layer_norm_backward :382-436, rms_norm_backward :438-485,
 *   thunder_rms_norm_backward :487-535 (in the textbook form the Python trace
 *   benchmarks/python/test_rmsnorm_bwd.py:20-76 computes).
 * It exists to pin the ATen oracle (oracle/aten_oracle.py) from a second,
 * library-free direction and to give the GPU tests an "<= k ulp of the io dtype
 * against exact arithmetic" bound that is much tighter than the reference's
 * validator tolerances. All inputs/outputs are row-major doubles; callers
 * convert bf16/fp16/fp32 data exactly (every such value is a double).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define API __attribute__((visibility("default")))

API void oracle_layer_norm_fwd(const double* x, const double* w, const double* b, int64_t M, int64_t N,
                               double eps, double* y, double* mean, double* invstd) {
  for (int64_t r = 0; r < M; ++r) {
    const double* xr = x + r * N;
    double s = 0.0;
    for (int64_t c = 0; c < N; ++c) s += xr[c];
    const double mu = s / (double)N;
    double m2 = 0.0;
    for (int64_t c = 0; c < N; ++c) m2 += (xr[c] - mu) * (xr[c] - mu);
    const double is = 1.0 / sqrt(m2 / (double)N + eps);
    for (int64_t c = 0; c < N; ++c) {
      double v = (xr[c] - mu) * is;
      if (w) v *= w[c];
      if (b) v += b[c];
      y[r * N + c] = v;
    }
    if (mean) mean[r] = mu;
    if (invstd) invstd[r] = is;
  }
}

/* stat_is_rms: 0 -> stat = rsqrt(mean(x^2)+eps), 1 -> stat = sqrt(mean(x^2)+eps) */
API void oracle_rms_norm_fwd(const double* x, const double* w, int64_t M, int64_t N, double eps,
                             int stat_is_rms, double* y, double* stat) {
  for (int64_t r = 0; r < M; ++r) {
    const double* xr = x + r * N;
    double ss = 0.0;
    for (int64_t c = 0; c < N; ++c) ss += xr[c] * xr[c];
    const double rms = sqrt(ss / (double)N + eps);
    for (int64_t c = 0; c < N; ++c) {
      double v = xr[c] / rms;
      if (w) v *= w[c];
      y[r * N + c] = v;
    }
    if (stat) stat[r] = stat_is_rms ? rms : 1.0 / rms;
  }
}

API void oracle_softmax_fwd(const double* x, int64_t M, int64_t N, int is_log, double* y) {
  for (int64_t r = 0; r < M; ++r) {
    const double* xr = x + r * N;
    double mx = -INFINITY;
    int has_nan = 0;
    for (int64_t c = 0; c < N; ++c) {
      if (xr[c] != xr[c]) has_nan = 1;
      if (xr[c] > mx) mx = xr[c];
    }
    if (has_nan) mx = NAN; /* at::amax propagates NaN */
    double s = 0.0;
    for (int64_t c = 0; c < N; ++c) s += exp(xr[c] - mx);
    for (int64_t c = 0; c < N; ++c) {
      const double d = xr[c] - mx;
      y[r * N + c] = is_log ? d - log(s) : exp(d) / s;
    }
  }
}

API void oracle_softmax_bwd(const double* dy, const double* y, int64_t M, int64_t N, int is_log, double* dx) {
  for (int64_t r = 0; r < M; ++r) {
    const double* g = dy + r * N;
    const double* yr = y + r * N;
    double s = 0.0;
    for (int64_t c = 0; c < N; ++c) s += is_log ? g[c] : g[c] * yr[c];
    for (int64_t c = 0; c < N; ++c) {
      dx[r * N + c] = is_log ? g[c] - exp(yr[c]) * s : yr[c] * g[c] - yr[c] * s;
    }
  }
}

/*
 * Normalization backward.
 *   mean == NULL  -> RMSNorm (x_hat = x * invstd)
 *   stat_is_rms   -> `invstd` holds rms; invert it first
 *   centered      -> keep the -sum(g) term (always for LayerNorm; the C++
 *                    rms_norm_backward builder keeps it too, :461-476)
 * dx = invstd/N * (N*g - [centered] sum(g) - x_hat * sum(g*x_hat))
 * dw = sum_rows dy*x_hat ; db = sum_rows dy
 */
API void oracle_norm_bwd(const double* dy, const double* x, const double* mean, const double* invstd,
                         const double* w, int64_t M, int64_t N, int stat_is_rms, int centered, double* dx,
                         double* dw, double* db) {
  if (dw) for (int64_t c = 0; c < N; ++c) dw[c] = 0.0;
  if (db) for (int64_t c = 0; c < N; ++c) db[c] = 0.0;
  for (int64_t r = 0; r < M; ++r) {
    const double* g0 = dy + r * N;
    const double* xr = x + r * N;
    const double mu = mean ? mean[r] : 0.0;
    const double is = stat_is_rms ? 1.0 / invstd[r] : invstd[r];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t c = 0; c < N; ++c) {
      const double xh = (xr[c] - mu) * is;
      const double g = w ? g0[c] * w[c] : g0[c];
      s1 += g;
      s2 += g * xh;
      if (dw) dw[c] += g0[c] * xh;
      if (db) db[c] += g0[c];
    }
    if (!centered) s1 = 0.0;
    if (dx) {
      for (int64_t c = 0; c < N; ++c) {
        const double xh = (xr[c] - mu) * is;
        const double g = w ? g0[c] * w[c] : g0[c];
        dx[r * N + c] = (is / (double)N) * (((double)N * g - s1) - xh * s2);
      }
    }
  }
}

/* reductions: kind 0 = sum, 1 = amax (NaN-propagating), 2 = fmax (NaN-ignoring) */
API void oracle_reduce_inner(const double* x, int64_t M, int64_t N, int kind, double* out) {
  for (int64_t r = 0; r < M; ++r) {
    const double* xr = x + r * N;
    if (kind == 0) {
      double s = 0.0;
      for (int64_t c = 0; c < N; ++c) s += xr[c];
      out[r] = s;
    } else {
      double mx = -INFINITY;
      int has_nan = 0, n_ok = 0;
      for (int64_t c = 0; c < N; ++c) {
        if (xr[c] != xr[c]) { has_nan = 1; continue; }
        n_ok++;
        if (xr[c] > mx) mx = xr[c];
      }
      if (kind == 1) out[r] = has_nan ? NAN : mx;
      else out[r] = n_ok ? mx : NAN;
    }
  }
}

API void oracle_reduce_outer(const double* x, int64_t M, int64_t N, double* out) {
  for (int64_t c = 0; c < N; ++c) out[c] = 0.0;
  for (int64_t r = 0; r < M; ++r)
    for (int64_t c = 0; c < N; ++c) out[c] += x[r * N + c];
}
