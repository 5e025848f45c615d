/* CPU restatement (plain C + OpenMP) of the reprojection residual / analytic Jacobian / normal-equation
 * block evaluation -- TEST INFRASTRUCTURE ONLY (oracle).  Same arithmetic as oracle/ba_oracle.py
 * residuals_and_jacobians()/build_blocks(), which restates COLMAP 3.10's ReprojErrorCostFunction for
 * SIMPLE_PINHOLE / SIMPLE_RADIAL as reached from vggsfm/utils/triangulation.py:1050,1142 [3P-memory];
 * the in-repo statement of the same projection model is vggsfm/utils/triangulation_helpers.py:358-395 and
 * vggsfm/utils/distortion.py:119-128.  Used to make the timed CPU baseline of bench.py a fair one (hand
 * Jacobians, all host threads) and cross-checked against the numpy restatement in tests/test_ba_oracle.py.
 * PARITY UNPINNED like the rest of the BA oracle (pycolmap absent).
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -shared -fPIC ba_blocks_ref.c -o _build/libba_blocks_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* model: 0 pinhole, 1 radial; mode: 0 const, 1 per-frame, 2 shared.
 * Outputs (zeroed here): cost[1], g_c[S*dc], H_cc[S*dc*dc], g_p[N*3], H_pp[N*9], W[S*dc*N*3],
 * g_s[2], H_ss[4], H_cs[S*6*2] (stride ns), W_s[ns*N*3]. */
int ba_blocks_ref(int S, int N, const double* uv, const uint8_t* mask, const double* poses, const double* intr,
                  const double* points, const uint8_t* point_const, int model, int mode, double* cost, double* g_c,
                  double* H_cc, double* g_p, double* H_pp, double* W, double* g_s, double* H_ss, double* H_cs,
                  double* W_s) {
  const int ni = model == 0 ? 1 : 2;
  const int dc = mode == 1 ? 6 + ni : 6;
  const int ns = mode == 2 ? ni : 0;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  const size_t pstride = (size_t)N * (3 + 9 + 3 * 2);
  double* pacc = (double*)calloc((size_t)nthreads * pstride, sizeof(double));
  double* sacc = (double*)calloc((size_t)nthreads * 8, sizeof(double));
  if (!pacc || !sacc) return -1;
  memset(g_c, 0, sizeof(double) * S * dc);
  memset(H_cc, 0, sizeof(double) * S * dc * dc);
  if (ns) memset(H_cs, 0, sizeof(double) * S * 6 * ns);
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* pa = pacc + (size_t)tid * pstride;
    double* sa = sacc + (size_t)tid * 8;
#pragma omp for schedule(static)
    for (int s = 0; s < S; ++s) {
      const double* P = poses + (size_t)s * 12;
      const double f = intr[s * 4], cx = intr[s * 4 + 1], cy = intr[s * 4 + 2];
      const double k = model == 1 ? intr[s * 4 + 3] : 0.0;
      double* gc = g_c + (size_t)s * dc;
      double* Hc = H_cc + (size_t)s * dc * dc;
      for (int n = 0; n < N; ++n) {
        double* Wb = W + (((size_t)s * dc) * N + n) * 3;
        if (!mask[(size_t)s * N + n]) {
          for (int i = 0; i < dc; ++i) Wb[(size_t)i * N * 3] = Wb[(size_t)i * N * 3 + 1] = Wb[(size_t)i * N * 3 + 2] = 0.0;
          continue;
        }
        const double* X = points + (size_t)n * 3;
        const double a1 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2];
        const double a2 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2];
        const double a3 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2];
        const double px = a1 + P[3], py = a2 + P[7], pz = a3 + P[11];
        const double iz = 1.0 / pz, u = px * iz, v = py * iz;
        const double r2 = u * u + v * v, d = 1.0 + k * r2;
        const double rx = f * d * u + cx - uv[((size_t)s * N + n) * 2];
        const double ry = f * d * v + cy - uv[((size_t)s * N + n) * 2 + 1];
        sa[0] += 0.5 * (rx * rx + ry * ry);
        const double a00 = f * (d + 2.0 * k * u * u), a01 = f * (2.0 * k * u * v), a11 = f * (d + 2.0 * k * v * v);
        const double j00 = a00 * iz, j01 = a01 * iz, j02 = -(a00 * u + a01 * v) * iz;
        const double j10 = a01 * iz, j11 = a11 * iz, j12 = -(a01 * u + a11 * v) * iz;
        double jc0[8], jc1[8], jx0[3], jx1[3];
        jc0[0] = 2.0 * (-a3 * j01 + a2 * j02); jc1[0] = 2.0 * (-a3 * j11 + a2 * j12);
        jc0[1] = 2.0 * (a3 * j00 - a1 * j02);  jc1[1] = 2.0 * (a3 * j10 - a1 * j12);
        jc0[2] = 2.0 * (-a2 * j00 + a1 * j01); jc1[2] = 2.0 * (-a2 * j10 + a1 * j11);
        jc0[3] = j00; jc0[4] = j01; jc0[5] = j02; jc1[3] = j10; jc1[4] = j11; jc1[5] = j12;
        jc0[6] = d * u; jc1[6] = d * v; jc0[7] = f * u * r2; jc1[7] = f * v * r2;
        for (int c = 0; c < 3; ++c) {
          jx0[c] = j00 * P[c] + j01 * P[4 + c] + j02 * P[8 + c];
          jx1[c] = j10 * P[c] + j11 * P[4 + c] + j12 * P[8 + c];
        }
        if (point_const && point_const[n]) jx0[0] = jx0[1] = jx0[2] = jx1[0] = jx1[1] = jx1[2] = 0.0;
        double* gp = pa + (size_t)n * 18;
        for (int c = 0; c < 3; ++c) {
          gp[c] += jx0[c] * rx + jx1[c] * ry;
          for (int e = 0; e < 3; ++e) gp[3 + c * 3 + e] += jx0[c] * jx0[e] + jx1[c] * jx1[e];
        }
        for (int i = 0; i < dc; ++i) {
          gc[i] += jc0[i] * rx + jc1[i] * ry;
          for (int j = 0; j < dc; ++j) Hc[i * dc + j] += jc0[i] * jc0[j] + jc1[i] * jc1[j];
          for (int c = 0; c < 3; ++c) Wb[(size_t)i * N * 3 + c] = jc0[i] * jx0[c] + jc1[i] * jx1[c];
        }
        for (int j = 0; j < ns; ++j) {
          sa[1 + j] += jc0[6 + j] * rx + jc1[6 + j] * ry;
          for (int e = 0; e < ns; ++e) sa[3 + j * 2 + e] += jc0[6 + j] * jc0[6 + e] + jc1[6 + j] * jc1[6 + e];
          for (int i = 0; i < 6; ++i) H_cs[((size_t)s * 6 + i) * ns + j] += jc0[i] * jc0[6 + j] + jc1[i] * jc1[6 + j];
          for (int c = 0; c < 3; ++c) gp[12 + j * 3 + c] += jc0[6 + j] * jx0[c] + jc1[6 + j] * jx1[c];
        }
      }
    }
  }
  *cost = 0.0;
  memset(g_p, 0, sizeof(double) * N * 3);
  memset(H_pp, 0, sizeof(double) * N * 9);
  if (ns) { memset(W_s, 0, sizeof(double) * ns * N * 3); memset(g_s, 0, sizeof(double) * 2); memset(H_ss, 0, sizeof(double) * 4); }
  for (int t = 0; t < nthreads; ++t) {
    const double* pa = pacc + (size_t)t * pstride;
    const double* sa = sacc + (size_t)t * 8;
    *cost += sa[0];
    for (int j = 0; j < ns; ++j) {
      g_s[j] += sa[1 + j];
      for (int e = 0; e < ns; ++e) H_ss[j * ns + e] += sa[3 + j * 2 + e];
    }
    for (int n = 0; n < N; ++n) {
      for (int c = 0; c < 3; ++c) g_p[n * 3 + c] += pa[(size_t)n * 18 + c];
      for (int c = 0; c < 9; ++c) H_pp[n * 9 + c] += pa[(size_t)n * 18 + 3 + c];
      for (int j = 0; j < ns; ++j)
        for (int c = 0; c < 3; ++c) W_s[((size_t)j * N + n) * 3 + c] += pa[(size_t)n * 18 + 12 + j * 3 + c];
    }
  }
  free(pacc);
  free(sacc);
  return 0;
}
