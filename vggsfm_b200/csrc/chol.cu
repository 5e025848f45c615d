// Blocked right-looking Cholesky of the reduced camera system (row-major, lower triangle, in place).
//
// The reduced system of configuration C3 is 2402 x 2402 float64: cuSOLVER's potrf spends ~3.4 ms in ~250
// tiny launches on it, more than the Schur build once the tracks are sharded over GPUs.  This
// factorisation uses two launches per 64-column panel:
//   chol_panel_kernel     every CTA re-factors the 64x64 diagonal block in shared memory in 8-column
//                         micro-panels (8x8 leaf in registers with shuffles + rsqrt, 3 CTA barriers per
//                         micro-panel), then a right-looking triangular solve, 4 threads per row, turns its 64
//                         panel rows into L_ik.  CTA 0 parks the factored diagonal block in a side buffer.
//   chol_trailing_kernel  A22 -= P P^T on 64x64 tiles with the whole K=64 panel resident in shared memory.
// and one copy-back of the diagonal blocks at the end.  Ceres' counterpart: DENSE_SCHUR's LLT / LAPACK potrf
// inside SchurComplementSolver (reached from pycolmap.bundle_adjustment).
#include <stdlib.h>
#include "common.cuh"

namespace vgg {

constexpr int CH_NB = 64;
constexpr int CH_LD = 66;     // shared-memory row stride (even: 16-byte aligned pairs; 66*2 mod 32 = 4: conflict-free)

// Cholesky of the 64x64 block in shared memory Ls (row stride CH_LD) by the whole CTA (256 threads), in
// 8-column micro-panels: (a) warp 0 factors the 8x8 diagonal micro-block in registers (one row per lane,
// shuffles for the pivot column, rsqrt instead of sqrt+divide), (b) one thread per row below solves its 8
// entries against it, (c) all threads apply the rank-8 update to the rest of the block.  Three CTA barriers per
// micro-panel instead of three per column.  dinv[j] = 1 / L[j][j].  Returns 0 or 1 + first bad pivot (all threads).
__device__ __forceinline__ int cta_chol64(double* Ls, double* dinv, int* fail_sm, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  if (tid == 0) *fail_sm = 0;
  __syncthreads();
  for (int p = 0; p < 8; ++p) {
    const int c0 = p * 8;
    if (warp == 0) {
      double a[8];
      const int r = c0 + (lane & 7);
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = Ls[r * CH_LD + c0 + c];
      int fail = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double pj = __shfl_sync(0xffffffffu, a[j], j);
        if (!(pj > 0.0) && fail == 0) fail = c0 + j + 1;
        const double inv = rsqrt(pj > 0.0 ? pj : 1.0);
        a[j] = (lane == j) ? pj * inv : a[j] * inv;
        if (lane == j) dinv[c0 + j] = inv;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) {
          const double lc = __shfl_sync(0xffffffffu, a[j], c);
          if (lane >= c) a[c] = fma(-a[j], lc, a[c]);
        }
      }
      if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) Ls[r * CH_LD + c0 + c] = (c <= lane) ? a[c] : 0.0;
      }
      if (lane == 0 && fail && *fail_sm == 0) *fail_sm = fail;
    }
    __syncthreads();
    // (b) rows below the micro-block: x L^T = a  (one thread per row)
    {
      const int r = c0 + 8 + tid;
      if (r < CH_NB) {
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = Ls[r * CH_LD + c0 + c];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          x[m] *= dinv[c0 + m];
#pragma unroll
          for (int j = m + 1; j < 8; ++j) x[j] = fma(-x[m], Ls[(c0 + j) * CH_LD + c0 + m], x[j]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) Ls[r * CH_LD + c0 + c] = x[c];
      }
    }
    __syncthreads();
    // (c) rank-8 update of the remaining lower triangle: rows/cols >= c0+8
    {
      const int m = CH_NB - (c0 + 8);           // remaining dimension
      for (int e = tid; e < m * m; e += 256) {
        const int i = e / m, c = e - i * m;
        if (c <= i) {
          const double* li = Ls + (c0 + 8 + i) * CH_LD + c0;
          const double* lc = Ls + (c0 + 8 + c) * CH_LD + c0;
          double s = Ls[(c0 + 8 + i) * CH_LD + c0 + 8 + c];
#pragma unroll
          for (int k = 0; k < 8; ++k) s = fma(-li[k], lc[k], s);
          Ls[(c0 + 8 + i) * CH_LD + c0 + 8 + c] = s;
        }
      }
    }
    __syncthreads();
  }
  return *fail_sm;
}

// grid.x = 1 + number of 64-row chunks below the diagonal block; block 256
__global__ void __launch_bounds__(256) chol_panel_kernel(int n, int lda, int k0, double* __restrict__ A,
                                                         double* __restrict__ Ldiag /*[nblk][64*64]*/,
                                                         int* __restrict__ info) {
  extern __shared__ __align__(16) double panel_smem[];
  double* Ls = panel_smem;
  double* Ts = panel_smem + CH_NB * CH_LD;
  __shared__ double dinv[CH_NB];
  __shared__ int fail_sm;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nb = min(CH_NB, n - k0);
  const int r0 = k0 + CH_NB + ((int)blockIdx.x - 1) * CH_NB;
  // diagonal block (lower part, identity padding) and this CTA's 64 panel rows: a warp reads one 512 B row
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int i = e >> 6, j = e & 63;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j < nb && j <= i) v = A[(size_t)(k0 + i) * lda + k0 + j];
    Ls[i * CH_LD + j] = v;
    if (blockIdx.x > 0) Ts[i * CH_LD + j] = (r0 + i < n && j < nb) ? A[(size_t)(r0 + i) * lda + k0 + j] : 0.0;
  }
  const int fail = cta_chol64(Ls, dinv, &fail_sm, tid);
  if (fail && blockIdx.x == 0 && tid == 0) atomicCAS(info, 0, k0 + fail);
  if (blockIdx.x == 0) {
    double* dst = Ldiag + (size_t)(k0 / CH_NB) * CH_NB * CH_NB;
    for (int e = tid; e < CH_NB * CH_NB; e += 256) dst[e] = Ls[(e >> 6) * CH_LD + (e & 63)];
    return;
  }
  // X L^T = A_ik, right-looking: 4 threads per row, thread q owns columns j = q + 4*jj
  {
    const int r = tid >> 2, q = tid & 3;
    double a[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) a[jj] = Ts[r * CH_LD + q + 4 * jj];
#pragma unroll
    for (int m = 0; m < CH_NB; ++m) {
      const int qm = m & 3, jm = m >> 2;
      double x = a[jm] * dinv[m];
      x = __shfl_sync(0xffffffffu, x, (lane & ~3) | qm);
      if (q == qm) a[jm] = x;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        if (4 * jj + 3 > m) {                       // compile-time prune; exact test below
          const int j = q + 4 * jj;
          if (j > m) a[jj] = fma(-x, Ls[j * CH_LD + m], a[jj]);
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) Ts[r * CH_LD + q + 4 * jj] = a[jj];
  }
  __syncthreads();
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (r0 + r < n && c < nb) A[(size_t)(r0 + r) * lda + k0 + c] = Ts[r * CH_LD + c];
  }
}

// A[t0.., t0..] -= P P^T, P = A[t0.., k0..k0+63]; 64x64 tiles (lower), 256 threads, 4x4 outputs per thread
// mode 0: all lower tiles; mode 1: first tile column only (the next panel's columns); mode 2: the tiles right of it
__global__ void __launch_bounds__(256) chol_trailing_kernel(int n, int lda, int k0, int t0, int mode,
                                                            double* __restrict__ A) {
  extern __shared__ __align__(16) double ch_smem[];
  double* As = ch_smem;                         // [64 rows][CH_LD]
  double* Bs = ch_smem + CH_NB * CH_LD;
  int bi, bj;
  if (mode == 1) {
    bi = blockIdx.x; bj = 0;
  } else {
    const int t = blockIdx.x;
    bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    bj = t - bi * (bi + 1) / 2;
    if (mode == 2) { ++bi; ++bj; }
  }
  const bool diag = bi == bj;
  const int tid = threadIdx.x;
  const int ri = t0 + bi * 64, rj = t0 + bj * 64;
  // panel rows, row-major (k contiguous): a warp loads one 512 B row per step as 32 double2
  for (int e = tid; e < 64 * 32; e += 256) {
    const int row = e >> 5, kp = e & 31;
    double2 va = make_double2(0.0, 0.0);
    if (ri + row < n) va = *reinterpret_cast<const double2*>(A + (size_t)(ri + row) * lda + k0 + 2 * kp);
    *reinterpret_cast<double2*>(As + row * CH_LD + 2 * kp) = va;
    if (!diag) {
      double2 vb = make_double2(0.0, 0.0);
      if (rj + row < n) vb = *reinterpret_cast<const double2*>(A + (size_t)(rj + row) * lda + k0 + 2 * kp);
      *reinterpret_cast<double2*>(Bs + row * CH_LD + 2 * kp) = vb;
    }
  }
  __syncthreads();
  const double* bs = diag ? As : Bs;
  const int ty = tid >> 4, tx = tid & 15;         // rows ty + 16 i, cols tx + 16 j
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
  for (int kk = 0; kk < CH_NB; kk += 2) {
    double2 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = *reinterpret_cast<const double2*>(As + (ty + 16 * i) * CH_LD + kk);
      b[i] = *reinterpret_cast<const double2*>(bs + (tx + 16 * i) * CH_LD + kk);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fma(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fma(a[i].y, b[j].y, acc[i][j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ri + ty + 16 * i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = rj + tx + 16 * j;
      if (c > r || c >= n) continue;              // lower triangle only
      A[(size_t)r * lda + c] -= acc[i][j];
    }
  }
}

__global__ void chol_copy_diag_kernel(int n, int lda, const double* __restrict__ Ldiag, double* __restrict__ A) {
  const int blk = blockIdx.x;
  const int k0 = blk * CH_NB;
  for (int e = threadIdx.x; e < CH_NB * CH_NB; e += blockDim.x) {
    const int i = e / CH_NB, j = e % CH_NB;
    if (k0 + i < n && j <= i) A[(size_t)(k0 + i) * lda + k0 + j] = Ldiag[(size_t)blk * CH_NB * CH_NB + e];
  }
}

size_t chol_workspace_doubles(int n) {
  const int nblk = (n + CH_NB - 1) / CH_NB;
  return (size_t)nblk * CH_NB * CH_NB;
}

// In-place Cholesky of the row-major lower triangle of A[n x n] (lda even, A 16-byte aligned).
// info (device int): 0 on success, else 1-based index of the first non-positive pivot.
// One-panel lookahead: the trailing update of panel b first refreshes the next panel's 64 columns; panel b+1 is
// then factored on a side stream while the main stream finishes the rest of the update (the panel kernel is a
// latency chain on ~40 CTAs, the update is throughput work on the whole chip).
int chol_lower_inplace(int n, int lda, double* A, double* Ldiag, int* info, cudaStream_t st) {
  VGG_REQUIRE((lda % 2) == 0, "lda must be even");
  VGG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), st));
  const int nblk = (n + CH_NB - 1) / CH_NB;
  const size_t smem = sizeof(double) * 2 * CH_NB * CH_LD;
  static bool attr_set = false;
  if (!attr_set) {
    VGG_CUDA_CHECK(cudaFuncSetAttribute(chol_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGG_CUDA_CHECK(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static thread_local cudaStream_t side = nullptr;
  static thread_local cudaEvent_t ev_col = nullptr, ev_panel = nullptr;
  static const bool lookahead = [] { const char* e = getenv("VGG_CHOL_LOOKAHEAD"); return !(e && e[0] == '0'); }();
  if (lookahead && !side) {
    VGG_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    VGG_CUDA_CHECK(cudaEventCreateWithFlags(&ev_col, cudaEventDisableTiming));
    VGG_CUDA_CHECK(cudaEventCreateWithFlags(&ev_panel, cudaEventDisableTiming));
  }
  auto panel = [&](int b, cudaStream_t s2) -> int {
    const int k0 = b * CH_NB;
    const int below = n - (k0 + CH_NB);
    const int chunks = below > 0 ? (below + CH_NB - 1) / CH_NB : 0;
    chol_panel_kernel<<<1 + chunks, 256, smem, s2>>>(n, lda, k0, A, Ldiag, info);
    VGG_LAUNCH_CHECK();
    return VGG_OK;
  };
  int rc;
  if ((rc = panel(0, st))) return rc;
  for (int b = 0; b + 1 < nblk; ++b) {
    const int k0 = b * CH_NB, t0 = k0 + CH_NB;
    const int nt = (n - t0 + 63) / 64;
    if (!lookahead) {
      chol_trailing_kernel<<<nt * (nt + 1) / 2, 256, smem, st>>>(n, lda, k0, t0, 0, A);
      VGG_LAUNCH_CHECK();
      if ((rc = panel(b + 1, st))) return rc;
      continue;
    }
    chol_trailing_kernel<<<nt, 256, smem, st>>>(n, lda, k0, t0, 1, A);
    VGG_LAUNCH_CHECK();
    VGG_CUDA_CHECK(cudaEventRecord(ev_col, st));
    VGG_CUDA_CHECK(cudaStreamWaitEvent(side, ev_col, 0));
    if ((rc = panel(b + 1, side))) return rc;
    VGG_CUDA_CHECK(cudaEventRecord(ev_panel, side));
    if (nt > 1) {
      chol_trailing_kernel<<<(nt - 1) * nt / 2, 256, smem, st>>>(n, lda, k0, t0, 2, A);
      VGG_LAUNCH_CHECK();
    }
    VGG_CUDA_CHECK(cudaStreamWaitEvent(st, ev_panel, 0));
  }
  chol_copy_diag_kernel<<<nblk, 256, 0, st>>>(n, lda, Ldiag, A);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}

}  // namespace vgg
