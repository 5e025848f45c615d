// Blocked right-looking Cholesky of the reduced camera system (row-major, lower triangle, in place).
//
// The reduced system of configuration C3 is 2402 x 2402 float64: cuSOLVER's potrf spends ~3.4 ms in ~250
// tiny launches on it, more than the Schur build once the tracks are sharded over GPUs.  This
// factorisation uses two launches per 64-column panel:
//   chol_panel_kernel     every CTA re-factors the 64x64 diagonal block in shared memory (cheap, removes a
//                         launch and a dependency), inverts it, and multiplies its 64 rows of the panel by
//                         L_kk^-T (dense 64^3 product instead of per-row substitution chains);
//                         CTA 0 parks the factored diagonal block in a side buffer.
//   chol_trailing_kernel  A22 -= P P^T on 128x128 tiles with the whole K=64 panel resident in shared memory
//                         (8x8 register tiles on the FP64 FMA pipe).
// and one copy-back of the diagonal blocks at the end.  Ceres' counterpart: DENSE_SCHUR's LLT / LAPACK potrf
// inside SchurComplementSolver (reached from pycolmap.bundle_adjustment).
#include "common.cuh"

namespace vgg {

constexpr int CH_NB = 64;

// grid.x = 1 + number of 64-row chunks below the diagonal block; block 256
__global__ void __launch_bounds__(256) chol_panel_kernel(int n, int lda, int k0, double* __restrict__ A,
                                                         double* __restrict__ Ldiag /*[nblk][64*64]*/,
                                                         int* __restrict__ info) {
  extern __shared__ __align__(16) double panel_smem[];
  double (*L)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(panel_smem);
  double (*Li)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(panel_smem + CH_NB * (CH_NB + 1));
  double (*T)[CH_NB + 1] = reinterpret_cast<double (*)[CH_NB + 1]>(panel_smem + 2 * CH_NB * (CH_NB + 1));
  __shared__ int fail;
  const int tid = threadIdx.x;
  const int nb = min(CH_NB, n - k0);
  if (tid == 0) fail = 0;
  // load the (unfactored) diagonal block, lower part; pad with identity
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int i = e / CH_NB, j = e % CH_NB;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j < nb && j <= i) v = A[(size_t)(k0 + i) * lda + k0 + j];
    L[i][j] = v;
    Li[i][j] = 0.0;
  }
  __syncthreads();
  // unblocked Cholesky in shared memory
  for (int j = 0; j < CH_NB; ++j) {
    if (tid == 0) {
      const double d = L[j][j];
      if (!(d > 0.0)) { fail = j + 1; L[j][j] = 1.0; }
      else L[j][j] = sqrt(d);
    }
    __syncthreads();
    const double dj = L[j][j];
    if (tid > j && tid < CH_NB) L[tid][j] /= dj;
    __syncthreads();
    // trailing update of the lower triangle: rows i > j, cols j < c <= i
    const int m = CH_NB - 1 - j;
    for (int e = tid; e < m * m; e += 256) {
      const int i = j + 1 + e / m, c = j + 1 + e % m;
      if (c <= i) L[i][c] -= L[i][j] * L[c][j];
    }
    __syncthreads();
  }
  if (fail && blockIdx.x == 0 && tid == 0) atomicCAS(info, 0, k0 + fail);
  // inverse of the lower-triangular factor: column j by thread j (forward substitution)
  if (tid < CH_NB) {
    const int j = tid;
    Li[j][j] = 1.0 / L[j][j];
    for (int i = j + 1; i < CH_NB; ++i) {
      double s = 0.0;
      for (int m2 = j; m2 < i; ++m2) s += L[i][m2] * Li[m2][j];
      Li[i][j] = -s / L[i][i];
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    double* dst = Ldiag + (size_t)(k0 / CH_NB) * CH_NB * CH_NB;
    for (int e = tid; e < CH_NB * CH_NB; e += 256) dst[e] = L[e / CH_NB][e % CH_NB];
    return;
  }
  // panel rows: X = A_ik L^-T  ->  X[r][j] = sum_{m<=j} A[r][m] Li[j][m]
  const int r0 = k0 + CH_NB + (blockIdx.x - 1) * CH_NB;
  for (int e = tid; e < CH_NB * CH_NB; e += 256) {
    const int r = e / CH_NB, c = e % CH_NB;
    T[r][c] = (r0 + r < n && c < nb) ? A[(size_t)(r0 + r) * lda + k0 + c] : 0.0;
  }
  __syncthreads();
  // 256 threads: thread -> (row r = tid/4, 16 columns j = (tid%4) + 4*jj)
  {
    const int r = tid >> 2, q = tid & 3;
    double acc[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0;
    for (int m2 = 0; m2 < CH_NB; ++m2) {
      const double a = T[r][m2];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) acc[jj] = fma(a, Li[q + 4 * jj][m2], acc[jj]);   // Li[j][m] = 0 for m > j
    }
    if (r0 + r < n) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int j = q + 4 * jj;
        if (j < nb) A[(size_t)(r0 + r) * lda + k0 + j] = acc[jj];
      }
    }
  }
}

// A[t0.., t0..] -= P P^T, P = A[t0.., k0..k0+63]; 128x128 tiles (lower), 256 threads, 8x8 per thread
__global__ void __launch_bounds__(256) chol_trailing_kernel(int n, int lda, int k0, int t0, double* __restrict__ A) {
  extern __shared__ __align__(16) double ch_smem[];
  double* As = ch_smem;                       // [64][128]
  double* Bs = ch_smem + CH_NB * 128;         // [64][128]
  int t = blockIdx.x;
  int bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  while (bi * (bi + 1) / 2 > t) --bi;
  const int bj = t - bi * (bi + 1) / 2;
  const bool diag = bi == bj;
  const int tid = threadIdx.x;
  const int ri = t0 + bi * 128, rj = t0 + bj * 128;
  // load panel rows (row-major, k contiguous) transposed into [k][row]: lane <-> row, 16 B per load
  {
    const int row = tid & 127, half = tid >> 7;
    for (int kp = half; kp < CH_NB / 2; kp += 2) {
      double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
      if (ri + row < n) va = *reinterpret_cast<const double2*>(A + (size_t)(ri + row) * lda + k0 + 2 * kp);
      As[(2 * kp) * 128 + row] = va.x;
      As[(2 * kp + 1) * 128 + row] = va.y;
      if (!diag) {
        if (rj + row < n) vb = *reinterpret_cast<const double2*>(A + (size_t)(rj + row) * lda + k0 + 2 * kp);
        Bs[(2 * kp) * 128 + row] = vb.x;
        Bs[(2 * kp + 1) * 128 + row] = vb.y;
      }
    }
  }
  __syncthreads();
  const double* bs = diag ? As : Bs;
  const int ty = tid >> 4, tx = tid & 15;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < CH_NB; ++kk) {
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 av = *reinterpret_cast<const double2*>(As + kk * 128 + ty * 2 + 32 * i);
      a[2 * i] = av.x; a[2 * i + 1] = av.y;
      const double2 bv = *reinterpret_cast<const double2*>(bs + kk * 128 + tx * 2 + 32 * i);
      b[2 * i] = bv.x; b[2 * i + 1] = bv.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ri + ty * 2 + (i & 1) + 32 * (i >> 1);
    if (r >= n) continue;
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const int c = rj + tx * 2 + 32 * jp;
      if (c > r || c >= n) continue;           // lower triangle only (c even: pair c, c+1)
      double* p = A + (size_t)r * lda + c;
      double2 v = *reinterpret_cast<double2*>(p);
      v.x -= acc[i][2 * jp];
      v.y -= acc[i][2 * jp + 1];               // element (r, c+1) may sit above the diagonal: harmless
      *reinterpret_cast<double2*>(p) = v;
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
int chol_lower_inplace(int n, int lda, double* A, double* Ldiag, int* info, cudaStream_t st) {
  VGG_REQUIRE((lda % 2) == 0, "lda must be even");
  VGG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), st));
  const int nblk = (n + CH_NB - 1) / CH_NB;
  const size_t smem = sizeof(double) * 2 * CH_NB * 128;
  const size_t psmem = sizeof(double) * 3 * CH_NB * (CH_NB + 1);
  static bool attr_set = false;
  if (!attr_set) {
    VGG_CUDA_CHECK(cudaFuncSetAttribute(chol_trailing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGG_CUDA_CHECK(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    attr_set = true;
  }
  for (int b = 0; b < nblk; ++b) {
    const int k0 = b * CH_NB;
    const int below = n - (k0 + CH_NB);
    const int chunks = below > 0 ? (below + CH_NB - 1) / CH_NB : 0;
    chol_panel_kernel<<<1 + chunks, 256, psmem, st>>>(n, lda, k0, A, Ldiag, info);
    VGG_LAUNCH_CHECK();
    if (below > 0) {
      const int t0 = k0 + CH_NB;
      const int nt = (n - t0 + 127) / 128;
      chol_trailing_kernel<<<nt * (nt + 1) / 2, 256, smem, st>>>(n, lda, k0, t0, A);
      VGG_LAUNCH_CHECK();
    }
  }
  chol_copy_diag_kernel<<<nblk, 256, 0, st>>>(n, lda, Ldiag, A);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}

}  // namespace vgg
