// Fused reprojection residual + analytic Jacobian + normal-equation block kernel.
//
// Replaces, per LM iteration, what Ceres does on the CPU for the reference: evaluate COLMAP's
// ReprojErrorCostFunction<SimplePinhole|SimpleRadial> for every observation with autodiff Jets and
// hand the 2x(dc+3) Jacobians to the Schur eliminator (reached from
// vggsfm/utils/triangulation.py:213,1050,1142 through pycolmap.bundle_adjustment).
//
// One thread owns one track (point) and marches over a chunk of frames; observations are the dense
// [S,N] grid the reference already uses (tracks [S,N,2] + inlier mask [S,N]).
//   * HBM reads: uv (8 B) + mask (1 B) per observation, staged per 8-frame sub-tile with TMA 1-D bulk
//     copies (cp.async.bulk -> UBLKCP) into a two-stage shared-memory ring with mbarriers; poses and
//     intrinsics ride the same barrier.
//   * HBM writes: the camera-point coupling block W = J_c^T J_p (dc x 3 doubles per observation),
//     assembled per frame and per warp in shared memory as dc rows of [32 x 3] and written back with TMA
//     bulk stores (768 B rows), double buffered against the next frame's math; warps never wait for each
//     other inside a TS-frame sub-tile (no CTA barrier in the frame loop).
//   * per-point blocks (H_pp 3x3 sym, g_p) accumulate in registers over the frame chunk;
//   * per-camera blocks (g_c, H_cc upper-packed, H_cs) are reduced across the 32 tracks of a warp with
//     a reduce-scatter shuffle network (16 values at a time, 15+1 shuffles), then one f64 RED per value
//     per warp.
// Algorithmic HBM bytes per observation: 9 + 24*dc (+ amortised per-frame/per-point terms), see DESIGN.md.
#include <utility>
#include "common.cuh"

namespace vgg {

constexpr int TN = 128;   // tracks per CTA (threads)
constexpr int TS = 4;     // frames per TMA sub-tile (keeps smem <= 56 KB -> 4 CTAs/SM)
constexpr int NWARP = TN / 32;

template <int MODEL, int MODE>
struct BlkCfg {
  static constexpr int NI = (MODEL == VGG_SIMPLE_PINHOLE) ? 1 : 2;
  static constexpr int DC = (MODE == VGG_INTR_PER_FRAME) ? 6 + NI : 6;
  static constexpr int NS = (MODE == VGG_INTR_SHARED) ? NI : 0;
  static constexpr int NPACK = DC * (DC + 1) / 2;
  static constexpr int KR = DC + NPACK + 6 * NS;   // per-frame reduced values
  static constexpr int K1 = KR < 32 ? KR : 32;
  static constexpr int K2 = KR - K1;               // second group (<= 16)
};

struct BlkSmem {
  // input ring
  float2 uv[2][TS][TN];
  double pose[2][TS][12];
  double intr[2][TS][4];
  uint8_t mask[2][TS][TN];
  uint64_t bar[2];
};

// ---- compile-time layout of the per-frame camera record: g_c[DC] | H_cc upper-packed | H_cs[6][NS] ----
__host__ __device__ constexpr int pack_row(int dc, int p) {
  int i = 0;
  while (p >= dc - i) { p -= dc - i; ++i; }
  return i;
}
__host__ __device__ constexpr int pack_col(int dc, int p) {
  int i = 0;
  while (p >= dc - i) { p -= dc - i; ++i; }
  return i + p;
}

template <int DC, int NS, int K>
__device__ __forceinline__ double cam_value(const double* jc0, const double* jc1, double rx, double ry) {
  constexpr int NPACK = DC * (DC + 1) / 2;
  if constexpr (K < DC) {
    return jc0[K] * rx + jc1[K] * ry;
  } else if constexpr (K < DC + NPACK) {
    constexpr int i = pack_row(DC, K - DC), j = pack_col(DC, K - DC);
    return jc0[i] * jc0[j] + jc1[i] * jc1[j];
  } else {
    constexpr int q = K - DC - NPACK;
    constexpr int i = q / (NS > 0 ? NS : 1), j = q % (NS > 0 ? NS : 1);
    return jc0[i] * jc0[6 + j] + jc1[i] * jc1[6 + j];
  }
}

template <int DC, int NS, int KR, int BASE, int... G>
__device__ __forceinline__ void cam_batch(double (&a)[16], const double* jc0, const double* jc1, double rx, double ry,
                                          std::integer_sequence<int, G...>) {
  ((a[G] = (BASE + G < KR) ? cam_value<DC, NS, (BASE + G < KR ? BASE + G : 0)>(jc0, jc1, rx, ry) : 0.0), ...);
}

template <int MODEL, int MODE, bool USE_TMA>
__global__ void __launch_bounds__(TN) ba_blocks_kernel(
    int S, int N, int frames_per_cta, const float2* __restrict__ uv, const uint8_t* __restrict__ mask,
    const double* __restrict__ poses, const double* __restrict__ intr, const double* __restrict__ points,
    const uint8_t* __restrict__ point_const, double* __restrict__ cost, double* __restrict__ camrec,
    double* __restrict__ g_p, double* __restrict__ H_pp, double* __restrict__ W, double* __restrict__ shared_out) {
  using C = BlkCfg<MODEL, MODE>;
  constexpr int DC = C::DC, NS = C::NS, KR = C::KR;
  constexpr int NB16 = (KR + 15) / 16;          // reduce batches of 16 camera values
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BlkSmem& sm = *reinterpret_cast<BlkSmem*>(smem_raw);
  // per-warp W tiles: [NWARP][2][DC][32*3] doubles, then red[NWARP][8]
  double* wsm = reinterpret_cast<double*>(smem_raw + align_up(sizeof(BlkSmem), 128));
  double* red = wsm + NWARP * 2 * DC * 96;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * TN;
  const int n = n0 + tid;
  const int nvalid = min(TN, N - n0);
  const int nw0 = n0 + warp * 32;                       // first track of this warp
  const int nvalid_w = max(0, min(32, N - nw0));
  const int s_begin = blockIdx.y * frames_per_cta;
  const int s_end = min(S, s_begin + frames_per_cta);
  const int ntiles = (s_end - s_begin + TS - 1) / TS;
  const bool active = n < N;
  double* wwarp = wsm + (size_t)warp * 2 * DC * 96;

  if (USE_TMA && tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue_tile = [&](int tile) {
    const int st = tile & 1;
    const int s0 = s_begin + tile * TS;
    const int nf = min(TS, s_end - s0);
    if (USE_TMA) {
      if (tid == 0) {
        const uint32_t bytes = nf * (nvalid * 8 + nvalid + 96 + 32);
        mbar_expect_tx(&sm.bar[st], bytes);
        for (int f = 0; f < nf; ++f) {
          tma_load_1d(&sm.uv[st][f][0], uv + (size_t)(s0 + f) * N + n0, nvalid * 8, &sm.bar[st]);
          tma_load_1d(&sm.mask[st][f][0], mask + (size_t)(s0 + f) * N + n0, nvalid, &sm.bar[st]);
        }
        tma_load_1d(&sm.pose[st][0][0], poses + (size_t)s0 * 12, nf * 96, &sm.bar[st]);
        tma_load_1d(&sm.intr[st][0][0], intr + (size_t)s0 * 4, nf * 32, &sm.bar[st]);
      }
    } else {
      for (int f = 0; f < nf; ++f) {
        if (active) {
          sm.uv[st][f][tid] = uv[(size_t)(s0 + f) * N + n];
          sm.mask[st][f][tid] = mask[(size_t)(s0 + f) * N + n];
        }
      }
      for (int i = tid; i < nf * 12; i += TN) (&sm.pose[st][0][0])[i] = poses[(size_t)s0 * 12 + i];
      for (int i = tid; i < nf * 4; i += TN) (&sm.intr[st][0][0])[i] = intr[(size_t)s0 * 4 + i];
    }
  };

  double X0 = 0, X1 = 0, X2 = 0;
  bool pconst = false;
  if (active) {
    X0 = points[(size_t)n * 3 + 0];
    X1 = points[(size_t)n * 3 + 1];
    X2 = points[(size_t)n * 3 + 2];
    pconst = point_const ? (point_const[n] != 0) : false;
  }
  double hpp[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
  double ws[NS > 0 ? NS : 1][3];
#pragma unroll
  for (int j = 0; j < (NS > 0 ? NS : 1); ++j) ws[j][0] = ws[j][1] = ws[j][2] = 0;
  double cost_acc = 0, gs[2] = {0, 0}, hss[3] = {0, 0, 0};

  if (ntiles > 0) issue_tile(0);
  int fcount = 0;
  for (int tile = 0; tile < ntiles; ++tile) {
    const int st = tile & 1;
    const int s0 = s_begin + tile * TS;
    const int nf = min(TS, s_end - s0);
    // every warp is done with the other stage (tile-1) once it gets here: one CTA sync per TS frames
    if (tile > 0) __syncthreads();
    if (tile + 1 < ntiles) issue_tile(tile + 1);
    if (USE_TMA) mbar_wait(&sm.bar[st], (tile >> 1) & 1);
    else __syncthreads();

    for (int f = 0; f < nf; ++f, ++fcount) {
      const int s = s0 + f;
      double* wt = wwarp + (fcount & 1) * DC * 96;      // this warp's tile: DC rows of [32][3] doubles
      double jc0[8], jc1[8];
      double rx = 0.0, ry = 0.0;
      const bool valid = active && sm.mask[st][f][tid] != 0;
      if (valid) {
        const double* P = sm.pose[st][f];
        const double* I = sm.intr[st][f];
        const float2 ob = sm.uv[st][f][tid];
        const double R00 = P[0], R01 = P[1], R02 = P[2], t0 = P[3];
        const double R10 = P[4], R11 = P[5], R12 = P[6], t1 = P[7];
        const double R20 = P[8], R21 = P[9], R22 = P[10], t2 = P[11];
        const double fo = I[0], cx = I[1], cy = I[2];
        const double kk = (MODEL == VGG_SIMPLE_RADIAL) ? I[3] : 0.0;
        const double a1 = R00 * X0 + R01 * X1 + R02 * X2;
        const double a2 = R10 * X0 + R11 * X1 + R12 * X2;
        const double a3 = R20 * X0 + R21 * X1 + R22 * X2;
        const double px = a1 + t0, py = a2 + t1, pz = a3 + t2;
        const double iz = 1.0 / pz;
        const double u = px * iz, w_ = py * iz;
        const double r2 = u * u + w_ * w_;
        const double d = 1.0 + kk * r2;
        rx = fo * d * u + cx - (double)ob.x;
        ry = fo * d * w_ + cy - (double)ob.y;
        cost_acc += 0.5 * (rx * rx + ry * ry);
        // f*A, A = d(distorted)/d(u,v)
        double a00, a01, a11;
        if (MODEL == VGG_SIMPLE_RADIAL) {
          a00 = fo * (d + 2.0 * kk * u * u);
          a01 = fo * (2.0 * kk * u * w_);
          a11 = fo * (d + 2.0 * kk * w_ * w_);
        } else {
          a00 = fo; a01 = 0.0; a11 = fo;
        }
        // Jproj (2x3) = f*A * iz*[[1,0,-u],[0,1,-v]]
        const double j00 = a00 * iz, j01 = a01 * iz, j02 = -(a00 * u + a01 * w_) * iz;
        const double j10 = a01 * iz, j11 = a11 * iz, j12 = -(a01 * u + a11 * w_) * iz;
        // camera columns: delta(3) = Jproj * (-2[RX]x), t(3) = Jproj, f, k
        jc0[0] = 2.0 * (-a3 * j01 + a2 * j02);  jc1[0] = 2.0 * (-a3 * j11 + a2 * j12);
        jc0[1] = 2.0 * (a3 * j00 - a1 * j02);   jc1[1] = 2.0 * (a3 * j10 - a1 * j12);
        jc0[2] = 2.0 * (-a2 * j00 + a1 * j01);  jc1[2] = 2.0 * (-a2 * j10 + a1 * j11);
        jc0[3] = j00; jc0[4] = j01; jc0[5] = j02;
        jc1[3] = j10; jc1[4] = j11; jc1[5] = j12;
        jc0[6] = d * u;            jc1[6] = d * w_;
        jc0[7] = fo * u * r2;      jc1[7] = fo * w_ * r2;
        // point columns: Jproj * R
        double jx0[3], jx1[3];
        jx0[0] = j00 * R00 + j01 * R10 + j02 * R20;
        jx0[1] = j00 * R01 + j01 * R11 + j02 * R21;
        jx0[2] = j00 * R02 + j01 * R12 + j02 * R22;
        jx1[0] = j10 * R00 + j11 * R10 + j12 * R20;
        jx1[1] = j10 * R01 + j11 * R11 + j12 * R21;
        jx1[2] = j10 * R02 + j11 * R12 + j12 * R22;
        if (pconst) { jx0[0] = jx0[1] = jx0[2] = jx1[0] = jx1[1] = jx1[2] = 0.0; }
        // point blocks
        gp[0] += jx0[0] * rx + jx1[0] * ry;
        gp[1] += jx0[1] * rx + jx1[1] * ry;
        gp[2] += jx0[2] * rx + jx1[2] * ry;
        hpp[0] += jx0[0] * jx0[0] + jx1[0] * jx1[0];
        hpp[1] += jx0[0] * jx0[1] + jx1[0] * jx1[1];
        hpp[2] += jx0[0] * jx0[2] + jx1[0] * jx1[2];
        hpp[3] += jx0[1] * jx0[1] + jx1[1] * jx1[1];
        hpp[4] += jx0[1] * jx0[2] + jx1[1] * jx1[2];
        hpp[5] += jx0[2] * jx0[2] + jx1[2] * jx1[2];
        // coupling blocks: row i of the warp tile holds [32][3] doubles
#pragma unroll
        for (int i = 0; i < DC; ++i) {
          double* row = wt + i * 96 + lane * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) row[c] = jc0[i] * jx0[c] + jc1[i] * jx1[c];
        }
        if (NS > 0) {
#pragma unroll
          for (int j = 0; j < NS; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) ws[j][c] += jc0[6 + j] * jx0[c] + jc1[6 + j] * jx1[c];
            gs[j] += jc0[6 + j] * rx + jc1[6 + j] * ry;
          }
          hss[0] += jc0[6] * jc0[6] + jc1[6] * jc1[6];
          if (NS > 1) {
            hss[1] += jc0[6] * jc0[7] + jc1[6] * jc1[7];
            hss[2] += jc0[7] * jc0[7] + jc1[7] * jc1[7];
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { jc0[i] = 0.0; jc1[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < DC; ++i) {
          double* row = wt + i * 96 + lane * 3;
          row[0] = 0.0; row[1] = 0.0; row[2] = 0.0;
        }
      }
      // hand the warp tile to the TMA engine; nobody else touches it
      if (USE_TMA) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && nvalid_w > 0) {
#pragma unroll
          for (int i = 0; i < DC; ++i)
            tma_store_1d(W + ((size_t)(s * DC + i) * N + nw0) * 3, wt + i * 96, nvalid_w * 24);
          tma_store_commit();
          tma_store_wait_read<1>();     // the other buffer (frame-1) has been read out
        }
      } else {
        __syncwarp();
        for (int i = 0; i < DC; ++i)
          for (int e = lane; e < nvalid_w * 3; e += 32) W[((size_t)(s * DC + i) * N + nw0) * 3 + e] = wt[i * 96 + e];
      }
      // camera record: reduce 16 values at a time across the warp, one f64 RED per value per warp
#pragma unroll
      for (int b = 0; b < NB16; ++b) {
        double a[16];
        if (b == 0) cam_batch<DC, NS, KR, 0>(a, jc0, jc1, rx, ry, std::make_integer_sequence<int, 16>{});
        if (b == 1) cam_batch<DC, NS, KR, 16>(a, jc0, jc1, rx, ry, std::make_integer_sequence<int, 16>{});
        if (b == 2) cam_batch<DC, NS, KR, 32>(a, jc0, jc1, rx, ry, std::make_integer_sequence<int, 16>{});
        const double r = warp_reduce_scatter<16>(a, lane);
        const int k = b * 16 + lane;
        if (lane < 16 && k < KR && r != 0.0) atomicAdd(&camrec[(size_t)s * KR + k], r);
      }
      __syncwarp();      // lanes may not overwrite the other buffer before lane 0 returned from wait_group
    }
  }
  if (USE_TMA && lane == 0) tma_store_wait_all<0>();

  // flush per-point accumulators
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (gp[c] != 0.0) atomicAdd(&g_p[(size_t)n * 3 + c], gp[c]);
#pragma unroll
    for (int c = 0; c < 6; ++c)
      if (hpp[c] != 0.0) atomicAdd(&H_pp[(size_t)n * 6 + c], hpp[c]);
    if (NS > 0) {
#pragma unroll
      for (int j = 0; j < NS; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (ws[j][c] != 0.0) atomicAdd(&W[((size_t)(S * DC + j) * N + n) * 3 + c], ws[j][c]);
    }
  }
  // CTA-wide scalars: cost, g_s, H_ss
  {
    double a[8] = {cost_acc, gs[0], gs[1], hss[0], hss[1], hss[2], 0.0, 0.0};
    const double r = warp_reduce_scatter<8>(a, lane);
    if (lane < 8) red[warp * 8 + lane] = r;
    __syncthreads();
    if (tid < 6) {
      double acc = 0;
#pragma unroll
      for (int wq = 0; wq < NWARP; ++wq) acc += red[wq * 8 + tid];
      if (tid == 0) atomicAdd(cost, acc);
      else if (NS > 0 && acc != 0.0) atomicAdd(&shared_out[tid - 1], acc);
    }
  }
}

template <int MODEL, int MODE>
static int launch_blocks(const vgg_ba_problem* p, double* cost, double* camrec, double* g_p, double* H_pp,
                         double* W, double* shared_out, int frames_per_cta, cudaStream_t stream, bool outputs_zeroed) {
  using C = BlkCfg<MODEL, MODE>;
  const int S = p->S, N = p->N;
  const size_t smem = align_up(sizeof(BlkSmem), 128) + sizeof(double) * (NWARP * 2 * C::DC * 96 + NWARP * 8);
  const bool tma_ok = (N % 16 == 0) && ((reinterpret_cast<uintptr_t>(p->uv) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p->mask) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p->poses) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p->intr) & 15) == 0);
  if (frames_per_cta <= 0) {
    // small problems: one full wave of 148 SMs x 4 resident CTAs (no tail); large ones: whole frame range
    // per CTA so the per-point accumulators are flushed once.  Chunks are a multiple of TS frames.
    const int nb = (N + TN - 1) / TN;
    int chunks = (148 * 4) / nb;
    if (chunks < 1) chunks = 1;
    if (chunks > (S + TS - 1) / TS) chunks = (S + TS - 1) / TS;
    frames_per_cta = (S + chunks - 1) / chunks;
    frames_per_cta = ((frames_per_cta + TS - 1) / TS) * TS;
  }
  dim3 grid((N + TN - 1) / TN, (S + frames_per_cta - 1) / frames_per_cta);
  // zero the accumulated outputs (cost | camrec | g_p | H_pp are contiguous in the solver workspace,
  // but this entry point does not assume it)
  if (!outputs_zeroed) {
    VGG_CUDA_CHECK(cudaMemsetAsync(cost, 0, sizeof(double), stream));
    VGG_CUDA_CHECK(cudaMemsetAsync(camrec, 0, sizeof(double) * (size_t)S * C::KR, stream));
    VGG_CUDA_CHECK(cudaMemsetAsync(g_p, 0, sizeof(double) * (size_t)N * 3, stream));
    VGG_CUDA_CHECK(cudaMemsetAsync(H_pp, 0, sizeof(double) * (size_t)N * 6, stream));
    VGG_CUDA_CHECK(cudaMemsetAsync(shared_out, 0, sizeof(double) * 8, stream));
    if (C::NS > 0)
      VGG_CUDA_CHECK(cudaMemsetAsync(W + (size_t)S * C::DC * N * 3, 0, sizeof(double) * (size_t)C::NS * N * 3, stream));
  }
  if (tma_ok) {
    auto kern = ba_blocks_kernel<MODEL, MODE, true>;
    VGG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TN, smem, stream>>>(S, N, frames_per_cta, reinterpret_cast<const float2*>(p->uv), p->mask, p->poses,
                                     p->intr, p->points, p->point_const, cost, camrec, g_p, H_pp, W, shared_out);
  } else {
    auto kern = ba_blocks_kernel<MODEL, MODE, false>;
    VGG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, TN, smem, stream>>>(S, N, frames_per_cta, reinterpret_cast<const float2*>(p->uv), p->mask, p->poses,
                                     p->intr, p->points, p->point_const, cost, camrec, g_p, H_pp, W, shared_out);
  }
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}

int ba_build_blocks(const vgg_ba_problem* p, double* cost, double* camrec, double* g_p, double* H_pp, double* W,
                    double* shared_out, int frames_per_cta, cudaStream_t stream, bool outputs_zeroed) {
  const int key = p->camera_model * 3 + p->intr_mode;
  switch (key) {
    case 0: return launch_blocks<0, 0>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
    case 1: return launch_blocks<0, 1>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
    case 2: return launch_blocks<0, 2>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
    case 3: return launch_blocks<1, 0>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
    case 4: return launch_blocks<1, 1>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
    case 5: return launch_blocks<1, 2>(p, cost, camrec, g_p, H_pp, W, shared_out, frames_per_cta, stream, outputs_zeroed);
  }
  set_error("bad camera_model/intr_mode %d/%d", p->camera_model, p->intr_mode);
  return VGG_EINVAL;
}

}  // namespace vgg
