// Coarse-tracker correlation on the 5th-generation tensor cores (tcgen05.mma kind::f16, TMEM accumulators).
//
// CorrBlock.corr + CorrBlock.sample of the coarse tracker (vggsfm/models/track_modules/blocks.py:363-416: per level one
// torch.matmul of the [N,C] targets with all H*W positions, fp16 under autocast, then (2r+1)^2 bilinear taps) for
// C = 128 channels.  r02 measurement at BASELINE's C4 shape (128 frames x 1024 queries, 128x128 maps, 5 levels, r = 4):
// the CUDA-core footprint kernel (csrc/corr.cu) needs 11.9 ms per refinement iteration -- 16.8 GB of L2 gathers, 0.2 of
// the HBM roofline -- and loses to the reference's own GPU path (dense fp16 GEMM + grid_sample: 4.7 ms).  The dense
// product is the right shape for this chip after all, as long as the 5.7 GB volume is never written: here one CTA per
// SM walks work items (frame, 128 queries); per level and per tile of 256 positions it issues
// tcgen05.mma M=128 x N=256 x K=128 (8 instructions of K=16) from shared memory into one of two 256-column TMEM
// accumulators, while eight epilogue warps (thread = query = TMEM lane, two warps per lane quarter) drain the other one with tcgen05.ld and keep
// only the (2r+2)^2 footprint values each query needs, in a shared-memory footprint table; after a level's last tile the
// same warps interpolate the (2r+1)^2 taps and write them.  The correlation volume lives 2 us in TMEM.
//   operands: K-major, 64-byte swizzle (k-blocks of 32 fp16 channels), stored in global memory as ready-made tile
//   images -- B (feature positions) once per CorrBlock, A (targets) once per call -- so the producer warp moves a whole
//   256 x 128 operand tile with ONE 64 KB bulk copy (cp.async.bulk -> UBLKCP), no tensor map.
// grid_sample semantics kept: align_corners=True, padding "zeros" (taps outside the map read 0), tap order
// out[a*(2r+1)+b] at x = cx + (a-r), y = cy + (b-r).  Maps whose width is a power of two (128 -> 8 at C4).
#include <cuda_fp16.h>
#include <algorithm>
#include "common.cuh"

namespace vgg {

namespace {

constexpr int CT_M = 128;                 // queries per work item (TMEM lanes)
constexpr int CT_N = 256;                 // positions per tile (TMEM columns)
constexpr int CT_C = 128;                 // channels
constexpr int CT_KB = 4;                  // k-blocks of 32 channels (64 bytes)
constexpr int CT_A_BYTES = CT_M * CT_C * 2;          // 32 KB
constexpr int CT_B_BYTES = CT_N * CT_C * 2;          // 64 KB
constexpr int CT_STAGES = 2;
constexpr int CT_THREADS = 320;           // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (TMEM quarter = warp & 3, two per quarter)
constexpr int CT_FB_LD = 101;             // footprint table row stride (floats): 10*10 (+1: conflict-free per-query rows)
constexpr size_t CT_SMEM = 1024 + CT_A_BYTES + (size_t)CT_STAGES * CT_B_BYTES + (size_t)CT_M * CT_FB_LD * 4 + 256;

struct CtLevels {
  int H[8], W[8], logW[8], ntiles[8];
  size_t tile_off[8];                     // byte offset of level l's tile images inside one image's block
  size_t img_stride;                      // bytes of tile images per image (all levels)
};

__device__ __forceinline__ void tcf_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcf_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  // K-major, 64-byte swizzle, 8-row atoms 512 B apart, sm_100 descriptor version (as csrc/syrk_i8.cu)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// D = f32, A = B = f16, both K-major, N = 256, M = 128
constexpr uint32_t CT_IDESC = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(CT_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of (row r, channel c) inside a k-block tile image of `rows` x 64 B (64-byte swizzle, Swizzle<2,4,3>)
__device__ __forceinline__ int sw64_off(int r, int c_in_kb /*0..31*/) {
  const int byte = c_in_kb * 2;
  return r * 64 + ((((byte >> 4) ^ ((r >> 1) & 3))) << 4) + (byte & 15);
}

// ---- B operand: channels-last half pyramid level -> tile images [img][ntile][kb][256 x 64 B] -----------------------
__global__ void ct_build_b_kernel(int BS, int HW, int ntiles, const __half* __restrict__ lvl /*[BS,HW,128]*/,
                                  uint8_t* __restrict__ tiles, size_t tile_off, size_t img_stride) {
  // one thread per (img, position, 16-byte chunk of 8 channels): 16 chunks per position
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)BS * ntiles * CT_N * 16;
  if (i >= total) return;
  const int chunk = (int)(i & 15);
  const size_t pp = i >> 4;
  const int r = (int)(pp % CT_N);
  const size_t tt = pp / CT_N;
  const int t = (int)(tt % ntiles);
  const size_t img = tt / ntiles;
  const int p = t * CT_N + r;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (p < HW) v = *reinterpret_cast<const uint4*>(lvl + ((size_t)img * HW + p) * CT_C + chunk * 8);
  const int kb = chunk >> 2, cin = (chunk & 3) * 8;
  uint8_t* dst = tiles + img * img_stride + tile_off + ((size_t)t * CT_KB + kb) * (CT_N * 64) + sw64_off(r, cin);
  *reinterpret_cast<uint4*>(dst) = v;
}

// ---- A operand: float targets [BS,N,128] -> fp16 tile images [img][mtile][kb][128 x 64 B] ---------------------------
__global__ void ct_build_a_kernel(int BS, int N, int mtiles, const float* __restrict__ targets, uint8_t* __restrict__ a_tiles) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)BS * mtiles * CT_M * 16;
  if (i >= total) return;
  const int chunk = (int)(i & 15);
  const size_t pp = i >> 4;
  const int r = (int)(pp % CT_M);
  const size_t tt = pp / CT_M;
  const int m = (int)(tt % mtiles);
  const size_t img = tt / mtiles;
  const int n = m * CT_M + r;
  __half h[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) h[k] = __float2half(0.f);
  if (n < N) {
    const float4 a = *reinterpret_cast<const float4*>(targets + ((size_t)img * N + n) * CT_C + chunk * 8);
    const float4 b = *reinterpret_cast<const float4*>(targets + ((size_t)img * N + n) * CT_C + chunk * 8 + 4);
    h[0] = __float2half(a.x); h[1] = __float2half(a.y); h[2] = __float2half(a.z); h[3] = __float2half(a.w);
    h[4] = __float2half(b.x); h[5] = __float2half(b.y); h[6] = __float2half(b.z); h[7] = __float2half(b.w);
  }
  const int kb = chunk >> 2, cin = (chunk & 3) * 8;
  uint8_t* dst = a_tiles + ((img * mtiles + m) * CT_KB + kb) * (size_t)(CT_M * 64) + sw64_off(r, cin);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

template <int R>
__global__ void __launch_bounds__(CT_THREADS, 1)
    corr_tc_kernel(int BS, int N, int L, int mtiles, const __grid_constant__ CtLevels lv, const uint8_t* __restrict__ b_tiles,
                   const uint8_t* __restrict__ a_tiles, const float* __restrict__ coords, float* __restrict__ out) {
  constexpr int FP = 2 * R + 2, K = 2 * R + 1;
  extern __shared__ __align__(1024) uint8_t ct_smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(ct_smem), 1024));
  uint8_t* a_sm = base;
  uint8_t* b_sm = base + CT_A_BYTES;
  float* fb = reinterpret_cast<float*>(b_sm + (size_t)CT_STAGES * CT_B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(fb + CT_M * CT_FB_LD);
  uint64_t* a_full = bars;            // 1
  uint64_t* a_empty = bars + 1;       // 1
  uint64_t* b_full = bars + 2;        // [2]
  uint64_t* b_empty = bars + 4;       // [2]
  uint64_t* t_full = bars + 6;        // [2]
  uint64_t* t_empty = bars + 8;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int s = 0; s < CT_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcf_before();
  __syncthreads();
  tcf_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int nitems = BS * mtiles;
  int tiles_per_item = 0;
  for (int l = 0; l < L; ++l) tiles_per_item += lv.ntiles[l];

  if (warp == 0) {
    // ===== producer: one bulk copy per operand tile =====
    if (lane == 0) {
      int stage = 0, phase = 0, item_no = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++item_no) {
        const int img = item / mtiles;
        mbar_wait(a_empty, (uint32_t)((item_no & 1) ^ 1));
        mbar_expect_tx(a_full, CT_A_BYTES);
        tma_load_1d(a_sm, a_tiles + (size_t)item * CT_A_BYTES, CT_A_BYTES, a_full);
        for (int l = 0; l < L; ++l) {
          const uint8_t* src = b_tiles + (size_t)img * lv.img_stride + lv.tile_off[l];
          for (int t = 0; t < lv.ntiles[l]; ++t) {
            mbar_wait(&b_empty[stage], (uint32_t)(phase ^ 1));
            mbar_expect_tx(&b_full[stage], CT_B_BYTES);
            tma_load_1d(b_sm + (size_t)stage * CT_B_BYTES, src + (size_t)t * CT_B_BYTES, CT_B_BYTES, &b_full[stage]);
            if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0, phase = 0, item_no = 0;
      uint32_t tile_no = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++item_no) {
        mbar_wait(a_full, (uint32_t)(item_no & 1));
        tcf_after();
        const uint32_t a_addr = smem_u32(a_sm);
        for (int t = 0; t < tiles_per_item; ++t, ++tile_no) {
          const uint32_t buf = tile_no & 1u;
          mbar_wait(&t_empty[buf], (uint32_t)(((tile_no >> 1) & 1u) ^ 1u));
          mbar_wait(&b_full[stage], (uint32_t)phase);
          tcf_after();
          const uint32_t b_addr = smem_u32(b_sm + (size_t)stage * CT_B_BYTES);
#pragma unroll
          for (int kb = 0; kb < CT_KB; ++kb)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_f16(tmem_base + buf * CT_N, desc_sw64(a_addr + kb * (CT_M * 64) + ks * 32),
                       desc_sw64(b_addr + kb * (CT_N * 64) + ks * 32), (kb | ks) ? 1u : 0u);
          umma_commit_to(&b_empty[stage]);       // the stage may be refilled once these MMAs have read it
          umma_commit_to(&t_full[buf]);          // ... and the accumulator is complete
          if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_to(a_empty);                 // all MMAs of the item have consumed A
      }
    }
  } else {
    // ===== epilogue: thread = query row = TMEM lane; two warps per lane quarter (quarter = warp & 3), warp `sub` of a
    //       pair drains the 32-column chunks with (chunk & 1) == sub; the pair meets on a named barrier per level =====
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;                       // 0..127
    float* myfb = fb + row * CT_FB_LD;
    const float inv_sqrt_c = rsqrtf((float)CT_C);
    uint32_t tile_no = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int img = item / mtiles, m = item % mtiles;
      const int n = m * CT_M + row;
      float cx0 = 0.f, cy0 = 0.f;
      if (n < N) {
        cx0 = coords[((size_t)img * N + n) * 2];
        cy0 = coords[((size_t)img * N + n) * 2 + 1];
      }
      for (int l = 0; l < L; ++l) {
        const int W = lv.W[l], logW = lv.logW[l];
        const float scale = 1.0f / (float)(1 << l);
        const float cx = cx0 * scale, cy = cy0 * scale;
        const float fxf = floorf(cx), fyf = floorf(cy);
        const int x0 = (int)fxf - R, y0 = (int)fyf - R;        // footprint origin
        for (int i = sub; i < FP * FP; i += 2) myfb[i] = 0.f;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");       // table zeroed by both warps of the pair
        for (int t = 0; t < lv.ntiles[l]; ++t, ++tile_no) {
          const uint32_t buf = tile_no & 1u;
          mbar_wait(&t_full[buf], (uint32_t)((tile_no >> 1) & 1u));
          tcf_after();
#pragma unroll 1
          for (int ch = sub; ch < CT_N / 32; ch += 2) {
            const int p0 = t * CT_N + ch * 32;
            bool need;
            int dy = 0, xc = 0;
            if (W >= 32) {
              dy = (p0 >> logW) - y0;
              xc = p0 & (W - 1);
              need = (unsigned)dy < (unsigned)FP && xc + 31 >= x0 && xc < x0 + FP;
            } else {
              const int ya = p0 >> logW, yb = (p0 + 31) >> logW;
              need = yb >= y0 && ya < y0 + FP;
            }
            if (!__any_sync(0xffffffffu, need)) continue;
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * CT_N + ch * 32, v);
            if (need) {
              if (W >= 32) {
                float* dstrow = myfb + dy * FP - x0 + xc;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int dx = xc + j - x0;
                  if ((unsigned)dx < (unsigned)FP) dstrow[j] = __uint_as_float(v[j]) * inv_sqrt_c;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int p = p0 + j;
                  const int ddy = (p >> logW) - y0, dx = (p & (W - 1)) - x0;
                  if ((unsigned)ddy < (unsigned)FP && (unsigned)dx < (unsigned)FP)
                    myfb[ddy * FP + dx] = __uint_as_float(v[j]) * inv_sqrt_c;
                }
              }
            }
          }
          tcf_before();
          __syncwarp();
          if (lane == 0) mbar_arrive1(&t_empty[buf]);
        }
        // ---- interpolation of the K*K taps: the pair shares its 32 queries (even / odd), one query at a time so that
        //      the stores of a query are contiguous
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");       // both halves of every footprint are in
        for (int qq = sub; qq < 32; qq += 2) {
          const int nq = m * CT_M + quarter * 32 + qq;
          if (nq >= N) break;
          const float qcx = __shfl_sync(0xffffffffu, cx, qq), qcy = __shfl_sync(0xffffffffu, cy, qq);
          const float wx = qcx - floorf(qcx), wy = qcy - floorf(qcy);
          const float* f = fb + (quarter * 32 + qq) * CT_FB_LD;
          float* orow = out + ((size_t)img * N + nq) * (size_t)(L * K * K) + (size_t)l * K * K;
          for (int o = lane; o < K * K; o += 32) {
            const int a = o / K, b = o % K;
            const float d00 = f[b * FP + a], d01 = f[b * FP + a + 1], d10 = f[(b + 1) * FP + a], d11 = f[(b + 1) * FP + a + 1];
            orow[o] = d00 * (1.f - wx) * (1.f - wy) + d01 * wx * (1.f - wy) + d10 * (1.f - wx) * wy + d11 * wx * wy;
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");       // tables are re-zeroed next level
      }
    }
  }
  tcf_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int ct_levels(int H, int W, int L, CtLevels* lv) {
  size_t off = 0;
  int h = H, w = W;
  for (int l = 0; l < L; ++l) {
    if (h <= 0 || w <= 0 || (w & (w - 1)) != 0) return -1;
    lv->H[l] = h;
    lv->W[l] = w;
    int lg = 0;
    while ((1 << lg) < w) ++lg;
    lv->logW[l] = lg;
    lv->ntiles[l] = (h * w + CT_N - 1) / CT_N;
    lv->tile_off[l] = off;
    off += (size_t)lv->ntiles[l] * CT_B_BYTES;
    h /= 2;
    w /= 2;
  }
  lv->img_stride = off;
  return 0;
}

}  // namespace

}  // namespace vgg

using namespace vgg;

extern "C" {

int vgg_corr_tc_supported(int C, int H, int W, int num_levels, int radius) {
  CtLevels lv;
  return C == CT_C && num_levels >= 1 && num_levels <= 8 && (radius == 3 || radius == 4) && ct_levels(H, W, num_levels, &lv) == 0;
}

int vgg_corr_tc_bytes(int BS, int C, int H, int W, int num_levels, int N, size_t* tile_bytes, size_t* target_bytes) {
  VGG_REQUIRE(BS > 0 && N >= 0, "bad sizes");
  CtLevels lv;
  VGG_REQUIRE(C == CT_C && num_levels >= 1 && num_levels <= 8 && ct_levels(H, W, num_levels, &lv) == 0,
              "tensor-core correlation: C must be 128 and the map width a power of two");
  if (tile_bytes) *tile_bytes = (size_t)BS * lv.img_stride;
  if (target_bytes) *target_bytes = (size_t)BS * ((N + CT_M - 1) / CT_M) * CT_A_BYTES;
  return VGG_OK;
}

int vgg_corr_tc_build(int BS, int C, int H, int W, int num_levels, const void* pyramid_half, void* tiles, void* stream) {
  VGG_REQUIRE(pyramid_half && tiles, "null pointer");
  CtLevels lv;
  VGG_REQUIRE(C == CT_C && ct_levels(H, W, num_levels, &lv) == 0, "tensor-core correlation: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  g_launch_count = 0;
  const char* p = reinterpret_cast<const char*>(pyramid_half);
  for (int l = 0; l < num_levels; ++l) {
    const int HW = lv.H[l] * lv.W[l];
    const size_t total = (size_t)BS * lv.ntiles[l] * CT_N * 16;
    ct_build_b_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(BS, HW, lv.ntiles[l], reinterpret_cast<const __half*>(p),
                                                                      reinterpret_cast<uint8_t*>(tiles), lv.tile_off[l],
                                                                      lv.img_stride);
    VGG_LAUNCH_CHECK();
    p += align_up((size_t)BS * HW * C * 2, 256);      // level stride of vgg_corr_build_pyramid (elem_size 2)
  }
  return VGG_OK;
}

int vgg_corr_tc_sample(int BS, int N, int C, int H, int W, int num_levels, int radius, const void* tiles, const float* targets,
                       const float* coords, void* target_tiles, float* out, void* stream) {
  VGG_REQUIRE(tiles && targets && coords && target_tiles && out, "null pointer");
  VGG_REQUIRE(radius == 3 || radius == 4, "tensor-core correlation: radius 3 or 4");
  CtLevels lv;
  VGG_REQUIRE(C == CT_C && ct_levels(H, W, num_levels, &lv) == 0, "tensor-core correlation: unsupported shape");
  cudaStream_t st = (cudaStream_t)stream;
  g_launch_count = 0;
  if (BS == 0 || N == 0) return VGG_OK;
  const int mtiles = (N + CT_M - 1) / CT_M;
  {
    const size_t total = (size_t)BS * mtiles * CT_M * 16;
    ct_build_a_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(BS, N, mtiles, targets, reinterpret_cast<uint8_t*>(target_tiles));
    VGG_LAUNCH_CHECK();
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    VGG_CUDA_CHECK(cudaGetDevice(&dev));
    VGG_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VGG_CUDA_CHECK(cudaFuncSetAttribute(corr_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM));
    VGG_CUDA_CHECK(cudaFuncSetAttribute(corr_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM));
  }
  const int grid = std::min(sms, BS * mtiles);
  if (radius == 4)
    corr_tc_kernel<4><<<grid, CT_THREADS, CT_SMEM, st>>>(BS, N, num_levels, mtiles, lv, reinterpret_cast<const uint8_t*>(tiles),
                                                        reinterpret_cast<const uint8_t*>(target_tiles), coords, out);
  else
    corr_tc_kernel<3><<<grid, CT_THREADS, CT_SMEM, st>>>(BS, N, num_levels, mtiles, lv, reinterpret_cast<const uint8_t*>(tiles),
                                                        reinterpret_cast<const uint8_t*>(target_tiles), coords, out);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}

}  // extern "C"
