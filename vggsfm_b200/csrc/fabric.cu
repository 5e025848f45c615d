// Multi-GPU exchange kernels of the track-sharded bundle adjustment (one process per GPU), all on the
// symmetric (peer-mapped) allocation the caller hands over in vgg_ba_fabric -- no NCCL call and no host callback inside
// the LM loop:
//   fabric_barrier_kernel    cross-rank barrier: every rank stores the epoch into its slot of every peer's flag row
//                            (st.release.sys over NVLink) and polls its own row (ld.acquire.sys, bounded spin).
//   fabric_gather_kernel     second half of the reduce-scatter / all-gather of the reduced camera system: the Schur
//                            kernels RED each 128-row block of the lower triangle ONLY into its owner's copy (block b ->
//                            rank b mod world, csrc/syrk_i8.cu epilogue), so inbound traffic per GPU does not grow with
//                            the number of ranks; after the barrier every rank PULLS the blocks it does not own from
//                            their owners.  All ranks then hold bit-identical systems (no arrival-order differences).
//   fabric_allreduce_kernel  small vectors (candidate cost, model terms, camera gradient: <= Dpad+8 doubles): one CTA
//                            writes the vector into its mailbox slot on every peer, barrier, then sums the slots in
//                            rank order -- identical bits on every rank; one slot can be combined with max instead.
#include "common.cuh"

namespace vgg {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// threads 0..world-1 of the calling CTA: signal `epoch` to every peer, wait for every peer's signal
__device__ __forceinline__ void signal_and_wait(const FabricDev& fd, size_t flags_off, unsigned long long epoch, int* err,
                                                int tid) {
  if (tid < fd.world) {
    __threadfence_system();
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(fd.peer[tid] + flags_off) + fd.rank;
    st_release_sys(remote, epoch);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(fd.peer[fd.rank] + flags_off) + tid;
    unsigned long long v = ld_acquire_sys(mine);
    long spins = 0;
    while (v < epoch && ++spins < (1L << 22)) v = ld_acquire_sys(mine);      // bounded (~seconds): a lost peer must not hang the GPU
    if (v < epoch) *err = 1;
  }
}

__global__ void fabric_barrier_kernel(FabricDev fd, size_t flags_off, unsigned long long epoch, int* err) {
  signal_and_wait(fd, flags_off, epoch, err, threadIdx.x);
}

// fd.peer[r] = base of rank r's reduced-system buffer of THIS iteration.  rows [0, nrows) of `lda` doubles; matrix rows
// (r < nmat) carry their lower-triangle part [0, r], the vector rows after them ncols_vec entries.
__global__ void __launch_bounds__(256) fabric_gather_kernel(FabricDev fd, int nrows, int nmat, int ncols_vec, int lda) {
  double* local = fd.peer[fd.rank];
  for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
    const int owner = (r >> 7) % fd.world;
    if (owner == fd.rank) continue;
    const int ncols = r < nmat ? r + 1 : ncols_vec;
    const int n2 = (ncols + 1) >> 1;
    const double2* src = reinterpret_cast<const double2*>(fd.peer[owner] + (size_t)r * lda);
    double2* dst = reinterpret_cast<double2*>(local + (size_t)r * lda);
    for (int c = threadIdx.x; c < n2; c += blockDim.x) dst[c] = __ldcv(src + c);
  }
}

// mailbox of rank r at fd.peer[r] + mail_off: [2 parities][world slots][mail_len]
__global__ void __launch_bounds__(512) fabric_allreduce_kernel(FabricDev fd, size_t flags_off, size_t mail_off, int mail_len,
                                                               int parity, unsigned long long epoch, double* vec, int count,
                                                               int max_slot, int* err) {
  const int tid = threadIdx.x;
  for (int i = tid; i < count; i += blockDim.x) {
    const double v = vec[i];
    for (int r = 0; r < fd.world; ++r)
      fd.peer[r][mail_off + ((size_t)parity * fd.world + fd.rank) * mail_len + i] = v;
  }
  __syncthreads();
  signal_and_wait(fd, flags_off, epoch, err, tid);
  __syncthreads();
  const volatile double* box = fd.peer[fd.rank] + mail_off + (size_t)parity * fd.world * mail_len;
  for (int i = tid; i < count; i += blockDim.x) {
    double s = box[i];
    for (int r = 1; r < fd.world; ++r) {
      const double x = box[(size_t)r * mail_len + i];
      s = (i == max_slot) ? fmax(s, x) : s + x;
    }
    vec[i] = s;
  }
}

}  // namespace

int launch_fabric_barrier(const FabricDev& fd, size_t flags_off, unsigned long long epoch, int* err, cudaStream_t st) {
  fabric_barrier_kernel<<<1, 32, 0, st>>>(fd, flags_off, epoch, err);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}
int launch_fabric_gather(const FabricDev& fd, int nrows, int nmat, int ncols_vec, int lda, cudaStream_t st) {
  fabric_gather_kernel<<<296, 256, 0, st>>>(fd, nrows, nmat, ncols_vec, lda);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}
int launch_fabric_allreduce(const FabricDev& fd, size_t flags_off, size_t mail_off, int mail_len, int parity,
                            unsigned long long epoch, double* vec, int count, int max_slot, int* err, cudaStream_t st) {
  VGG_REQUIRE(count <= mail_len, "fabric all-reduce: vector longer than the mailbox");
  fabric_allreduce_kernel<<<1, 512, 0, st>>>(fd, flags_off, mail_off, mail_len, parity, epoch, vec, count, max_slot, err);
  VGG_LAUNCH_CHECK();
  return VGG_OK;
}

}  // namespace vgg
