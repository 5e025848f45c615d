// Development probes (NOT part of the public C ABI in include/vggsfm_b200.h): exported from the library for
// tools/syrk_i8_check.py only.
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
/* Tensor-pipe probe used for the SYRK's roofline: cycles per back-to-back tcgen05.mma.kind::i8 (M=128, K=32); mode
 * bit0 selects N=256 (else 128), mode>>1 the shared-memory layout (0 = 64 B swizzle, 1 = 128 B swizzle, 2 = none).
 * out_cycles is a host pointer. */
int vgg_syrk_ozaki_mma_rate(int iters, int mode, double* out_cycles, void* stream);
/* Cluster hardware-rule probe used while developing the CTA-pair SYRK (bounded, cannot hang): out_host[0..2] int. */
int vgg_probe_remote_mbarrier(int* out_host, void* stream);

#ifdef __cplusplus
}
#endif
