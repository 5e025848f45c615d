/*
 * vggsfm_b200 -- C ABI of the B200-native geometry hot path for VGGSfM.
 *
 * The reference (facebookresearch/vggsfm @ e1d9d2e) has no FFI of its own: its seam for this path
 * is ordinary Python symbols plus the pycolmap object API.  Every entry
 * point below names the reference call site it replaces.  All pointers are DEVICE pointers owned
 * by the caller unless marked "host"; `stream` is a cudaStream_t passed as void*; no entry point
 * allocates device memory (the caller passes a workspace sized by the matching *_workspace_bytes).
 * Return value: 0 on success, negative VGG_E* on error; vgg_last_error() gives the message.
 *
 * Layouts (row-major, densely packed):
 *   observations  uv   float  [S,N,2]   pixels (or normalised coordinates where stated)
 *                 mask uint8  [S,N]     1 = observation participates
 *   cameras       poses  double [S,12]  cam_from_world R|t, 3x4 row-major
 *                 intr   double [S,4]   f, cx, cy, k   (k unused for SIMPLE_PINHOLE)
 *   points        double [N,3]
 */
#ifndef VGGSFM_B200_H
#define VGGSFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGG_OK 0
#define VGG_EINVAL (-1)
#define VGG_ECUDA (-2)
#define VGG_EWORKSPACE (-3)
#define VGG_ESOLVER (-4)

#define VGG_SIMPLE_PINHOLE 0
#define VGG_SIMPLE_RADIAL 1

#define VGG_INTR_CONST 0      /* intrinsics held constant (video_runner.py:813-815)            */
#define VGG_INTR_PER_FRAME 1  /* one pycolmap.Camera per frame (tensor_to_pycolmap.py:78-110)  */
#define VGG_INTR_SHARED 2     /* shared_camera=True: one Camera for all frames                 */

const char* vgg_last_error(void);
int vgg_version(void);

/* ------------------------------------------------------------------------------------------- */
/* Bundle adjustment: replaces pycolmap.bundle_adjustment(reconstruction, ba_options)           */
/* (vggsfm/utils/triangulation.py:213,1050,1142; vggsfm/runners/video_runner.py:508,1321-1331)  */
/* and the tensor<->Reconstruction marshalling around it (tensor_to_pycolmap.py:16-214).        */
/* ------------------------------------------------------------------------------------------- */

typedef struct vgg_ba_problem {
  int32_t S, N;
  int32_t camera_model;        /* VGG_SIMPLE_* */
  int32_t intr_mode;           /* VGG_INTR_*   */
  const float* uv;             /* [S,N,2] pixels */
  const uint8_t* mask;         /* [S,N] */
  const uint8_t* param_const;  /* [S*dc+ns] 1 = reduced parameter held constant (gauge, fixed poses) */
  const uint8_t* point_const;  /* [N] 1 = point held constant (video_runner.py:821-829), or NULL */
  double* poses;               /* [S,12] in/out */
  double* intr;                /* [S,4]  in/out */
  double* points;              /* [N,3]  in/out */
} vgg_ba_problem;

/* Ceres solver options as COLMAP's BundleAdjustmentOptions sets them (triangulation_helpers.py:626-635). */
typedef struct vgg_ba_options {
  int32_t max_num_iterations;
  int32_t max_num_consecutive_invalid_steps;
  int32_t jacobi_scaling;
  int32_t reserved;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
} vgg_ba_options;

typedef struct vgg_ba_summary {
  int32_t iterations, successful, termination, reserved;
  double initial_cost, final_cost, final_radius;
  double device_ms;            /* CUDA-event time of the LM loop on `stream` */
  int64_t kernel_launches;     /* kernels of this library launched by the call */
} vgg_ba_summary;

/* termination codes */
#define VGG_BA_NO_CONVERGENCE 0
#define VGG_BA_CONVERGENCE_GRADIENT 1
#define VGG_BA_CONVERGENCE_FUNCTION 2
#define VGG_BA_CONVERGENCE_PARAMETER 3
#define VGG_BA_MIN_TRUST_REGION 4
#define VGG_BA_FAILURE 5

/* Sum/max all-reduce hook over track shards (one process per GPU).  `buf` is a device pointer
 * inside the caller's workspace; op 0 = sum, 1 = max.  NULL = single GPU. */
typedef int (*vgg_allreduce_fn)(void* user, double* buf, size_t count, int op, void* stream);

void vgg_ba_default_options(vgg_ba_options* opt);
int vgg_ba_dims(int camera_model, int intr_mode, int* dc, int* ns);
int vgg_ba_workspace_bytes(int S, int N, int camera_model, int intr_mode, size_t* bytes);

/* Fused residual + analytic 2x(dc+3) Jacobian + normal-equation block kernel (one launch).
 * Outputs (all double): cost[1]; camrec[S,KR] = per frame (g_c[dc] | H_cc upper-packed | H_cs[6,ns]);
 * g_p[N,3]; H_pp[N,6] (xx,xy,xz,yy,yz,zz); W[(S*dc+ns),N,3] coupling blocks J_c^T J_p (shared
 * intrinsics rows last); shared[8] = (g_s[2], H_ss xx,xy,yy).  KR = vgg_ba_camrec_len(). */
int vgg_ba_camrec_len(int camera_model, int intr_mode);
int vgg_ba_build_blocks(const vgg_ba_problem* prob, double* cost, double* camrec, double* g_p,
                        double* H_pp, double* W, double* shared, int frames_per_cta, void* stream);

/* Schur complement of the point blocks onto the camera system (second kernel of the path):
 * given the blocks above, the Jacobi scales and the trust-region radius, writes
 * Sraw[D,Dpad] (lower triangle valid) = H_cc - sum_j W_j V_j^-1 W_j^T and rhs[Dpad] =
 * -(g_c - sum_j W_j V_j^-1 g_pj).  Exposed for the parity tests and profiling. */
int vgg_ba_schur(const vgg_ba_problem* prob, const double* camrec, const double* g_p,
                 const double* H_pp, const double* W, const double* shared, const double* scale_p,
                 double radius, double min_diag, double max_diag, void* workspace, size_t ws_bytes,
                 double* Sraw, double* rhs, int* Dpad_out, void* stream);

/* Whole Levenberg-Marquardt solve (Ceres trust-region semantics).  `trace` is a HOST array
 * [max_num_iterations, 8] (it, cost, candidate_cost, model_change, rho, radius, step_norm, flags) or NULL. */
int vgg_ba_solve(const vgg_ba_problem* prob, const vgg_ba_options* opt, void* workspace,
                 size_t ws_bytes, vgg_allreduce_fn allreduce, void* allreduce_user,
                 vgg_ba_summary* summary, double* trace, void* stream);

#ifdef __cplusplus
}
#endif
#endif
