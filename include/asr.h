/*
 * include/asr.h -- C ABI of libasr.so, the B200 (sm_100a) implementation of the Augmented
 * Super-Resolution hot path of nicoloalbergoni/DeepLabV3Plus-Augmented-SuperResolution.
 *
 * The reference has no FFI: its boundary is a Python call surface executed by TensorFlow ops.  Each
 * entry point below replaces the TensorFlow op sequence behind one reference function (cited as
 * file:line relative to the reference root).  The Python mirror of that surface lives in
 * deeplabv3plus_augmented_superresolution_b200/superresolution_scripts/ and calls these symbols
 * through ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; `d_` = device pointer on the current CUDA device, `h_` = host.
 *   - every function returns 0 on success or a negative ASR_E* code; asr_last_error() gives the
 *     thread-local message.  Nothing throws or aborts across the ABI.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  No call
 *     synchronises the device; small host tables are uploaded with cudaMemcpyAsync on `stream`
 *     from pageable memory, which only blocks the calling host thread until the bytes are staged.
 *   - the library keeps no device memory: callers pass workspaces sized by *_workspace_bytes; the
 *     only allocations are transient per-copy transform tables (32 bytes per copy) made and freed
 *     in stream order inside asr_warp_affine / asr_backproject_batched.
 *   - all images are fp32, C-contiguous.  There is no CPU fallback.
 *   - alignment: every device array must be 16-byte aligned and every `d_workspace` 256-byte aligned
 *     (128-bit loads/stores and TMA tensor maps); any cudaMalloc / torch allocation satisfies both.
 *     A misaligned pointer is rejected with ASR_EINVAL, never dereferenced.
 *   - thread-safety: re-entrant; one host thread per GPU may call concurrently, and one process may
 *     drive several GPUs in turn (per-device kernel attributes are tracked per device).
 */
#ifndef ASR_H
#define ASR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASR_VERSION 200

enum {
    ASR_OK = 0,
    ASR_EINVAL = -1,       /* bad shape / parameter */
    ASR_ENULL = -2,        /* null pointer */
    ASR_EUNSUPPORTED = -3, /* valid in the reference, not implemented here (message says what) */
    ASR_ECUDA = -4,        /* CUDA runtime error (message carries cudaGetErrorString) */
    ASR_EWORKSPACE = -5,   /* workspace too small */
    ASR_EDTYPE = -6        /* DLPack tensor is not fp32 / not on CUDA / not contiguous */
};

enum { ASR_OPT_ADAM = 0, ASR_OPT_SGD = 1, ASR_OPT_ADAGRAD = 2, ASR_OPT_ADADELTA = 3, ASR_OPT_ADAMAX = 4 };
enum { ASR_INTERP_NEAREST = 0, ASR_INTERP_BILINEAR = 1 };
enum { ASR_OPM_ARGMAX = 0, ASR_OPM_SLICE = 1, ASR_OPM_SLICE_MAX = 2 };
enum { ASR_BACKPROJECT_MAX = 0, ASR_BACKPROJECT_MEAN = 1 };

/* One solve's hyper-parameters: the kwargs of Superresolution.__init__
 * (superresolution_scripts/superresolution.py:26-42) and Optimizer.__init__
 * (superresolution_scripts/optimizer.py:4-48), plus the shared optimizer's step counter at entry
 * (`optimizer.iterations`, which the reference never resets between images: SR_single_class.py:66-107). */
typedef struct AsrSolveParams {
    float lambda_df, lambda_tv, lambda_l2, lambda_l1;
    int32_t num_iter;
    int32_t optimizer;                /* ASR_OPT_* */
    float learning_rate, epsilon, beta_1, beta_2;
    int32_t amsgrad;
    float initial_accumulator_value, momentum;
    int32_t nesterov;
    int32_t lr_scheduler;             /* ExponentialDecay on/off (optimizer.py:43-52) */
    float decay_steps, decay_rate;
    int64_t step_offset;
    int32_t use_btv;                  /* bilateral TV instead of TV (superresolution.py:8-23,78) */
    int32_t images_in_flight;         /* perf knob read from params[0]: images per launch group, 0 = all */
} AsrSolveParams;

int asr_version(void);
const char* asr_last_error(void);

/* Measurement hooks (bench.py): cumulative number of kernels this library has launched, and optional
 * CUDA-event timing of the two solve kernels on their launching stream.  asr_profile_read waits for
 * the recorded events, returns {forward-residual, gradient/update} total milliseconds and launch
 * counts since the previous read, and clears them. */
long long asr_kernel_launches(void);
int asr_profile_enable(int on);
int asr_profile_read(double* ms2, long long* count2);

/* Measurement hook (bench.py roofline.l2): one launch that reads `bytes` of d_buf `passes` times with L1 bypassed.
 * With a buffer that fits the L2 and passes >> 1 the caller's CUDA-event time gives the L2 read bandwidth of the
 * device the solve's L2-resident regime is compared against.  d_sink: 4 bytes.                                  */
int asr_l2_read_probe(const void* d_buf, size_t bytes, int passes, void* d_sink, void* stream);

/* ---- superresolution.py:102-137  Superresolution.augmented_superresolution ------------------
 * Solves B independent images in one call (SR_single_class.py:83-107 loops over them one by one).
 *   params      h_ array of n_params structs; n_params == 1 (shared) or B (one per image, as the
 *               hyper-parameter sweep of sweep_script.py:88-130 needs)
 *   d_copies    [B,N,h,w] low-resolution class maps y_k
 *   h_angles    [B,N] radians, h_shifts [B,N,2] (dx,dy) HR pixels (augmentation_utils.py:14-20)
 *   h_keep      NULL, or [B,N] bytes: 0 drops the copy (copy dropout, superresolution.py:47-53)
 *   d_x_out     [B,H,W] solved high-resolution maps
 *   d_loss_out  NULL or [B]: loss of the last iteration before its update (superresolution.py:137)
 * Supported: H == S*h, W == S*w for an even integer S; S == 4 (every caller in the reference) runs the tuned kernels,
 * any other even S (x2, x6, x8: Superresolution's default feature_size is 64x64 -> 512x512) the literal per-output kernels. */
int asr_solve_workspace_bytes(int B, int N, int h, int w, int H, int W, int max_iter, size_t* bytes);
int asr_solve_batched(const AsrSolveParams* params, int n_params,
                      const float* d_copies, const float* h_angles, const float* h_shifts,
                      const uint8_t* h_keep, int B, int N, int h, int w, int H, int W,
                      float* d_x_out, float* d_loss_out,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* The same solve with the reference's verbose trace (superresolution.py:129-130 prints the loss every 10 iterations):
 * d_loss_trace [B, trace_cols] receives, in column j, the loss of iteration j*loss_every evaluated on the iterate BEFORE that
 * iteration's update (the value the reference prints as "{j*loss_every+1}/{num_iter}"); columns past an image's num_iter are
 * left untouched.  Each traced iteration costs one extra reduction over the residuals and x.                              */
int asr_solve_batched_traced(const AsrSolveParams* params, int n_params,
                             const float* d_copies, const float* h_angles, const float* h_shifts,
                             const uint8_t* h_keep, int B, int N, int h, int w, int H, int W,
                             float* d_x_out, float* d_loss_out, int loss_every, float* d_loss_trace, int trace_cols,
                             void* d_workspace, size_t workspace_bytes, void* stream);

/* Hyper-parameter sweeps (sweep_script.py:88-130, check_robustness-style grids): n_points independent
 * solves, point i using LR stack h_stack_index[i] of d_copies [n_stacks,N,h,w] (angles [n_stacks,N],
 * shifts [n_stacks,N,2]) with its own params[i]; many points may share one stack without copying it.
 * d_x_out [n_points,H,W]; workspace sized by asr_solve_workspace_bytes(n_points, ...).              */
int asr_solve_sweep(const AsrSolveParams* params, int n_points, const float* d_copies, const float* h_angles,
                    const float* h_shifts, const int32_t* h_stack_index, int n_stacks, int N, int h, int w, int H, int W,
                    float* d_x_out, float* d_loss_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* One evaluation of loss_function's residual and tape.gradient (superresolution.py:44-100,126-133)
 * at a caller-supplied x; used by parity tests to check single steps.
 *   d_x [B,H,W] -> d_resid [B,N,h,w] (D T_k R_k x - y_k), d_grad [B,H,W] (data + TV + L2 + L1)   */
int asr_loss_grad_batched(const AsrSolveParams* params, int n_params,
                          const float* d_x, const float* d_copies, const float* h_angles,
                          const float* h_shifts, const uint8_t* h_keep,
                          int B, int N, int h, int w, int H, int W,
                          float* d_resid, float* d_grad, float* d_loss_out,
                          void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- superresolution.py:139-161  max_superresolution / mean_superresolution ----------------- */
int asr_backproject_batched(int mode, const float* d_copies, const float* h_angles, const float* h_shifts,
                            int B, int N, int h, int w, int H, int W, float* d_out, void* stream);
/* The same with a caller-owned workspace (the per-copy transform table) instead of a stream-ordered allocation inside
 * the call: what a per-image loop should use.                                                                      */
int asr_backproject_workspace_bytes(int B, int N, size_t* bytes);
int asr_backproject_batched_ws(int mode, const float* d_copies, const float* h_angles, const float* h_shifts,
                               int B, int N, int h, int w, int H, int W, float* d_out,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- augmentation_utils.py:11-27  create_augmented_copies (rotate then translate) ------------
 * d_image [H,W,C] -> d_out [N,H,W,C]; also check_robustness.py:44-50 (interp NEAREST for labels).
 * angles/shifts are the already-drawn values: the RNG stays in Python (global NumPy stream).      */
int asr_warp_affine(const float* d_image, const float* h_angles, const float* h_shifts, int N,
                    int H, int W, int C, int interp, float* d_out, void* stream);
/* The same with a caller-owned workspace (transform table + the image padded to 4 channels). */
int asr_warp_affine_workspace_bytes(int N, int H, int W, int C, size_t* bytes);
int asr_warp_affine_ws(const float* d_image, const float* h_angles, const float* h_shifts, int N,
                       int H, int W, int C, int interp, float* d_out,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- augmentation_utils.py:80-115 + utils.py:115-119  OPM extraction ------------------------
 * d_logits [N,h,w,K] NHWC -> d_class_out [N,h,w]; d_max_out [N,h,w] only for SLICE_MAX.
 * d_workspace: 2*N floats (SLICE mode per-copy min/max), may be NULL otherwise.                  */
int asr_opm_extract(const float* d_logits, int N, int h, int w, int K, int class_id, int mode,
                    float* d_class_out, float* d_max_out, void* d_workspace, void* stream);

/* ---- superres_utils.py:56-62,186-194  global min-max normalisation of a stack ----------------
 * in place allowed; d_workspace: 2 floats.                                                       */
int asr_minmax_normalize(const float* d_in, int64_t n, float new_min, float new_max, float* d_out,
                         void* d_workspace, void* stream);

/* ---- superres_utils.py:118-139  threshold_image ----------------------------------------------
 * B images of n pixels each.  d_th_mask NULL: x > max(x)*th_factor (max per image); else x >= mask.
 * d_out int32 {0, th_value}.  d_workspace: 2*B floats.                                            */
int asr_threshold(const float* d_x, int B, int64_t n, int32_t th_value, float th_factor,
                  const float* d_th_mask, int32_t* d_out, void* d_workspace, void* stream);

/* ---- utils.py:180-204  single_class_IOU, the counting part ---------------------------------------
 * B label images of n pixels (int32).  d_counts [B,4] = {inter, union} for class_id, then for class 0
 * (ground truth relabelled to {class_id, 0} when include_bg).  The caller divides and drops empty
 * unions (NaN) before averaging, as the reference does.                                           */
int asr_iou_counts(const int32_t* d_true, const int32_t* d_pred, int B, int64_t n, int class_id, int include_bg,
                   unsigned long long* d_counts, void* stream);

/* ---- DLPack front door (north_star: "ctypes over DLPack buffers") ----------------------------
 * Same calls taking DLTensor* (dlpack.h v0.8 layout) for the device arrays; they validate
 * device type kDLCUDA, dtype float32, compact strides, shapes, then forward to the functions above. */
struct DLTensor;
int asr_solve_batched_dlpack(const AsrSolveParams* params, int n_params, const struct DLTensor* copies,
                             const float* h_angles, const float* h_shifts, const uint8_t* h_keep,
                             struct DLTensor* x_out, struct DLTensor* loss_out /* may be NULL */,
                             struct DLTensor* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASR_H */
