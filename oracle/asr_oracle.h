/*
 * oracle/asr_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the Augmented Super-Resolution hot path of
 * nicoloalbergoni/DeepLabV3Plus-Augmented-SuperResolution.  Every function
 * here follows the reference call site it cites, op by op, materialising the
 * same intermediates TensorFlow would.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in tensorflow==2.7.0 and
 * tensorflow-addons==0.15.0 (reference configs/requirements.txt:114-115),
 * neither of which is installed or vendored; the reference has no tests or
 * golden vectors for the path.  The operator semantics restated here are the
 * published behaviour of those releases.
 */
#ifndef ASR_ORACLE_H
#define ASR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_OPT_ADAM = 0, ORC_OPT_SGD = 1, ORC_OPT_ADAGRAD = 2, ORC_OPT_ADADELTA = 3, ORC_OPT_ADAMAX = 4 };
enum { ORC_INTERP_NEAREST = 0, ORC_INTERP_BILINEAR = 1 };
enum { ORC_OPM_ARGMAX = 0, ORC_OPM_SLICE = 1, ORC_OPM_SLICE_MAX = 2 };

/* Mirrors the kwargs of Superresolution.__init__ (superresolution.py:26-42)
 * and Optimizer.__init__ (optimizer.py:4-48). */
typedef struct {
    float lambda_df, lambda_tv, lambda_l2, lambda_l1;
    int32_t num_iter;
    int32_t optimizer;                /* ORC_OPT_* */
    float learning_rate, epsilon, beta_1, beta_2;
    int32_t amsgrad;
    float initial_accumulator_value, momentum;
    int32_t nesterov;
    int32_t lr_scheduler;
    float decay_steps, decay_rate;
    int64_t step_offset;              /* optimizer.iterations at entry */
    int32_t use_btv;                  /* superresolution.py:78 */
    int32_t n_keep;                   /* copy dropout: number of copies kept, 0 = all; keep mask passed separately */
} orc_params;

int orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */

/* tfa.image.angles_to_projective_transforms / translations_to_projective_transforms */
void orc_rotate_matrix(float angle, int H, int W, float t[8]);
void orc_translate_matrix(float dx, float dy, float t[8]);
/* gradient of ImageProjectiveTransformV3: 3x3 inverse, renormalised */
void orc_invert_transform(const float t[8], float tinv[8]);

/* ImageProjectiveTransformV3, fill_mode CONSTANT.  in [N,H,W,C] -> out [N,H,W,C].
 * n_tr == 1 broadcasts one transform over the batch. */
void orc_projective_transform(const float* in, int N, int H, int W, int C,
                              const float* transforms, int n_tr, int interp,
                              float fill_value, float* out);

/* tf.image.resize(method=bilinear, antialias=False) == ResizeBilinear(half_pixel_centers=True) */
void orc_resize_bilinear(const float* in, int N, int h, int w, int C, float* out, int H, int W);
/* ResizeBilinearGrad: grad [N,H,W,C] (output side) -> in_grad [N,h,w,C] (input side) */
void orc_resize_bilinear_grad(const float* grad, int N, int H, int W, int C, float* in_grad, int h, int w);

/* superresolution.py:44-100 forward + tape.gradient (:126-133).
 * x [H,W], copies [N,h,w], returns loss, writes grad [H,W].
 * keep may be NULL (no dropout) or N bytes (1 = keep). */
float orc_loss_and_grad(const float* x, int H, int W, const float* copies, int N, int h, int w,
                        const float* angles, const float* shifts, const orc_params* p,
                        const uint8_t* keep, float* grad, float* resid_out /* [N,h,w] or NULL */);

/* superresolution.py:102-137.  x_out [H,W]; returns last-iteration loss.
 * trace (optional, may be NULL): x after iterations listed in trace_iters[n_trace] -> trace[n_trace,H,W] */
float orc_augmented_superresolution(const float* copies, int N, int h, int w, int H, int W,
                                    const float* angles, const float* shifts, const orc_params* p,
                                    const uint8_t* keep, float* x_out,
                                    const int32_t* trace_iters, int n_trace, float* trace);

/* superresolution.py:139-161: mode 0 = max, 1 = mean */
void orc_backproject(const float* copies, int N, int h, int w, int H, int W,
                     const float* angles, const float* shifts, int mode, float* out);

/* superres_utils.py:118-139.  th_mask may be NULL -> th_factor path. out int32 [n] */
void orc_threshold(const float* x, int64_t n, int32_t th_value, float th_factor, const float* th_mask, int32_t* out);
/* superres_utils.py:56-62 with global min/max over the whole buffer (load_SR_data :186-194) */
void orc_minmax_normalize_global(const float* in, int64_t n, float new_min, float new_max, float* out);
/* augmentation_utils.py:80-115 + utils.py:115-119. logits [N,h,w,K] -> class_out [N,h,w], max_out [N,h,w] (slice_max only) */
void orc_opm_extract(const float* logits, int N, int h, int w, int K, int class_id, int mode,
                     float* class_out, float* max_out);
/* utils.py:180-204 single_class_IOU on int32 label images; returns NaN-filtered mean */
double orc_single_class_iou(const int32_t* y_true, const int32_t* y_pred, int64_t n, int class_id, int include_bg);

#ifdef __cplusplus
}
#endif
#endif
