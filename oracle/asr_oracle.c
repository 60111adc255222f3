/*
 * oracle/asr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see asr_oracle.h).
 *
 * Literal, op-by-op CPU restatement of the reference hot path.  It deliberately
 * materialises every intermediate TensorFlow would ([N,H,W] tensors) and keeps
 * TensorFlow's fp32 evaluation order; build with -ffp-contract=off so that no
 * multiply-add is fused.  Speed comes only from OpenMP over independent rows.
 *
 * Reference call sites (paths relative to the reference repo root):
 *   superresolution_scripts/superresolution.py:44-100   loss_function
 *   superresolution_scripts/superresolution.py:102-137  augmented_superresolution
 *   superresolution_scripts/superresolution.py:139-161  max/mean_superresolution
 *   superresolution_scripts/superresolution.py:8-23     bilateral_tv
 *   superresolution_scripts/optimizer.py:4-52           Optimizer
 *   superresolution_scripts/superres_utils.py:56-62,118-139
 *   superresolution_scripts/augmentation_utils.py:80-115, utils.py:115-119,180-204
 * Third-party operator semantics (tensorflow 2.7.0 / tensorflow-addons 0.15.0,
 * not vendored): SURVEY.md Appendix A.1-A.8.
 */
#include "asr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- tfa.image transform matrices -------------------------------------------- */

void orc_rotate_matrix(float angle, int H, int W, float t[8]) {
    /* tfa angles_to_projective_transforms: every op in fp32 */
    float c = cosf(angle), s = sinf(angle);
    float wm = (float)W - 1.0f, hm = (float)H - 1.0f;
    float xoff = (wm - (c * wm - s * hm)) / 2.0f;
    float yoff = (hm - (s * wm + c * hm)) / 2.0f;
    t[0] = c; t[1] = -s; t[2] = xoff;
    t[3] = s; t[4] = c;  t[5] = yoff;
    t[6] = 0.0f; t[7] = 0.0f;
}

void orc_translate_matrix(float dx, float dy, float t[8]) {
    /* tfa translations_to_projective_transforms: [1,0,-dx,0,1,-dy,0,0] */
    t[0] = 1.0f; t[1] = 0.0f; t[2] = -dx;
    t[3] = 0.0f; t[4] = 1.0f; t[5] = -dy;
    t[6] = 0.0f; t[7] = 0.0f;
}

void orc_invert_transform(const float t[8], float tinv[8]) {
    /* Gradient of ImageProjectiveTransformV3: flat -> 3x3, tf.linalg.inv (fp32 LU with
     * partial pivoting), divide by element [2][2], keep the first 8. */
    float a[3][3] = {{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, {t[6], t[7], 1.0f}};
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        float best = fabsf(a[k][k]);
        for (int r = k + 1; r < 3; ++r)
            if (fabsf(a[r][k]) > best) { best = fabsf(a[r][k]); piv = r; }
        if (piv != k) {
            for (int c = 0; c < 3; ++c) { float tmp = a[k][c]; a[k][c] = a[piv][c]; a[piv][c] = tmp; }
            int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
        }
        for (int r = k + 1; r < 3; ++r) {
            a[r][k] = a[r][k] / a[k][k];
            for (int c = k + 1; c < 3; ++c) a[r][c] = a[r][c] - a[r][k] * a[k][c];
        }
    }
    float inv[3][3];
    for (int col = 0; col < 3; ++col) {
        float b[3];
        for (int r = 0; r < 3; ++r) b[r] = (perm[r] == col) ? 1.0f : 0.0f;
        /* forward substitution, unit lower */
        for (int r = 1; r < 3; ++r)
            for (int c = 0; c < r; ++c) b[r] = b[r] - a[r][c] * b[c];
        /* back substitution */
        for (int r = 2; r >= 0; --r) {
            for (int c = r + 1; c < 3; ++c) b[r] = b[r] - a[r][c] * b[c];
            b[r] = b[r] / a[r][r];
        }
        for (int r = 0; r < 3; ++r) inv[r][col] = b[r];
    }
    float d = inv[2][2];
    tinv[0] = inv[0][0] / d; tinv[1] = inv[0][1] / d; tinv[2] = inv[0][2] / d;
    tinv[3] = inv[1][0] / d; tinv[4] = inv[1][1] / d; tinv[5] = inv[1][2] / d;
    tinv[6] = inv[2][0] / d; tinv[7] = inv[2][1] / d;
}

/* ---- ImageProjectiveTransformV3 ---------------------------------------------- */

static inline float read_fill(const float* img, int H, int W, int C, long y, long x, int c, float fill) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? img[((size_t)y * W + x) * C + c] : fill;
}

void orc_projective_transform(const float* in, int N, int H, int W, int C,
                              const float* transforms, int n_tr, int interp,
                              float fill_value, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < N; ++b) {
        for (int oy = 0; oy < H; ++oy) {
            const float* t = (n_tr == 1) ? transforms : transforms + 8 * (size_t)b;
            const float* img = in + (size_t)b * H * W * C;
            float* orow = out + ((size_t)b * H + oy) * W * C;
            for (int ox = 0; ox < W; ++ox) {
                float projection = t[6] * (float)ox + t[7] * (float)oy + 1.0f;
                if (projection == 0.0f) {
                    for (int c = 0; c < C; ++c) orow[ox * C + c] = fill_value;
                    continue;
                }
                float x = (t[0] * (float)ox + t[1] * (float)oy + t[2]) / projection;
                float y = (t[3] * (float)ox + t[4] * (float)oy + t[5]) / projection;
                if (interp == ORC_INTERP_NEAREST) {
                    long yy = (long)roundf(y), xx = (long)roundf(x);
                    for (int c = 0; c < C; ++c) orow[ox * C + c] = read_fill(img, H, W, C, yy, xx, c, fill_value);
                } else {
                    float y_floor = floorf(y), x_floor = floorf(x);
                    float y_ceil = y_floor + 1.0f, x_ceil = x_floor + 1.0f;
                    long y0 = (long)y_floor, x0 = (long)x_floor, y1 = (long)y_ceil, x1 = (long)x_ceil;
                    for (int c = 0; c < C; ++c) {
                        float v_yfloor = (x_ceil - x) * read_fill(img, H, W, C, y0, x0, c, fill_value) +
                                         (x - x_floor) * read_fill(img, H, W, C, y0, x1, c, fill_value);
                        float v_yceil = (x_ceil - x) * read_fill(img, H, W, C, y1, x0, c, fill_value) +
                                        (x - x_floor) * read_fill(img, H, W, C, y1, x1, c, fill_value);
                        orow[ox * C + c] = (y_ceil - y) * v_yfloor + (y - y_floor) * v_yceil;
                    }
                }
            }
        }
    }
}

/* ---- ResizeBilinear, half_pixel_centers=True --------------------------- */

typedef struct { int lower, upper; float lerp; } interp_w;

static void resize_weights(int out_size, int in_size, interp_w* w) {
    /* compute_interpolation_weights with HalfPixelScaler; scale = in/out in fp32 */
    float scale = (float)in_size / (float)out_size;
    for (int i = 0; i < out_size; ++i) {
        float in = ((float)i + 0.5f) * scale - 0.5f;
        float in_f = floorf(in);
        int lo = (int)in_f; if (lo < 0) lo = 0;
        int hi = (int)ceilf(in); if (hi > in_size - 1) hi = in_size - 1;
        w[i].lower = lo; w[i].upper = hi; w[i].lerp = in - in_f;
    }
}

void orc_resize_bilinear(const float* in, int N, int h, int w, int C, float* out, int H, int W) {
    interp_w* ys = (interp_w*)malloc(sizeof(interp_w) * H);
    interp_w* xs = (interp_w*)malloc(sizeof(interp_w) * W);
    resize_weights(H, h, ys);
    resize_weights(W, w, xs);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < N; ++b) {
        for (int y = 0; y < H; ++y) {
            const float* img = in + (size_t)b * h * w * C;
            const float* top = img + (size_t)ys[y].lower * w * C;
            const float* bot = img + (size_t)ys[y].upper * w * C;
            float yl = ys[y].lerp;
            float* orow = out + ((size_t)b * H + y) * W * C;
            for (int x = 0; x < W; ++x) {
                int xl = xs[x].lower * C, xu = xs[x].upper * C;
                float xlerp = xs[x].lerp;
                for (int c = 0; c < C; ++c) {
                    float tl = top[xl + c], tr = top[xu + c], bl = bot[xl + c], br = bot[xu + c];
                    float t = tl + (tr - tl) * xlerp;
                    float bb = bl + (br - bl) * xlerp;
                    orow[x * C + c] = t + (bb - t) * yl;
                }
            }
        }
    }
    free(ys); free(xs);
}

void orc_resize_bilinear_grad(const float* grad, int N, int H, int W, int C, float* in_grad, int h, int w) {
    /* ResizeBilinearGrad CPU functor: here (H,W) is the *resized* (small) side the gradient arrives on
     * and (h,w) the original (large) side it is scattered to. Argument naming follows the op:
     * grad [N,H,W,C] -> in_grad [N,h,w,C]. */
    float hs = (float)h / (float)H, ws = (float)w / (float)W;
    memset(in_grad, 0, sizeof(float) * (size_t)N * h * w * C);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < N; ++b) {
        float* ig = in_grad + (size_t)b * h * w * C;
        const float* g = grad + (size_t)b * H * W * C;
        for (int y = 0; y < H; ++y) {
            float in_y = ((float)y + 0.5f) * hs - 0.5f;
            int top = (int)floorf(in_y); if (top < 0) top = 0;
            int bot = (int)ceilf(in_y); if (bot > h - 1) bot = h - 1;
            float y_lerp = in_y - floorf(in_y);
            float inv_y_lerp = 1.0f - y_lerp;
            for (int x = 0; x < W; ++x) {
                float in_x = ((float)x + 0.5f) * ws - 0.5f;
                int left = (int)floorf(in_x); if (left < 0) left = 0;
                int right = (int)ceilf(in_x); if (right > w - 1) right = w - 1;
                float x_lerp = in_x - floorf(in_x);
                float inv_x_lerp = 1.0f - x_lerp;
                for (int c = 0; c < C; ++c) {
                    float gv = g[((size_t)y * W + x) * C + c];
                    float dtop = inv_y_lerp * gv;
                    ig[((size_t)top * w + left) * C + c] += dtop * inv_x_lerp;
                    ig[((size_t)top * w + right) * C + c] += dtop * x_lerp;
                    float dbot = y_lerp * gv;
                    ig[((size_t)bot * w + left) * C + c] += dbot * inv_x_lerp;
                    ig[((size_t)bot * w + right) * C + c] += dbot * x_lerp;
                }
            }
        }
    }
}

/* ---- bilateral TV (superresolution.py:8-23), value and gradient --------------------------- */

static float sgnf(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

static double btv_value_and_grad(const float* x, int H, int W, float scale, float* grad_acc) {
    /* 15 integer (h,v) shifts, h in [-2,2], v in [0,2]; tfa.image.translate default NEAREST, zero fill;
     * weight 0.6^(|h|+|v|).  Gradient: sign(diff)*w to x, and -(that) through the translate op's
     * registered gradient (same op, inverse transform, NEAREST). scale = lambda_tv. */
    const float alpha = 0.6f;
    double total = 0.0;
    size_t n = (size_t)H * W;
    float* shifted = (float*)malloc(sizeof(float) * n);
    float* gdiff = (float*)malloc(sizeof(float) * n);
    float* gback = (float*)malloc(sizeof(float) * n);
    for (int hh = -2; hh <= 2; ++hh) {
        for (int vv = 0; vv <= 2; ++vv) {
            float t[8], tinv[8];
            orc_translate_matrix((float)hh, (float)vv, t);
            orc_projective_transform(x, 1, H, W, 1, t, 1, ORC_INTERP_NEAREST, 0.0f, shifted);
            float wgt = powf(alpha, (float)(abs(hh) + abs(vv)));
            double l1 = 0.0;
            for (size_t i = 0; i < n; ++i) {
                float d = x[i] - shifted[i];
                l1 += fabs((double)d);
                gdiff[i] = (scale * wgt) * sgnf(d);
            }
            total += (double)wgt * l1;
            if (grad_acc) {
                orc_invert_transform(t, tinv);
                orc_projective_transform(gdiff, 1, H, W, 1, tinv, 1, ORC_INTERP_NEAREST, 0.0f, gback);
                for (size_t i = 0; i < n; ++i) grad_acc[i] += gdiff[i] - gback[i];
            }
        }
    }
    free(shifted); free(gdiff); free(gback);
    return total;
}

/* ---- loss_function + tape.gradient (superresolution.py:44-100, :126-133) --- */

float orc_loss_and_grad(const float* x, int H, int W, const float* copies, int N, int h, int w,
                        const float* angles, const float* shifts, const orc_params* p,
                        const uint8_t* keep, float* grad, float* resid_out) {
    /* copy dropout (:47-53): boolean_mask of copies / angles / shifts */
    int M = 0;
    int* idx = (int*)malloc(sizeof(int) * N);
    for (int k = 0; k < N; ++k) if (!keep || keep[k]) idx[M++] = k;

    size_t hr = (size_t)H * W, lr = (size_t)h * w;
    float* rot_t = (float*)malloc(sizeof(float) * 8 * M);
    float* tr_t = (float*)malloc(sizeof(float) * 8 * M);
    for (int j = 0; j < M; ++j) {
        orc_rotate_matrix(angles[idx[j]], H, W, rot_t + 8 * j);
        orc_translate_matrix(shifts[2 * idx[j]], shifts[2 * idx[j] + 1], tr_t + 8 * j);
    }
    float* a = (float*)malloc(sizeof(float) * hr * M);
    float* b = (float*)malloc(sizeof(float) * hr * M);
    float* d = (float*)malloc(sizeof(float) * lr * M);

    /* tf.tile (:59-60) */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < M; ++j) memcpy(a + hr * j, x, sizeof(float) * hr);
    /* rotate (:61-62), translate (:63-64) */
    orc_projective_transform(a, M, H, W, 1, rot_t, M, ORC_INTERP_BILINEAR, 0.0f, b);
    orc_projective_transform(b, M, H, W, 1, tr_t, M, ORC_INTERP_BILINEAR, 0.0f, a);
    /* D operator (:67-68) */
    orc_resize_bilinear(a, M, H, W, 1, d, h, w);

    /* data term (:71-72) and its gradient 2*lambda_df*(D - y) */
    double df = 0.0;
    float two_ldf = 2.0f * p->lambda_df;
#pragma omp parallel for schedule(static) reduction(+ : df)
    for (int j = 0; j < M; ++j) {
        const float* y = copies + lr * idx[j];
        float* dj = d + lr * j;
        for (size_t i = 0; i < lr; ++i) {
            float r = dj[i] - y[i];
            df += (double)r * (double)r;
            if (resid_out) resid_out[lr * idx[j] + i] = r;
            dj[i] = two_ldf * r;
        }
    }
    if (resid_out)
        for (int k = 0; k < N; ++k)
            if (keep && !keep[k]) memset(resid_out + lr * k, 0, sizeof(float) * lr);

    /* regularisers (:78-98) */
    double tv = 0.0, l2 = 0.0, l1 = 0.0;
    for (size_t i = 0; i < hr; ++i) { l2 += (double)x[i] * (double)x[i]; l1 += fabs((double)x[i]); }

    if (grad) {
        /* backward: ResizeBilinearGrad, then the warp op's registered gradient twice */
        orc_resize_bilinear_grad(d, M, h, w, 1, a, H, W);
        float* inv_t = (float*)malloc(sizeof(float) * 8 * M);
        for (int j = 0; j < M; ++j) orc_invert_transform(tr_t + 8 * j, inv_t + 8 * j);
        orc_projective_transform(a, M, H, W, 1, inv_t, M, ORC_INTERP_BILINEAR, 0.0f, b);
        for (int j = 0; j < M; ++j) orc_invert_transform(rot_t + 8 * j, inv_t + 8 * j);
        orc_projective_transform(b, M, H, W, 1, inv_t, M, ORC_INTERP_BILINEAR, 0.0f, a);
        free(inv_t);
        /* gradient of tf.tile: sum over the copy axis, ascending k, sequential fp32 adds */
#pragma omp parallel for schedule(static)
        for (int yy = 0; yy < H; ++yy) {
            for (int xx = 0; xx < W; ++xx) {
                size_t i = (size_t)yy * W + xx;
                float s = 0.0f;
                for (int j = 0; j < M; ++j) s += a[hr * j + i];
                grad[i] = s;
            }
        }
    }

    if (p->use_btv) {
        tv = btv_value_and_grad(x, H, W, p->lambda_tv, grad);
    } else {
        /* tf.image.image_gradients (:81-83): dy[i]=x[i+1]-x[i] (last row 0), dx likewise */
        for (int yy = 0; yy < H; ++yy) {
            for (int xx = 0; xx < W; ++xx) {
                size_t i = (size_t)yy * W + xx;
                float dyv = (yy < H - 1) ? x[i + W] - x[i] : 0.0f;
                float dxv = (xx < W - 1) ? x[i + 1] - x[i] : 0.0f;
                tv += fabs((double)dyv) + fabs((double)dxv);
                if (grad) {
                    float gy = p->lambda_tv * sgnf(dyv), gx = p->lambda_tv * sgnf(dxv);
                    if (yy < H - 1) { grad[i + W] += gy; grad[i] -= gy; }
                    if (xx < W - 1) { grad[i + 1] += gx; grad[i] -= gx; }
                }
            }
        }
    }
    if (grad) {
        for (size_t i = 0; i < hr; ++i) {
            grad[i] += p->lambda_l2 * (x[i] * 2.0f);
            if (p->lambda_l1 > 0.0f) grad[i] += p->lambda_l1 * sgnf(x[i]);
        }
    }

    float loss = p->lambda_df * (float)df + p->lambda_tv * (float)tv;
    loss = loss + p->lambda_l2 * (float)l2;
    if (p->lambda_l1 > 0.0f) loss = loss + p->lambda_l1 * (float)l1;

    free(a); free(b); free(d); free(rot_t); free(tr_t); free(idx);
    return loss;
}

/* ---- Optimizer (optimizer.py:4-52) -------------------------------------------- */

static float lr_at(const orc_params* p, int i) {
    if (!p->lr_scheduler) return p->learning_rate;
    /* ExponentialDecay, non-staircase, fp32: lr0 * rate^(i/steps) */
    float pw = (float)i / p->decay_steps;
    return p->learning_rate * powf(p->decay_rate, pw);
}

typedef struct { float *s0, *s1, *s2; } opt_state;

static void opt_step(const orc_params* p, float lr, int64_t t, float* x, const float* g, opt_state* st, size_t n) {
    switch (p->optimizer) {
    case ORC_OPT_SGD: {
        if (p->momentum == 0.0f) {
            for (size_t i = 0; i < n; ++i) x[i] -= lr * g[i];                       /* ApplyGradientDescent */
        } else {
            for (size_t i = 0; i < n; ++i) {                                         /* ApplyKerasMomentum */
                st->s0[i] = st->s0[i] * p->momentum - lr * g[i];
                if (p->nesterov) x[i] += st->s0[i] * p->momentum - lr * g[i];
                else x[i] += st->s0[i];
            }
        }
    } break;
    case ORC_OPT_ADAGRAD: {                                                          /* ApplyAdagradV2 */
        for (size_t i = 0; i < n; ++i) {
            st->s0[i] += g[i] * g[i];
            x[i] -= g[i] * lr / (sqrtf(st->s0[i]) + p->epsilon);
        }
    } break;
    case ORC_OPT_ADADELTA: {                                                         /* ApplyAdadelta, rho=.95 eps=1e-7 */
        const float rho = 0.95f, eps = 1e-7f;
        for (size_t i = 0; i < n; ++i) {
            st->s0[i] = st->s0[i] * rho + g[i] * g[i] * (1.0f - rho);
            float upd = sqrtf(st->s1[i] + eps) * (1.0f / sqrtf(st->s0[i] + eps)) * g[i];
            x[i] -= upd * lr;
            st->s1[i] = st->s1[i] * rho + upd * upd * (1.0f - rho);
        }
    } break;
    case ORC_OPT_ADAMAX: {                                                           /* ApplyAdaMax */
        float b1p = powf(p->beta_1, (float)t);
        for (size_t i = 0; i < n; ++i) {
            st->s0[i] += (g[i] - st->s0[i]) * (1.0f - p->beta_1);
            float bv = p->beta_2 * st->s1[i], ag = fabsf(g[i]);
            st->s1[i] = bv > ag ? bv : ag;
            x[i] -= lr / (1.0f - b1p) * (st->s0[i] / (st->s1[i] + p->epsilon));
        }
    } break;
    default: {                                                                       /* ApplyAdam[WithAmsgrad] */
        float b1p = powf(p->beta_1, (float)t), b2p = powf(p->beta_2, (float)t);
        float alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
        for (size_t i = 0; i < n; ++i) {
            st->s0[i] += (g[i] - st->s0[i]) * (1.0f - p->beta_1);
            st->s1[i] += (g[i] * g[i] - st->s1[i]) * (1.0f - p->beta_2);
            if (p->amsgrad) {
                st->s2[i] = st->s2[i] > st->s1[i] ? st->s2[i] : st->s1[i];
                x[i] -= (st->s0[i] * alpha) / (sqrtf(st->s2[i]) + p->epsilon);
            } else {
                x[i] -= (st->s0[i] * alpha) / (sqrtf(st->s1[i]) + p->epsilon);
            }
        }
    } break;
    }
}

float orc_augmented_superresolution(const float* copies, int N, int h, int w, int H, int W,
                                    const float* angles, const float* shifts, const orc_params* p,
                                    const uint8_t* keep, float* x_out,
                                    const int32_t* trace_iters, int n_trace, float* trace) {
    size_t hr = (size_t)H * W;
    /* initial value: bilinear upsample of the un-augmented copy 0 (:112-113) */
    orc_resize_bilinear(copies, 1, h, w, 1, x_out, H, W);
    float* g = (float*)malloc(sizeof(float) * hr);
    opt_state st;
    st.s0 = (float*)calloc(hr, sizeof(float));
    st.s1 = (float*)calloc(hr, sizeof(float));
    st.s2 = (float*)calloc(hr, sizeof(float));
    if (p->optimizer == ORC_OPT_ADAGRAD)
        for (size_t i = 0; i < hr; ++i) st.s0[i] = p->initial_accumulator_value;
    float loss = 0.0f;
    for (int i = 0; i < p->num_iter; ++i) {
        float lr = lr_at(p, i);                                                     /* :121-122 */
        loss = orc_loss_and_grad(x_out, H, W, copies, N, h, w, angles, shifts, p, keep, g, NULL);
        int64_t t = p->step_offset + (int64_t)i + 1;                                /* iterations + 1 */
        opt_step(p, lr, t, x_out, g, &st, hr);                                      /* :134-135 */
        for (int j = 0; j < n_trace; ++j)
            if (trace_iters[j] == i + 1) memcpy(trace + hr * j, x_out, sizeof(float) * hr);
    }
    free(g); free(st.s0); free(st.s1); free(st.s2);
    return loss;
}

/* ---- max / mean back-projection (superresolution.py:139-161) ------------------------------ */

void orc_backproject(const float* copies, int N, int h, int w, int H, int W,
                     const float* angles, const float* shifts, int mode, float* out) {
    size_t hr = (size_t)H * W;
    float* a = (float*)malloc(sizeof(float) * hr * N);
    float* b = (float*)malloc(sizeof(float) * hr * N);
    float* t = (float*)malloc(sizeof(float) * 8 * N);
    orc_resize_bilinear(copies, N, h, w, 1, a, H, W);
    for (int k = 0; k < N; ++k) orc_translate_matrix(-shifts[2 * k], -shifts[2 * k + 1], t + 8 * k);
    orc_projective_transform(a, N, H, W, 1, t, N, ORC_INTERP_BILINEAR, 0.0f, b);
    for (int k = 0; k < N; ++k) orc_rotate_matrix(-angles[k], H, W, t + 8 * k);
    orc_projective_transform(b, N, H, W, 1, t, N, ORC_INTERP_BILINEAR, 0.0f, a);
#pragma omp parallel for schedule(static)
    for (int yy = 0; yy < H; ++yy) {
        for (int xx = 0; xx < W; ++xx) {
            size_t i = (size_t)yy * W + xx;
            if (mode == 0) {
                float m = a[i];
                for (int k = 1; k < N; ++k) m = a[hr * k + i] > m ? a[hr * k + i] : m;
                out[i] = m;
            } else {
                float s = 0.0f;
                for (int k = 0; k < N; ++k) s += a[hr * k + i];
                out[i] = s / (float)N;
            }
        }
    }
    free(a); free(b); free(t);
}

/* ---- post-processing (superres_utils.py) -------------------------------------------------- */

void orc_threshold(const float* x, int64_t n, int32_t th_value, float th_factor, const float* th_mask, int32_t* out) {
    if (th_mask) {
        for (int64_t i = 0; i < n; ++i) out[i] = (x[i] >= th_mask[i]) ? th_value : 0;
    } else {
        float mx = x[0];
        for (int64_t i = 1; i < n; ++i) mx = x[i] > mx ? x[i] : mx;
        float th = mx * th_factor;
        for (int64_t i = 0; i < n; ++i) out[i] = (x[i] > th) ? th_value : 0;
    }
}

void orc_minmax_normalize_global(const float* in, int64_t n, float new_min, float new_max, float* out) {
    float mn = in[0], mx = in[0];
    for (int64_t i = 1; i < n; ++i) { mn = in[i] < mn ? in[i] : mn; mx = in[i] > mx ? in[i] : mx; }
    float den = (mx - mn) != 0.0f ? (mx - mn) : 1.0f;
    for (int64_t i = 0; i < n; ++i) {
        float num = (in[i] - mn) * (new_max - new_min);
        out[i] = new_min + (num / den);
    }
}

void orc_opm_extract(const float* logits, int N, int h, int w, int K, int class_id, int mode,
                     float* class_out, float* max_out) {
    size_t px = (size_t)h * w;
    for (int n = 0; n < N; ++n) {
        const float* lg = logits + (size_t)n * px * K;
        float* co = class_out + (size_t)n * px;
        if (mode == ORC_OPM_ARGMAX) {
            for (size_t i = 0; i < px; ++i) {
                int best = 0;
                for (int k = 1; k < K; ++k) if (lg[i * K + k] > lg[i * K + best]) best = k;   /* ties -> lowest index */
                co[i] = (best == class_id) ? (float)class_id : 0.0f;
            }
        } else if (mode == ORC_OPM_SLICE) {
            /* per-copy min/max over ALL channels (augmentation_utils.py:100-104) */
            float mn = lg[0], mx = lg[0];
            for (size_t i = 1; i < px * K; ++i) { mn = lg[i] < mn ? lg[i] : mn; mx = lg[i] > mx ? lg[i] : mx; }
            float den = (mx - mn) != 0.0f ? (mx - mn) : 1.0f;
            for (size_t i = 0; i < px; ++i) {
                float num = (lg[i * K + class_id] - mn) * (1.0f - 0.0f);
                co[i] = 0.0f + (num / den);
            }
        } else {
            float* mo = max_out + (size_t)n * px;
            for (size_t i = 0; i < px; ++i) {
                co[i] = lg[i * K + class_id];
                float m = -INFINITY;
                for (int k = 0; k < K; ++k) if (k != class_id && lg[i * K + k] > m) m = lg[i * K + k];
                mo[i] = m;
            }
        }
    }
}

double orc_single_class_iou(const int32_t* y_true, const int32_t* y_pred, int64_t n, int class_id, int include_bg) {
    int classes[2] = {class_id, 0};
    int nc = include_bg ? 2 : 1;
    double sum = 0.0; int cnt = 0;
    for (int ci = 0; ci < nc; ++ci) {
        int64_t inter = 0, uni = 0;
        for (int64_t i = 0; i < n; ++i) {
            int32_t yt = y_true[i];
            if (include_bg && yt != class_id) yt = 0;       /* utils.py:185-189 */
            int tl = (yt == classes[ci]), pl = (y_pred[i] == classes[ci]);
            inter += (tl & pl); uni += (tl | pl);
        }
        if (uni > 0) { sum += (double)inter / (double)uni; ++cnt; }   /* NaN entries dropped */
    }
    return cnt ? sum / cnt : NAN;
}
