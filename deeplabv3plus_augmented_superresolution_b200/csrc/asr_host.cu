// asr_host.cu -- host-side helpers of libasr: error plumbing and the per-copy transform tables.
//
// The transform coefficients are computed on the host in fp32, exactly as the TensorFlow-addons
// Python helpers do (tfa.image.angles_to_projective_transforms / translations_to_projective_transforms,
// called from superresolution_scripts/superresolution.py:61-64 and augmentation_utils.py:22-25), and the
// inverse used by the registered gradient of ImageProjectiveTransformV3 is a fp32 3x3 LU inverse
// (tf.linalg.inv), renormalised by its [2][2] element.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "asr_common.cuh"

namespace asr {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void rotate_matrix(float angle, int H, int W, float t[8]) {
    const float c = cosf(angle), s = sinf(angle);
    const float wm1 = (float)W - 1.0f, hm1 = (float)H - 1.0f;
    const float cw = c * wm1, sh = s * hm1, sw = s * wm1, ch = c * hm1;
    t[0] = c;  t[1] = -s; t[2] = (wm1 - (cw - sh)) / 2.0f;
    t[3] = s;  t[4] = c;  t[5] = (hm1 - (sw + ch)) / 2.0f;
    t[6] = 0.0f; t[7] = 0.0f;
}

// 3x3 inverse by LU with partial pivoting, every operation rounded to fp32; columns of the
// inverse by forward/back substitution against the permuted identity.
void invert_transform(const float t[8], float tinv[8]) {
    float lu[9] = {t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], 1.0f};
    int row_of[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        int p = k;
        for (int r = k + 1; r < 3; ++r)
            if (fabsf(lu[3 * r + k]) > fabsf(lu[3 * p + k])) p = r;
        if (p != k) {
            for (int c = 0; c < 3; ++c) { float tmp = lu[3 * k + c]; lu[3 * k + c] = lu[3 * p + c]; lu[3 * p + c] = tmp; }
            int tr = row_of[k]; row_of[k] = row_of[p]; row_of[p] = tr;
        }
        for (int r = k + 1; r < 3; ++r) {
            const float l = lu[3 * r + k] / lu[3 * k + k];
            lu[3 * r + k] = l;
            for (int c = k + 1; c < 3; ++c) lu[3 * r + c] = lu[3 * r + c] - l * lu[3 * k + c];
        }
    }
    float inv[9];
    for (int col = 0; col < 3; ++col) {
        float y[3];
        for (int r = 0; r < 3; ++r) y[r] = (row_of[r] == col) ? 1.0f : 0.0f;
        for (int r = 1; r < 3; ++r)
            for (int c = 0; c < r; ++c) y[r] = y[r] - lu[3 * r + c] * y[c];
        for (int r = 2; r >= 0; --r) {
            for (int c = r + 1; c < 3; ++c) y[r] = y[r] - lu[3 * r + c] * y[c];
            y[r] = y[r] / lu[3 * r + r];
        }
        for (int r = 0; r < 3; ++r) inv[3 * r + col] = y[r];
    }
    for (int i = 0; i < 8; ++i) tinv[i] = inv[i] / inv[8];
}

}  // namespace asr

extern "C" int asr_version(void) { return ASR_VERSION; }
extern "C" const char* asr_last_error(void) { return asr::g_err; }
