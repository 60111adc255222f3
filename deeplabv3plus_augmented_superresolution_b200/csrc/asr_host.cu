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

#include <atomic>
#include <mutex>
#include <vector>

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

// ---- launch accounting and optional kernel timing ------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
struct ProfSpan { cudaEvent_t a, b; int slot; };
static std::vector<ProfSpan> g_spans;
static std::vector<cudaEvent_t> g_pool;
struct OpenMark { int slot; cudaStream_t st; cudaEvent_t ev; };
static std::vector<OpenMark> g_open;   // begin marks waiting for their end mark, keyed by (slot, stream)

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool profile_enabled() { return g_prof_on; }

static cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void profile_mark(int slot, cudaStream_t st, bool begin) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        cudaEvent_t e = take_event();
        cudaEventRecord(e, st);
        g_open.push_back(OpenMark{slot, st, e});
    } else {
        for (size_t i = g_open.size(); i-- > 0;) {
            if (g_open[i].slot != slot || g_open[i].st != st) continue;
            cudaEvent_t e = take_event();
            cudaEventRecord(e, st);
            g_spans.push_back(ProfSpan{g_open[i].ev, e, slot});
            g_open.erase(g_open.begin() + (long)i);
            return;
        }
    }
}

bool first_use_on_device(unsigned long long* mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const unsigned long long bit = 1ull << dev;
    return (__atomic_fetch_or(mask, bit, __ATOMIC_RELAXED) & bit) == 0;
}

}  // namespace asr

extern "C" long long asr_kernel_launches(void) { return asr::g_launches.load(); }

extern "C" int asr_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(asr::g_prof_mu);
    asr::g_prof_on = on != 0;
    return ASR_OK;
}

extern "C" int asr_profile_read(double* ms, long long* count) {
    if (!ms || !count) return asr::fail(ASR_ENULL, "null argument");
    std::lock_guard<std::mutex> lk(asr::g_prof_mu);
    ms[0] = ms[1] = 0.0;
    count[0] = count[1] = 0;
    for (auto& s : asr::g_spans) {
        ASR_CUDA_TRY(cudaEventSynchronize(s.b));
        float t = 0.f;
        ASR_CUDA_TRY(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.slot] += t;
        count[s.slot] += 1;
        asr::g_pool.push_back(s.a);
        asr::g_pool.push_back(s.b);
    }
    asr::g_spans.clear();
    return ASR_OK;
}

extern "C" int asr_version(void) { return ASR_VERSION; }
extern "C" const char* asr_last_error(void) { return asr::g_err; }
