// asr_aux.cu -- producers and consumers either side of the solve:
//   asr_warp_affine         augmentation_utils.py:11-27   create_augmented_copies (rotate -> translate)
//   asr_opm_extract         augmentation_utils.py:80-115  argmax / slice / slice_max OPM (+ utils.py:115-119)
//   asr_minmax_normalize    superres_utils.py:56-62,186-194
//   asr_backproject_batched superresolution.py:139-161    max_superresolution / mean_superresolution
//   asr_threshold           superres_utils.py:118-139     threshold_image
// Same numerical contract as the solve: un-fused fp32, TensorFlow's evaluation order.
#include <math.h>
#include <vector>

#include "asr_common.cuh"

namespace asr {

// monotone float <-> uint map so that min/max reductions can use integer atomics (exact, order-free)
__device__ __forceinline__ unsigned enc(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned e) {
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// ================================================================================================
// K3: augmentation warp.  out[k] = translate(rotate(image, angle_k), shift_k), zero fill.
// ================================================================================================
// One CTA per 32x32 output tile per copy.  The translate stage reads the rotated image p only on
// the integer window (Z+s, Z+s+1), s = floor(-d), so the CTA first evaluates p once per position of
// the 33x33 window into shared memory (each p would otherwise be recomputed by its 4 consumers),
// then forms the outputs with per-column/row tap tables carrying the literal translate weights.
// The source image is first padded to 4 channels so that every tap is one 128-bit load, p lives in
// shared memory as float4, and the finished tile is staged in shared memory and written with
// row-contiguous 128-bit stores (a 32-pixel RGB row segment is 384 contiguous bytes).
constexpr int K3_T = 32;
constexpr int K3_P = K3_T + 1;
constexpr int K3_THREADS = 256;
constexpr int K3_CMAX = 4;

struct WarpXf { float r0, r1, r2, r3, r4, r5, tx, ty; };

__device__ __forceinline__ float2 warp_taps(int Z, float t, int s, int limit, int interp) {
    const float iz = fadd((float)Z, t);
    const int base = Z + s;
    float wa, wb;
    if (interp == ASR_INTERP_BILINEAR) {
        const float f = floorf(iz);
        const float w0 = fsub(fadd(f, 1.0f), iz), w1 = fsub(iz, f);
        if ((int)f == base) { wa = w0; wb = w1; } else { wa = 0.0f; wb = w0; }
    } else {  // NEAREST: std::round, half away from zero
        const int n = (int)roundf(iz);
        wa = (n == base) ? 1.0f : 0.0f;
        wb = (n == base + 1) ? 1.0f : 0.0f;
    }
    if (base < 0 || base >= limit) wa = 0.0f;
    if (base + 1 < 0 || base + 1 >= limit) wb = 0.0f;
    return make_float2(wa, wb);
}

__global__ void k_pad_channels(const float* __restrict__ img, float4* __restrict__ out, size_t npx, int C) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (size_t)gridDim.x * blockDim.x) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) v[c] = img[i * C + c];
        out[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// bilinear_interpolation() of the op on the C live channels of four padded pixels: channel pairs travel as the two
// lanes of packed fp32 instructions with the weights broadcast (asr_common.cuh: products packed, sums scalar)
template <int C>
__device__ __forceinline__ float4 bilerpC(float4 a, float4 b, float4 c, float4 d, float wx0, float wx1, float wy0, float wy1) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const f32x2 x0 = pk(wx0, wx0), x1 = pk(wx1, wx1), y0 = pk(wy0, wy0), y1 = pk(wy1, wy1);
    if (C >= 2) {
        const f32x2 r = bilerp2(pk(a.x, a.y), pk(b.x, b.y), pk(c.x, c.y), pk(d.x, d.y), x0, x1, y0, y1);
        o.x = pk_lo(r); o.y = pk_hi(r);
    } else {
        o.x = bilerp(a.x, b.x, c.x, d.x, wx0, wx1, wy0, wy1);
    }
    if (C == 4) {
        const f32x2 r = bilerp2(pk(a.z, a.w), pk(b.z, b.w), pk(c.z, c.w), pk(d.z, d.w), x0, x1, y0, y1);
        o.z = pk_lo(r); o.w = pk_hi(r);
    } else if (C == 3) {
        o.z = bilerp(a.z, b.z, c.z, d.z, wx0, wx1, wy0, wy1);
    }
    return o;
}

template <int C>
__global__ void __launch_bounds__(K3_THREADS)
k_warp_affine(const float4* __restrict__ img, const WarpXf* __restrict__ xf, float* __restrict__ out, int H, int W, int interp) {
    __shared__ float4 p[K3_P * K3_P];
    __shared__ __align__(16) float outs[K3_T * K3_T * C];
    __shared__ float2 colw[K3_T], roww[K3_T];
    const int k = blockIdx.y, tid = threadIdx.x;
    const int ntx = (W + K3_T - 1) / K3_T;
    const int X0 = (blockIdx.x % ntx) * K3_T, Y0 = (blockIdx.x / ntx) * K3_T;
    const WarpXf T = xf[k];
    const int sx = (int)floorf(T.tx), sy = (int)floorf(T.ty);
    const int qx_lo = X0 + sx, qy_lo = Y0 + sy;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    if (tid < K3_T) colw[tid] = warp_taps(X0 + tid, T.tx, sx, W, interp);
    else if (tid < 2 * K3_T) roww[tid - K3_T] = warp_taps(Y0 + tid - K3_T, T.ty, sy, H, interp);

    // ---- rotated image on the (T+1)^2 window ---------------------------------------------------------
    for (int e = tid; e < K3_P * K3_P; e += K3_THREADS) {
        const int py = e / K3_P, px = e - py * K3_P;
        const float qx = (float)(qx_lo + px), qy = (float)(qy_lo + py);
        const float ix = affine_coord(T.r0, qx, T.r1, qy, T.r2), iy = affine_coord(T.r3, qx, T.r4, qy, T.r5);
        float4 v;
        if (interp == ASR_INTERP_BILINEAR) {
            // floor of both coordinates with the magic-number add (asr_common.cuh), x and y as the two packed lanes
            const f32x2 ixy = pk(ix, iy), magic2 = pk(kMagic, kMagic);
            const f32x2 raw = add2_rd(ixy, magic2);
            const f32x2 fl2 = sub2(raw, magic2);
            const f32x2 w1 = sub2(ixy, fl2), w0 = sub2(add2(fl2, pk(1.0f, 1.0f)), ixy);   // (floor + 1) - v, v - floor
            const float wx0 = pk_lo(w0), wx1 = pk_lo(w1), wy0 = pk_hi(w0), wy1 = pk_hi(w1);
            const int x0 = (int)__float_as_uint(pk_lo(raw)) - kMagicBits, y0 = (int)__float_as_uint(pk_hi(raw)) - kMagicBits;
            const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
            const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
            const float4* b00 = img + ((ptrdiff_t)y0 * W + x0);
            const float4 v00 = (vy0 && vx0) ? __ldg(b00) : zero4;
            const float4 v01 = (vy0 && vx1) ? __ldg(b00 + 1) : zero4;
            const float4 v10 = (vy1 && vx0) ? __ldg(b00 + W) : zero4;
            const float4 v11 = (vy1 && vx1) ? __ldg(b00 + W + 1) : zero4;
            v = bilerpC<C>(v00, v01, v10, v11, wx0, wx1, wy0, wy1);
        } else {
            const long xn = (long)roundf(ix), yn = (long)roundf(iy);
            v = (xn >= 0 && xn < W && yn >= 0 && yn < H) ? __ldg(img + ((size_t)yn * W + xn)) : zero4;
        }
        p[e] = v;
    }
    __syncthreads();

    // ---- translate stage: 4 output pixels per thread into the staging tile -----------------------------
#pragma unroll
    for (int j = 0; j < K3_T * K3_T / K3_THREADS; ++j) {
        const int ty = (tid >> 5) + 8 * j, tx = tid & 31;
        const float2 wc = colw[tx], wr = roww[ty];
        const float4 a = p[ty * K3_P + tx], b = p[ty * K3_P + tx + 1], c = p[(ty + 1) * K3_P + tx], d = p[(ty + 1) * K3_P + tx + 1];
        float4 v;
        if (interp == ASR_INTERP_BILINEAR) {
            v = bilerpC<C>(a, b, c, d, wc.x, wc.y, wr.x, wr.y);
        } else {   // exactly one tap has weight 1 (or none: zero fill)
            const float4 top = (wc.x != 0.0f) ? a : ((wc.y != 0.0f) ? b : zero4);
            const float4 bot = (wc.x != 0.0f) ? c : ((wc.y != 0.0f) ? d : zero4);
            v = (wr.x != 0.0f) ? top : ((wr.y != 0.0f) ? bot : zero4);
        }
        float* o = outs + (ty * K3_T + tx) * C;
        o[0] = v.x;
        if (C > 1) o[1] = v.y;
        if (C > 2) o[2] = v.z;
        if (C > 3) o[3] = v.w;
    }
    __syncthreads();

    // ---- write the tile: each row segment is K3_T*C contiguous floats ------------------------------------
    float* ok = out + (size_t)k * H * W * C;
    const bool full = (X0 + K3_T <= W) && (Y0 + K3_T <= H) && (((size_t)W * C) % 4 == 0) && ((X0 * C) % 4 == 0) && ((K3_T * C) % 4 == 0);
    if (full) {
        constexpr int ROW4 = K3_T * C / 4;
        for (int e = tid; e < K3_T * ROW4; e += K3_THREADS) {
            const int ty = e / ROW4, c4 = e - ty * ROW4;
            const float4 v = *reinterpret_cast<const float4*>(outs + ty * K3_T * C + 4 * c4);
            *reinterpret_cast<float4*>(ok + ((size_t)(Y0 + ty) * W + X0) * C + 4 * c4) = v;
        }
    } else {
        for (int e = tid; e < K3_T * K3_T * C; e += K3_THREADS) {
            const int ty = e / (K3_T * C), rem = e - ty * (K3_T * C), tx = rem / C;
            if (X0 + tx < W && Y0 + ty < H) ok[((size_t)(Y0 + ty) * W + X0) * C + rem] = outs[e];
        }
    }
}

// ================================================================================================
// K4: OPM extraction from NHWC logits
// ================================================================================================
constexpr int K4_THREADS = 256;
constexpr int K4_KMAX = 64;

// per-copy min / max over all K channels (slice mode, augmentation_utils.py:100-101)
__global__ void k_copy_minmax(const float* __restrict__ logits, size_t per_copy, unsigned* __restrict__ mm) {
    const int n = blockIdx.y;
    const float* p = logits + (size_t)n * per_copy;
    float lo = INFINITY, hi = -INFINITY;
    const size_t n4 = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? per_copy / 4 : 0;   // odd per-copy sizes misalign every other copy: scalar path
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(p4 + i);
        lo = fminf(fminf(lo, fminf(v.x, v.y)), fminf(v.z, v.w));
        hi = fmaxf(fmaxf(hi, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_copy; i += (size_t)gridDim.x * blockDim.x) {
        lo = fminf(lo, p[i]); hi = fmaxf(hi, p[i]);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(mm + 2 * n, enc(lo)); atomicMax(mm + 2 * n + 1, enc(hi)); }
}

__global__ void k_mm_init(unsigned* mm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

// pixels are staged through shared memory so that global reads are fully coalesced float4 streams;
// each thread then scans its pixel's K channels at stride K (odd K -> conflict-free)
__global__ void __launch_bounds__(K4_THREADS)
k_opm_extract(const float* __restrict__ logits, int K, size_t px_per_copy, int class_id, int mode,
              const unsigned* __restrict__ mm, float* __restrict__ class_out, float* __restrict__ max_out) {
    extern __shared__ __align__(16) float sm[];   // [K4_THREADS * K]
    const int n = blockIdx.y;
    const size_t px0 = (size_t)blockIdx.x * K4_THREADS;
    const size_t npx = min((size_t)K4_THREADS, px_per_copy - px0);
    const float* src = logits + ((size_t)n * px_per_copy + px0) * K;
    const size_t nel = npx * K;
    if ((((uintptr_t)src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(sm);
        for (size_t i = threadIdx.x; i < nel / 4; i += K4_THREADS) d4[i] = __ldg(s4 + i);
        for (size_t i = (nel / 4) * 4 + threadIdx.x; i < nel; i += K4_THREADS) sm[i] = __ldg(src + i);
    } else {
        for (size_t i = threadIdx.x; i < nel; i += K4_THREADS) sm[i] = __ldg(src + i);
    }
    __syncthreads();
    if (threadIdx.x >= npx) return;
    const float* v = sm + threadIdx.x * K;
    const size_t o = (size_t)n * px_per_copy + px0 + threadIdx.x;
    if (mode == ASR_OPM_ARGMAX) {
        int best = 0;
        float bv = v[0];
        for (int k = 1; k < K; ++k) { const float t = v[k]; if (t > bv) { bv = t; best = k; } }   // ties -> lowest index
        class_out[o] = (best == class_id) ? (float)class_id : 0.0f;
    } else if (mode == ASR_OPM_SLICE) {
        const float mn = dec(mm[2 * n]), mx = dec(mm[2 * n + 1]);
        const float den = (fsub(mx, mn) != 0.0f) ? fsub(mx, mn) : 1.0f;
        const float num = fmul(fsub(v[class_id], mn), fsub(1.0f, 0.0f));
        class_out[o] = fadd(0.0f, __fdiv_rn(num, den));
    } else {
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) if (k != class_id) m = fmaxf(m, v[k]);
        class_out[o] = v[class_id];
        max_out[o] = m;
    }
}

// ================================================================================================
// global min-max normalisation (load_SR_data) and threshold_image
// ================================================================================================
__global__ void k_minmax_reduce(const float* __restrict__ p, size_t n, unsigned* __restrict__ mm, size_t per_image) {
    // blockIdx.y selects the image when per_image != 0 (threshold: one max per image)
    const int b = blockIdx.y;
    const float* q = p + (size_t)b * per_image;
    const size_t cnt = per_image ? per_image : n;
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
        const float v = q[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { atomicMin(mm + 2 * b, enc(lo)); atomicMax(mm + 2 * b + 1, enc(hi)); }
}

__global__ void k_minmax_apply(const float* __restrict__ in, size_t n, const unsigned* __restrict__ mm, float new_min,
                               float new_max, float* __restrict__ out) {
    const float mn = dec(mm[0]), mx = dec(mm[1]);
    const float den = (fsub(mx, mn) != 0.0f) ? fsub(mx, mn) : 1.0f;
    const float scale = fsub(new_max, new_min);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fadd(new_min, __fdiv_rn(fmul(fsub(in[i], mn), scale), den));
}

__global__ void k_threshold(const float* __restrict__ x, size_t per_image, const unsigned* __restrict__ mm, float th_factor,
                            const float* __restrict__ th_mask, int th_value, int* __restrict__ out) {
    const int b = blockIdx.y;
    const size_t o = (size_t)b * per_image;
    if (th_mask) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x)
            out[o + i] = (x[o + i] >= th_mask[o + i]) ? th_value : 0;
    } else {
        const float th = fmul(dec(mm[2 * b + 1]), th_factor);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x)
            out[o + i] = (x[o + i] > th) ? th_value : 0;
    }
}

// ================================================================================================
// K5: max / mean back-projection.  out = reduce_k rotate(translate(resize(y_k), -shift_k), -angle_k)
// ================================================================================================
struct BackXf { float r0, r1, r2, r3, r4, r5, tx, ty; };   // rotate by -angle; translate offsets (+dx,+dy)

__device__ __forceinline__ float upsampled_at(const float* __restrict__ y, int h, int w, int H, int W, int my, int mx,
                                              float ys, float xs) {
    // value of tf.image.resize(y,(H,W)) at integer (my,mx); zero outside the canvas (fill of the translate op)
    if (my < 0 || my >= H || mx < 0 || mx >= W) return 0.0f;
    const float in_y = fsub(fmul(fadd((float)my, 0.5f), ys), 0.5f), in_x = fsub(fmul(fadd((float)mx, 0.5f), xs), 0.5f);
    const float fy = floorf(in_y), fx = floorf(in_x);
    const int y0 = max((int)fy, 0), y1 = min((int)ceilf(in_y), h - 1), x0 = max((int)fx, 0), x1 = min((int)ceilf(in_x), w - 1);
    const float yl = fsub(in_y, fy), xl = fsub(in_x, fx);
    const float tl = __ldg(y + y0 * w + x0), tr = __ldg(y + y0 * w + x1), bl = __ldg(y + y1 * w + x0), br = __ldg(y + y1 * w + x1);
    const float t = fadd(tl, fmul(fsub(tr, tl), xl)), b = fadd(bl, fmul(fsub(br, bl), xl));
    return fadd(t, fmul(fsub(b, t), yl));
}

__device__ __forceinline__ float translated_at(const float* __restrict__ y, int h, int w, int H, int W, int qy, int qx,
                                               float tx, float ty, float ys, float xs) {
    // value of tfa.image.translate(up, -shift) at integer (qy,qx); zero outside the canvas (fill of the rotate op)
    if (qy < 0 || qy >= H || qx < 0 || qx >= W) return 0.0f;
    const float ix = fadd((float)qx, tx), iy = fadd((float)qy, ty);
    const float fx = floorf(ix), fy = floorf(iy);
    const float wx0 = fsub(fadd(fx, 1.0f), ix), wx1 = fsub(ix, fx), wy0 = fsub(fadd(fy, 1.0f), iy), wy1 = fsub(iy, fy);
    const int x0 = (int)fx, y0 = (int)fy;
    return bilerp(upsampled_at(y, h, w, H, W, y0, x0, ys, xs), upsampled_at(y, h, w, H, W, y0, x0 + 1, ys, xs),
                  upsampled_at(y, h, w, H, W, y0 + 1, x0, ys, xs), upsampled_at(y, h, w, H, W, y0 + 1, x0 + 1, ys, xs),
                  wx0, wx1, wy0, wy1);
}

__global__ void __launch_bounds__(256)
k_backproject(const float* __restrict__ copies, const BackXf* __restrict__ xf, float* __restrict__ out, int mode, int N, int h,
              int w, int H, int W) {
    const int b = blockIdx.z;
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= W || Y >= H) return;
    const float ys = (float)h / (float)H, xs = (float)w / (float)W;
    const float Xf = (float)X, Yf = (float)Y;
    float accv = 0.0f;
    for (int k = 0; k < N; ++k) {
        const BackXf T = xf[(size_t)b * N + k];
        const float* y = copies + ((size_t)b * N + k) * h * w;
        const float ix = affine_coord(T.r0, Xf, T.r1, Yf, T.r2), iy = affine_coord(T.r3, Xf, T.r4, Yf, T.r5);
        const float fx = floorf(ix), fy = floorf(iy);
        const float wx0 = fsub(fadd(fx, 1.0f), ix), wx1 = fsub(ix, fx), wy0 = fsub(fadd(fy, 1.0f), iy), wy1 = fsub(iy, fy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float v = bilerp(translated_at(y, h, w, H, W, y0, x0, T.tx, T.ty, ys, xs),
                               translated_at(y, h, w, H, W, y0, x0 + 1, T.tx, T.ty, ys, xs),
                               translated_at(y, h, w, H, W, y0 + 1, x0, T.tx, T.ty, ys, xs),
                               translated_at(y, h, w, H, W, y0 + 1, x0 + 1, T.tx, T.ty, ys, xs), wx0, wx1, wy0, wy1);
        if (mode == ASR_BACKPROJECT_MAX) accv = (k == 0) ? v : fmaxf(accv, v);
        else accv = fadd(accv, v);
    }
    if (mode == ASR_BACKPROJECT_MEAN) accv = __fdiv_rn(accv, (float)N);
    out[((size_t)b * H + Y) * W + X] = accv;
}

// ---- tiled form (output == 4 x feature size) -----------------------------------------------------
// One CTA per 64x64 output tile, looping over the copies.  Per copy the three ops are evaluated stage by stage
// on the bounding box of the rotated tile instead of 4 x 4 x 4 nested taps per pixel:
//   T  = the x-lerp of tf.image.resize on the LR rows the box touches          (LR rows x box columns + 1)
//   U  = the y-lerp: the upsampled image on the box (+1 row/column for the translate taps)
//   Z  = tfa.image.translate of U with per-column / per-row literal tap tables (zero outside the canvas)
//   out = max / mean over copies of the packed-fp32 rotate gather from Z      (same gather as the solve's K2)
// Every value is produced by the same fp32 expression as the nested form, so the two are bit-identical.
constexpr int K5_T = 64;               // output tile edge
constexpr int K5_THREADS = 512;        // thread owns pixels (lane + 32c, warp + 16r), c<2, r<4
constexpr int K5_ZS = 96;              // Z stride and rows: 64*sqrt(2)+2+3 < 96, multiple of 32
constexpr int K5_BC = K5_ZS / 4;       // box cells per axis (24)
constexpr int K5_US = 100;             // U / T stride (97 columns)
constexpr int K5_CHUNK = 128;
struct __align__(16) K5Box { unsigned cst; int cbx0, cby0, ncxy, qx_lo, qy_lo, pad0, pad1; };
struct __align__(16) K5Lerp { int i0, i1; float l; int ok; };   // tf.image.resize of one HR column / row: LR taps, weight, on-canvas
constexpr size_t K5_SMEM = sizeof(float) * (K5_ZS * K5_ZS + (K5_ZS + 1) * K5_US + (K5_BC + 2) * K5_US) + sizeof(float2) * 2 * K5_ZS +
                           sizeof(K5Lerp) * 2 * K5_US + (sizeof(K5Box) + sizeof(BackXf)) * K5_CHUNK;

__global__ void __launch_bounds__(K5_THREADS, 2)
k_backproject_tiled(const float* __restrict__ copies, const BackXf* __restrict__ xf, float* __restrict__ out, int mode, int N, int h,
                    int w, int H, int W) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Zt = reinterpret_cast<float*>(smem_raw);                 // [K5_ZS][K5_ZS]
    float* Ut = Zt + K5_ZS * K5_ZS;                                 // [K5_ZS + 1][K5_US]
    float* Tt = Ut + (K5_ZS + 1) * K5_US;                           // [K5_BC + 2][K5_US]
    float2* colw = reinterpret_cast<float2*>(Tt + (K5_BC + 2) * K5_US);   // [K5_ZS]
    float2* roww = colw + K5_ZS;                                    // [K5_ZS]
    K5Lerp* lx = reinterpret_cast<K5Lerp*>(roww + K5_ZS);           // [K5_US] resize taps of the box's HR columns
    K5Lerp* ly = lx + K5_US;                                        // [K5_US] ... rows
    K5Box* boxes = reinterpret_cast<K5Box*>(ly + K5_US);            // [K5_CHUNK]
    BackXf* xfs = reinterpret_cast<BackXf*>(boxes + K5_CHUNK);      // [K5_CHUNK]

    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntx = (W + K5_T - 1) / K5_T;
    const int tx0 = (blockIdx.x % ntx) * K5_T, ty0 = (blockIdx.x / ntx) * K5_T;
    const BackXf* xfb = xf + (size_t)b * N;
    const float X0f = (float)(tx0 + lane), X1f = (float)(tx0 + lane + 32);
    f32x2 accp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) accp[r] = pk(0.0f, 0.0f);
    const f32x2 magic2 = pk(kMagic, kMagic), one2 = pk(1.0f, 1.0f);

    for (int k0 = 0; k0 < N; k0 += K5_CHUNK) {
        const int nc = min(K5_CHUNK, N - k0);
        __syncthreads();
        // bounding box of R(tile) per copy: each rounded op of the coordinate is monotone, the corners bound every tap
        if (tid < nc) {
            const BackXf T = xfb[k0 + tid];
            float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
            for (int cnr = 0; cnr < 4; ++cnr) {
                const float X = (float)(tx0 + ((cnr & 1) ? K5_T - 1 : 0)), Y = (float)(ty0 + ((cnr & 2) ? K5_T - 1 : 0));
                const float cix = affine_coord(T.r0, X, T.r1, Y, T.r2), ciy = affine_coord(T.r3, X, T.r4, Y, T.r5);
                xmin = fminf(xmin, cix); xmax = fmaxf(xmax, cix); ymin = fminf(ymin, ciy); ymax = fmaxf(ymax, ciy);
            }
            const int qx0 = (int)floorf(xmin), qx1 = (int)floorf(xmax) + 1, qy0 = (int)floorf(ymin), qy1 = (int)floorf(ymax) + 1;
            const int sx = (int)floorf(T.tx), sy = (int)floorf(T.ty);
            K5Box bx;
            bx.cbx0 = (qx0 + sx) >> 2;
            bx.cby0 = (qy0 + sy) >> 2;
            const int cbx1 = (qx1 + sx) >> 2, cby1 = (qy1 + sy) >> 2;
            const int ncx = cbx1 - bx.cbx0 + 1, ncy = cby1 - bx.cby0 + 1;
            bx.qx_lo = 4 * bx.cbx0 - sx;
            bx.qy_lo = 4 * bx.cby0 - sy;
            // skip: no base tap of the box lies on the canvas, Z == 0
            const int skip = (cbx1 < 0 || bx.cbx0 >= w || cby1 < 0 || bx.cby0 >= h);
            if (!skip && (ncx > K5_BC || ncy > K5_BC)) __trap();
            bx.ncxy = (ncx & 0xff) | ((ncy & 0xff) << 8) | (skip << 16);
            bx.cst = (0u - (unsigned)(kMagicBits + bx.qy_lo) * K5_ZS - (unsigned)(kMagicBits + bx.qx_lo)) << 2;
            bx.pad0 = bx.pad1 = 0;
            boxes[tid] = bx;
            xfs[tid] = T;
        }
        __syncthreads();

        for (int kc = 0; kc < nc; ++kc) {
            const K5Box bx = boxes[kc];
            const BackXf T = xfs[kc];
            const bool live = !(bx.ncxy >> 16);
            f32x2 v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = pk(0.0f, 0.0f);
            if (live) {
                const int ncx = bx.ncxy & 0xff, ncy = (bx.ncxy >> 8) & 0xff;
                const int nux = 4 * ncx + 1, nuy = 4 * ncy + 1;           // U extent: HR positions 4*cb0 ... 4*cb0 + 4*nc
                const float* yk = copies + ((size_t)b * N + k0 + kc) * h * w;
                // ---- per-column / per-row tables: translate taps (source validity and the canvas of Z folded in)
                //      and the resize taps of every HR column / row of the box --------------------------------
                if (tid < 2 * K5_ZS) {
                    const bool col = tid < K5_ZS;
                    const int e = col ? tid : tid - K5_ZS;
                    const int q = (col ? bx.qx_lo : bx.qy_lo) + e, lim = col ? W : H;
                    const float t = col ? T.tx : T.ty;
                    float2 wv = make_float2(0.0f, 0.0f);
                    if (q >= 0 && q < lim) wv = warp_taps(q, t, (int)floorf(t), lim, ASR_INTERP_BILINEAR);
                    (col ? colw : roww)[e] = wv;
                } else if (tid < 2 * K5_ZS + 2 * K5_US) {
                    const int e2 = tid - 2 * K5_ZS;
                    const bool col = e2 < K5_US;
                    const int e = col ? e2 : e2 - K5_US;
                    const int m = 4 * (col ? bx.cbx0 : bx.cby0) + e, lim = col ? W : H, n = col ? w : h;
                    const float sc = (float)n / (float)lim;
                    const float in = fsub(fmul(fadd((float)m, 0.5f), sc), 0.5f);
                    const float f = floorf(in);
                    K5Lerp L;
                    L.i0 = max((int)f, 0); L.i1 = min((int)ceilf(in), n - 1); L.l = fsub(in, f);
                    L.ok = (m >= 0 && m < lim && e < (col ? nux : nuy)) ? 1 : 0;
                    if (!L.ok) { L.i0 = 0; L.i1 = 0; }
                    (col ? lx : ly)[e] = L;
                }
                __syncthreads();
                // ---- T = x-lerp of the LR rows the box touches ------------------------------------------------
                for (int ry = warp; ry < ncy + 2; ry += K5_THREADS / 32) {
                    const int lr = bx.cby0 - 1 + ry;
                    const bool row_ok = lr >= 0 && lr < h;
                    const float* yr = yk + (row_ok ? lr : 0) * w;
                    for (int ex = lane; ex < nux; ex += 32) {
                        const K5Lerp L = lx[ex];
                        float tv = 0.0f;
                        if (row_ok && L.ok) {
                            const float tl = __ldg(yr + L.i0), tr = __ldg(yr + L.i1);
                            tv = fadd(tl, fmul(fsub(tr, tl), L.l));
                        }
                        Tt[ry * K5_US + ex] = tv;
                    }
                }
                __syncthreads();
                // ---- U = upsampled image on the box ---------------------------------------------------------
                for (int ey = warp; ey < nuy; ey += K5_THREADS / 32) {
                    const K5Lerp L = ly[ey];
                    const float* t0 = Tt + (L.i0 - bx.cby0 + 1) * K5_US;
                    const float* t1 = Tt + (L.i1 - bx.cby0 + 1) * K5_US;
                    for (int ex = lane; ex < nux; ex += 32) {
                        float uv = 0.0f;
                        if (L.ok && lx[ex].ok) {
                            const float t = t0[ex], bb = t1[ex];
                            uv = fadd(t, fmul(fsub(bb, t), L.l));
                        }
                        Ut[ey * K5_US + ex] = uv;
                    }
                }
                __syncthreads();
                // ---- Z = translate(U) on the box ------------------------------------------------------------
                const int nzx = 4 * ncx;
                for (int ey = warp; ey < 4 * ncy; ey += K5_THREADS / 32) {
                    const float2 wr = roww[ey];
                    for (int ex = lane; ex < nzx; ex += 32) {
                        const float2 wc = colw[ex];
                        const float* u0 = Ut + ey * K5_US + ex;
                        Zt[ey * K5_ZS + ex] = bilerp(u0[0], u0[1], u0[K5_US], u0[K5_US + 1], wc.x, wc.y, wr.x, wr.y);
                    }
                }
                __syncthreads();
                // ---- rotate gather (packed fp32: the two lanes are columns lane, lane+32) ----------------------
                const f32x2 r2p = pk(T.r2, T.r2), r5p = pk(T.r5, T.r5);
                const f32x2 axp = pk(fmul(T.r0, X0f), fmul(T.r0, X1f)), ayp = pk(fmul(T.r3, X0f), fmul(T.r3, X1f));
                unsigned cst = bx.cst + smem_u32(Zt);
                asm volatile("" : "+r"(cst));
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float Yf = (float)(ty0 + warp + 16 * r);
                    const float bxr = fmul(T.r1, Yf), byr = fmul(T.r4, Yf);
                    const f32x2 ix = add2(add2(axp, pk(bxr, bxr)), r2p);
                    const f32x2 iy = add2(add2(ayp, pk(byr, byr)), r5p);
                    const f32x2 txx = add2_rd(ix, magic2), tyy = add2_rd(iy, magic2);
                    const f32x2 fxf = sub2(txx, magic2), fyf = sub2(tyy, magic2);
                    // (x_ceil - x) == 1 - (x - x_floor) bit for bit unless x in (-1,0), where that weight only multiplies
                    // the tap x_floor = -1, which lies outside the canvas and is an exact zero of Z
                    const f32x2 wx1 = sub2(ix, fxf), wx0 = sub2(one2, wx1);
                    const f32x2 wy1 = sub2(iy, fyf), wy0 = sub2(one2, wy1);
                    const unsigned ta = tap_offset<K5_ZS>(__float_as_uint(pk_lo(txx)), __float_as_uint(pk_lo(tyy)), cst);
                    const unsigned tb = tap_offset<K5_ZS>(__float_as_uint(pk_hi(txx)), __float_as_uint(pk_hi(tyy)), cst);
                    v[r] = bilerp2(pk(lds_tap<0>(ta), lds_tap<0>(tb)), pk(lds_tap<4>(ta), lds_tap<4>(tb)),
                                   pk(lds_tap<4 * K5_ZS>(ta), lds_tap<4 * K5_ZS>(tb)),
                                   pk(lds_tap<4 * K5_ZS + 4>(ta), lds_tap<4 * K5_ZS + 4>(tb)), wx0, wx1, wy0, wy1);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (mode == ASR_BACKPROJECT_MAX) {
                    accp[r] = (k0 + kc == 0) ? v[r] : pk(fmaxf(pk_lo(accp[r]), pk_lo(v[r])), fmaxf(pk_hi(accp[r]), pk_hi(v[r])));
                } else {
                    accp[r] = add2(accp[r], v[r]);
                }
            }
            // the next copy's first write to Z comes after two more barriers; T and U are dead here
        }
    }
    float* ob = out + (size_t)b * H * W;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int X = tx0 + lane + 32 * c, Y = ty0 + warp + 16 * r;
            if (X >= W || Y >= H) continue;
            float a = c ? pk_hi(accp[r]) : pk_lo(accp[r]);
            if (mode == ASR_BACKPROJECT_MEAN) a = __fdiv_rn(a, (float)N);
            ob[(size_t)Y * W + X] = a;
        }
    }
}

// ================================================================================================
// single_class_IOU counts (utils.py:180-204): per image inter/union for `class_id` and, with
// include_bg, for class 0 after relabelling every other ground-truth class to background
// ================================================================================================
__global__ void k_iou_counts(const int* __restrict__ yt, const int* __restrict__ yp, size_t n, int class_id, int include_bg,
                             unsigned long long* __restrict__ counts) {
    const int b = blockIdx.y;
    const int* t = yt + (size_t)b * n;
    const int* p = yp + (size_t)b * n;
    unsigned ic = 0, uc = 0, ib = 0, ub = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int tv = t[i];
        const int pv = p[i];
        if (include_bg && tv != class_id) tv = 0;
        const bool tc = tv == class_id, pc = pv == class_id, tb = tv == 0, pbk = pv == 0;
        ic += tc && pc; uc += tc || pc; ib += tb && pbk; ub += tb || pbk;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ic += __shfl_xor_sync(0xffffffffu, ic, o); uc += __shfl_xor_sync(0xffffffffu, uc, o);
        ib += __shfl_xor_sync(0xffffffffu, ib, o); ub += __shfl_xor_sync(0xffffffffu, ub, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(counts + 4 * b + 0, (unsigned long long)ic); atomicAdd(counts + 4 * b + 1, (unsigned long long)uc);
        atomicAdd(counts + 4 * b + 2, (unsigned long long)ib); atomicAdd(counts + 4 * b + 3, (unsigned long long)ub);
    }
}

// ================================================================================================
// measurement hook: L2 read bandwidth (bench.py's roofline.l2; SURVEY 8d asks for a measured L2 peak)
// ================================================================================================
// Every thread streams 16-byte words of a buffer that fits the L2 (ld.global.cg: L1 is bypassed), `passes` times over;
// the first pass warms the L2, the caller times the launch with CUDA events and divides bytes * passes by it.
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n16, int passes, unsigned* __restrict__ sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {   // four independent loads in flight per thread
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
        }
        for (; i < n16; i += stride) { const uint4 a = __ldcg(buf + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
    }
    if (acc == 0x9e3779b9u) *sink = acc;   // keeps the loads alive; practically never true
}

}  // namespace asr

using namespace asr;

extern "C" int asr_l2_read_probe(const void* d_buf, size_t bytes, int passes, void* d_sink, void* stream) {
    if (!d_buf || !d_sink) return fail(ASR_ENULL, "null argument");
    if (bytes < 16 || passes <= 0 || !aligned16(d_buf)) return fail(ASR_EINVAL, "need bytes >= 16, passes > 0 and a 16-byte aligned buffer");
    int dev = 0, n_sm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    ASR_LAUNCH(k_l2_read, n_sm * 8, 256, 0, static_cast<cudaStream_t>(stream), static_cast<const uint4*>(d_buf), bytes / 16, passes,
               static_cast<unsigned*>(d_sink));
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

extern "C" int asr_iou_counts(const int32_t* d_true, const int32_t* d_pred, int B, int64_t n, int class_id, int include_bg,
                              unsigned long long* d_counts, void* stream) {
    if (!d_true || !d_pred || !d_counts) return fail(ASR_ENULL, "null argument");
    if (B <= 0 || n <= 0 || B > 65535) return fail(ASR_EINVAL, "need 0 < B <= 65535 and n > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ASR_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * 4 * B, st));
    ASR_LAUNCH(k_iou_counts, dim3(16, B), 256, 0, st, d_true, d_pred, (size_t)n, class_id, include_bg, d_counts);
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

static size_t warp_xf_bytes(int N) { return (sizeof(WarpXf) * (size_t)N + 255) / 256 * 256; }

extern "C" int asr_warp_affine_workspace_bytes(int N, int H, int W, int C, size_t* bytes) {
    if (!bytes) return fail(ASR_ENULL, "bytes is NULL");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > K3_CMAX) return fail(ASR_EINVAL, "need N,H,W > 0 and 1 <= C <= %d", K3_CMAX);
    *bytes = warp_xf_bytes(N) + sizeof(float4) * (size_t)H * W;   // per-copy transform table + the image padded to 4 channels
    return ASR_OK;
}

static int warp_check(const float* d_image, const float* h_angles, const float* h_shifts, int N, int H, int W, int C, int interp,
                      const float* d_out) {
    if (!d_image || !h_angles || !h_shifts || !d_out) return fail(ASR_ENULL, "null argument");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || C > K3_CMAX) return fail(ASR_EINVAL, "need N,H,W > 0 and 1 <= C <= %d", K3_CMAX);
    if (N > 65535) return fail(ASR_EINVAL, "N must be <= 65535");
    if (H > (1 << 20) || W > (1 << 20)) return fail(ASR_EINVAL, "image too large for the fp32 floor trick");
    if (interp != ASR_INTERP_NEAREST && interp != ASR_INTERP_BILINEAR) return fail(ASR_EINVAL, "unknown interpolation %d", interp);
    if (!aligned16(d_image) || !aligned16(d_out)) return fail(ASR_EINVAL, "device pointers must be 16-byte aligned");
    return ASR_OK;
}

extern "C" int asr_warp_affine_ws(const float* d_image, const float* h_angles, const float* h_shifts, int N, int H, int W,
                                  int C, int interp, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (int e = warp_check(d_image, h_angles, h_shifts, N, H, W, C, interp, d_out)) return e;
    if (!d_workspace) return fail(ASR_ENULL, "null workspace");
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255u) return fail(ASR_EINVAL, "workspace must be 256-byte aligned");
    const size_t npx = (size_t)H * W, xf_bytes = warp_xf_bytes(N);
    if (workspace_bytes < xf_bytes + sizeof(float4) * npx) return fail(ASR_EWORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, xf_bytes + sizeof(float4) * npx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<WarpXf> xf(N);
    for (int k = 0; k < N; ++k) {
        float r[8];
        rotate_matrix(h_angles[k], H, W, r);
        xf[k] = WarpXf{r[0], r[1], r[2], r[3], r[4], r[5], -h_shifts[2 * k], -h_shifts[2 * k + 1]};
    }
    unsigned char* scratch = static_cast<unsigned char*>(d_workspace);
    WarpXf* d_xf = reinterpret_cast<WarpXf*>(scratch);
    float4* d_pad = reinterpret_cast<float4*>(scratch + xf_bytes);
    ASR_CUDA_TRY(cudaMemcpyAsync(d_xf, xf.data(), sizeof(WarpXf) * N, cudaMemcpyHostToDevice, st));
    ASR_LAUNCH(k_pad_channels, 296, 256, 0, st, d_image, d_pad, npx, C);
    const int tiles = ((W + K3_T - 1) / K3_T) * ((H + K3_T - 1) / K3_T);
    switch (C) {
    case 1: ASR_LAUNCH(k_warp_affine<1>, dim3(tiles, N), K3_THREADS, 0, st, d_pad, d_xf, d_out, H, W, interp); break;
    case 2: ASR_LAUNCH(k_warp_affine<2>, dim3(tiles, N), K3_THREADS, 0, st, d_pad, d_xf, d_out, H, W, interp); break;
    case 3: ASR_LAUNCH(k_warp_affine<3>, dim3(tiles, N), K3_THREADS, 0, st, d_pad, d_xf, d_out, H, W, interp); break;
    default: ASR_LAUNCH(k_warp_affine<4>, dim3(tiles, N), K3_THREADS, 0, st, d_pad, d_xf, d_out, H, W, interp); break;
    }
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

// convenience form: the scratch is allocated and released in stream order inside the call
extern "C" int asr_warp_affine(const float* d_image, const float* h_angles, const float* h_shifts, int N, int H, int W,
                               int C, int interp, float* d_out, void* stream) {
    if (int e = warp_check(d_image, h_angles, h_shifts, N, H, W, C, interp, d_out)) return e;
    size_t need = 0;
    if (int e = asr_warp_affine_workspace_bytes(N, H, W, C, &need)) return e;
    AsyncScratch guard;   // released in stream order on every return path
    ASR_CUDA_TRY(guard.alloc(need, static_cast<cudaStream_t>(stream)));
    return asr_warp_affine_ws(d_image, h_angles, h_shifts, N, H, W, C, interp, d_out, guard.p, need, stream);
}

extern "C" int asr_opm_extract(const float* d_logits, int N, int h, int w, int K, int class_id, int mode,
                               float* d_class_out, float* d_max_out, void* d_workspace, void* stream) {
    if (!d_logits || !d_class_out) return fail(ASR_ENULL, "null argument");
    if (N <= 0 || h <= 0 || w <= 0 || K <= 0 || K > K4_KMAX) return fail(ASR_EINVAL, "need N,h,w > 0 and 1 <= K <= %d", K4_KMAX);
    if (class_id < 0 || class_id >= K) return fail(ASR_EINVAL, "class_id %d outside [0,%d)", class_id, K);
    if (mode < ASR_OPM_ARGMAX || mode > ASR_OPM_SLICE_MAX) return fail(ASR_EINVAL, "unknown OPM mode %d", mode);
    if (mode == ASR_OPM_SLICE_MAX && !d_max_out) return fail(ASR_ENULL, "slice_max needs d_max_out");
    if (mode == ASR_OPM_SLICE && !d_workspace) return fail(ASR_ENULL, "slice needs a 2*N float workspace");
    if (N > 65535) return fail(ASR_EINVAL, "N must be <= 65535");
    if (!aligned16(d_logits) || !aligned16(d_class_out) || !aligned16(d_max_out) || !aligned16(d_workspace))
        return fail(ASR_EINVAL, "device pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t px = (size_t)h * w;
    unsigned* mm = static_cast<unsigned*>(d_workspace);
    if (mode == ASR_OPM_SLICE) {
        ASR_LAUNCH(k_mm_init, (N + 127) / 128, 128, 0, st, mm, N);
        ASR_LAUNCH(k_copy_minmax, dim3(32, N), 256, 0, st, d_logits, px * K, mm);
    }
    static unsigned long long attr = 0;
    if (first_use_on_device(&attr))
        ASR_CUDA_TRY(cudaFuncSetAttribute(k_opm_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * K4_THREADS * K4_KMAX)));
    const unsigned blocks = (unsigned)((px + K4_THREADS - 1) / K4_THREADS);
    ASR_LAUNCH(k_opm_extract, dim3(blocks, N), K4_THREADS, sizeof(float) * K4_THREADS * K, st, d_logits, K, px, class_id, mode, mm,
                                                                                      d_class_out, d_max_out);
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

extern "C" int asr_minmax_normalize(const float* d_in, int64_t n, float new_min, float new_max, float* d_out,
                                    void* d_workspace, void* stream) {
    if (!d_in || !d_out || !d_workspace) return fail(ASR_ENULL, "null argument");
    if (n <= 0) return fail(ASR_EINVAL, "n must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* mm = static_cast<unsigned*>(d_workspace);
    ASR_LAUNCH(k_mm_init, 1, 32, 0, st, mm, 1);
    ASR_LAUNCH(k_minmax_reduce, dim3(296, 1), 256, 0, st, d_in, (size_t)n, mm, 0);
    ASR_LAUNCH(k_minmax_apply, 296 * 2, 256, 0, st, d_in, (size_t)n, mm, new_min, new_max, d_out);
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

extern "C" int asr_threshold(const float* d_x, int B, int64_t n, int32_t th_value, float th_factor, const float* d_th_mask,
                             int32_t* d_out, void* d_workspace, void* stream) {
    if (!d_x || !d_out) return fail(ASR_ENULL, "null argument");
    if (!d_th_mask && !d_workspace) return fail(ASR_ENULL, "th_factor path needs a 2*B float workspace");
    if (B <= 0 || n <= 0 || B > 65535) return fail(ASR_EINVAL, "need 0 < B <= 65535 and n > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* mm = static_cast<unsigned*>(d_workspace);
    if (!d_th_mask) {
        ASR_LAUNCH(k_mm_init, (B + 127) / 128, 128, 0, st, mm, B);
        ASR_LAUNCH(k_minmax_reduce, dim3(16, B), 256, 0, st, d_x, 0, mm, (size_t)n);
    }
    ASR_LAUNCH(k_threshold, dim3(16, B), 256, 0, st, d_x, (size_t)n, mm, th_factor, d_th_mask, th_value, d_out);
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

static int backproject_check(int mode, const float* d_copies, const float* h_angles, const float* h_shifts, int B, int N, int h, int w,
                             int H, int W, const float* d_out) {
    if (!d_copies || !h_angles || !h_shifts || !d_out) return fail(ASR_ENULL, "null argument");
    if (mode != ASR_BACKPROJECT_MAX && mode != ASR_BACKPROJECT_MEAN) return fail(ASR_EINVAL, "mode must be max or mean");
    if (B <= 0 || N <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || B > 65535) return fail(ASR_EINVAL, "bad shape");
    if (!aligned16(d_copies) || !aligned16(d_out)) return fail(ASR_EINVAL, "device pointers must be 16-byte aligned");
    return ASR_OK;
}

extern "C" int asr_backproject_workspace_bytes(int B, int N, size_t* bytes) {
    if (!bytes) return fail(ASR_ENULL, "bytes is NULL");
    if (B <= 0 || N <= 0) return fail(ASR_EINVAL, "bad shape");
    *bytes = (sizeof(BackXf) * (size_t)B * N + 255) / 256 * 256;
    return ASR_OK;
}

extern "C" int asr_backproject_batched_ws(int mode, const float* d_copies, const float* h_angles, const float* h_shifts, int B,
                                          int N, int h, int w, int H, int W, float* d_out, void* d_workspace, size_t workspace_bytes,
                                          void* stream) {
    if (int e = backproject_check(mode, d_copies, h_angles, h_shifts, B, N, h, w, H, W, d_out)) return e;
    if (!d_workspace) return fail(ASR_ENULL, "null workspace");
    if (!aligned16(d_workspace)) return fail(ASR_EINVAL, "workspace must be 16-byte aligned");
    if (workspace_bytes < sizeof(BackXf) * (size_t)B * N) return fail(ASR_EWORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, sizeof(BackXf) * (size_t)B * N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<BackXf> xf((size_t)B * N);
    for (size_t i = 0; i < xf.size(); ++i) {
        float r[8];
        rotate_matrix(-h_angles[i], H, W, r);                       // tfa.image.rotate(..., -angles)
        // tfa.image.translate(..., -shifts): transform offsets -(-dx), -(-dy)
        xf[i] = BackXf{r[0], r[1], r[2], r[3], r[4], r[5], -(-h_shifts[2 * i]), -(-h_shifts[2 * i + 1])};
    }
    BackXf* d_xf = static_cast<BackXf*>(d_workspace);
    ASR_CUDA_TRY(cudaMemcpyAsync(d_xf, xf.data(), sizeof(BackXf) * xf.size(), cudaMemcpyHostToDevice, st));
    if (H == 4 * h && W == 4 * w) {
        static unsigned long long attr = 0;
        if (first_use_on_device(&attr))
            ASR_CUDA_TRY(cudaFuncSetAttribute(k_backproject_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K5_SMEM));
        const int tiles = ((W + K5_T - 1) / K5_T) * ((H + K5_T - 1) / K5_T);
        ASR_LAUNCH(k_backproject_tiled, dim3(tiles, B), K5_THREADS, K5_SMEM, st, d_copies, d_xf, d_out, mode, N, h, w, H, W);
    } else {
        ASR_LAUNCH(k_backproject, dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, st, d_copies, d_xf, d_out, mode, N, h, w, H, W);
    }
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

// convenience form: the per-copy transform table is allocated and released in stream order inside the call
extern "C" int asr_backproject_batched(int mode, const float* d_copies, const float* h_angles, const float* h_shifts, int B,
                                       int N, int h, int w, int H, int W, float* d_out, void* stream) {
    if (int e = backproject_check(mode, d_copies, h_angles, h_shifts, B, N, h, w, H, W, d_out)) return e;
    size_t need = 0;
    if (int e = asr_backproject_workspace_bytes(B, N, &need)) return e;
    AsyncScratch scratch;   // released in stream order on every return path
    ASR_CUDA_TRY(scratch.alloc(need, static_cast<cudaStream_t>(stream)));
    return asr_backproject_batched_ws(mode, d_copies, h_angles, h_shifts, B, N, h, w, H, W, d_out, scratch.p, need, stream);
}
