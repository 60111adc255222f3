// asr_solve.cu -- the multi-frame super-resolution inverse solve on B200 (sm_100a).
//
// Replaces the TensorFlow op sequence behind Superresolution.augmented_superresolution
// (superresolution_scripts/superresolution.py:102-137), i.e. per iteration
//   loss_function (:44-100): tile -> tfa.rotate -> tfa.translate -> tf.image.resize -> sum (D-y)^2, TV, L2, L1
//   tape.gradient (:126-133): ResizeBilinearGrad -> warp-grad(translate) -> warp-grad(rotate) -> sum over copies
//   optimizer.apply_gradients (:134-135, optimizer.py:21-41)
// with two kernels per iteration for a whole batch of images:
//   k_forward_residual   r_k = D T_k R_k x - y_k              (one CTA per 16x16 LR tile per copy, TMA-staged x box)
//   k_gradient_update    x' = opt(x, sum_k W_k^grad r_k + reg) (one CTA per 64x64 HR tile loops over the copies:
//                                                               8 gather warps + 4 TMA-fed fill warps)
// plus k_tap_tables and k_forward_tables once per solve.  The gathers run on packed fp32 (FADD2/FMUL2/FFMA2, two pixels
// per instruction).  Output/feature ratios other than 4 take the literal kg_* kernels further down.
// Both are gather-form (no atomics) and bit-reproduce the un-fused fp32 evaluation order of the
// TensorFlow ops (see asr_common.cuh).  DESIGN.md derives the restructurings used here and why
// each is bit-identical to the literal two-pass evaluation.
#include <cuda.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <vector>

#include "asr_common.cuh"

namespace asr {

// ================================================================================================
// device-side per-image parameters
// ================================================================================================
struct ImgParams {
    float two_ldf;      // fl(2 * lambda_df): _SquaredDifferenceGrad scalar
    float lambda_tv, lambda_l2, lambda_l1;
    float omb1, omb2;   // fl(1 - beta_1), fl(1 - beta_2)
    float beta_2;       // adamax
    float epsilon;
    float momentum;
    int optimizer, amsgrad, nesterov;
    int num_iter;
    int n_kept;
    int use_btv;        // bilateral TV (superresolution.py:8-23) instead of tf.image.image_gradients TV
    int stack;          // which LR stack of `copies` this solve reads (== image index unless a sweep shares stacks)
    float btv_w[5];     // 0.6^n, n = |h|+|v|
    float btv_lw[5];    // fl(lambda_tv * 0.6^n)
    int pad1, pad2;
};
// per (iteration, image): x = learning rate of that step, y = optimizer-specific step scale
// (Adam: lr*sqrt(1-b2^t)/(1-b1^t); Adamax: lr/(1-b1^t))
typedef float2 Sched;

constexpr int LOG2S = 2;  // HR/LR scale 4 (H == 4h): the 2x2 box of tf.image.resize sits at phases 1,2

// ================================================================================================
// K0: x0 = tf.image.resize(copies[0], (H,W))  (superresolution.py:112-113)
// ================================================================================================
__global__ void k_init_upsample(const float* __restrict__ copies, const ImgParams* __restrict__ ip, float* __restrict__ x0, int N,
                                int h, int w, int H, int W) {
    const int b = blockIdx.z;
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= W || Y >= H) return;
    const float* img = copies + (size_t)ip[b].stack * N * h * w;  // copy 0 is the un-augmented one
    const float ys = (float)h / (float)H, xs = (float)w / (float)W;
    const float in_y = fsub(fmul(fadd((float)Y, 0.5f), ys), 0.5f);
    const float in_x = fsub(fmul(fadd((float)X, 0.5f), xs), 0.5f);
    const float fy = floorf(in_y), fx = floorf(in_x);
    const int y0 = max((int)fy, 0), y1 = min((int)ceilf(in_y), h - 1);
    const int x0i = max((int)fx, 0), x1 = min((int)ceilf(in_x), w - 1);
    const float yl = fsub(in_y, fy), xl = fsub(in_x, fx);
    const float tl = img[y0 * w + x0i], tr = img[y0 * w + x1], bl = img[y1 * w + x0i], br = img[y1 * w + x1];
    const float t = fadd(tl, fmul(fsub(tr, tl), xl));
    const float bb = fadd(bl, fmul(fsub(br, bl), xl));
    x0[((size_t)b * H + Y) * W + X] = fadd(t, fmul(fsub(bb, t), yl));
}

// ================================================================================================
// K1: forward residual
// ================================================================================================
// For copy k and LR cell (i,j):  r = resize(translate(rotate(x)))[i,j] - y_k[i,j].
// The resize reads only z at rows {4i+1,4i+2} x cols {4j+1,4j+2}; each z is a 2x2 stencil of the
// rotated image p on integer positions, so one cell needs p on a 3x3 patch whose origin is
// (4j+1+floor(-dx), 4i+1+floor(-dy)).  A CTA (16x16 cells, one copy) pulls the x region those patches can
// touch into shared memory with ONE TMA tensor load (cp.async.bulk.tensor; out-of-image elements arrive as
// zeros, which is exactly the op's fill), evaluates the 48x48 needed p values with the op's arithmetic, and
// a second phase combines them per cell with per-column/row tap tables that carry the literal translate
// weights, the zero fill of p outside the canvas, and the rounding case floor(fl(Z-dx)) == Z+floor(-dx)+1.
// Everything that depends on the transforms only -- the box origin of every (tile, copy) and the tap tables of
// every copy -- is computed ONCE per solve by k_forward_tables, so the per-iteration CTA starts with one 8-byte
// descriptor load and the TMA issue (round 1 recomputed boxes and tables in every CTA of every iteration:
// half of the kernel's instructions).
constexpr int K1_TJ = 16;              // LR tile: 16 cells wide ...
constexpr int K1_TI = 16;              // ... 16 cells tall = 256 cells, one per thread in the second phase.  (Round 1's 16x12 tile gave every
                                       // gather thread 9 rows = 4 packed pairs + one scalar row that cost almost a pair, and tiled a 128-row map
                                       // with 11 x 12 = 132 rows; 16 rows = 6 packed pairs per thread and no ragged tile: K1 23.9 -> see DESIGN.md)
constexpr int K1_THREADS = 256;
constexpr int K1_GATHER = 192;         // threads of the gather phase: 48 p-columns x 4 row groups of 12 rows
constexpr int K1_PC = 3 * K1_TJ;       // needed p columns (48) and rows (48) per tile
constexpr int K1_PR = 3 * K1_TI;
constexpr int K1_PBS = K1_PC;          // p buffer stride 48: 3 rows = 144 = 16 (mod 32) floats, so the two cell rows a warp reads in
                                       // the second phase land on complementary bank sets (stride 49 collided on one bank: 2 wavefronts per load)
constexpr int K1_XS = 96;              // TMA box width (floats): 62*sqrt(2)+2+3 < 96, multiple of 32 (a pitch of 80 = 16 mod 32 saves TMA
                                       // bytes for small rotations but adds conflicts across source rows: measured 29.5 vs 28.0 us)
#ifndef ASR_K1_XR_SMALL
#define ASR_K1_XR_SMALL 74
#endif
constexpr int K1_XR_SMALL = ASR_K1_XR_SMALL;   // TMA box height when 62|sin|+62|cos|+3.05 <= 74 for every copy (|angle| <= 0.15 rad, the reference's
                                       // angle_max): 37.6 KB per CTA = 6 CTAs/SM, the register limit.  76 rows (|angle| <~ 0.19) was 5 CTAs/SM: 23.9 -> 22.8 us
constexpr int K1_XR_BIG = 92;          // ... for any rotation: 62*sqrt(2)+3 < 92: 44.6 KB per CTA = 5 CTAs/SM
constexpr int K1_SPAN_X = 4 * (K1_TJ - 1) + 2, K1_SPAN_Y = 4 * (K1_TI - 1) + 2;   // last needed p position = first + SPAN
constexpr int K1_EMPTY = INT_MIN;      // BoxDesc.by0 of a (tile, copy) whose rotated image is all zero
template <int XR>
constexpr size_t k1_smem() {
    return sizeof(float) * K1_XS * XR + sizeof(float) * (K1_PR * K1_PBS) + 16;
}
// per kept copy, device-side: rotate coefficients, floor of the translate offsets, index of the LR map y_k in `copies`
struct __align__(16) FwdCopy { float r0, r1, r2, r3, r4, r5; int sx, sy; int ysrc, pad0, pad1, pad2; };
typedef int2 BoxDesc;                  // x = TMA box start column (multiple of 4), y = start row or K1_EMPTY

// translate stencil weights of z-column Z on the window (Z+s, Z+s+1), validity of p folded in
__device__ __forceinline__ float2 translate_taps(int Z, float t, int s, int limit) {
    const float iz = fadd((float)Z, t);  // (1*Z + 0*Zy) + t
    const float f = floorf(iz);
    const float w0 = fsub(fadd(f, 1.0f), iz), w1 = fsub(iz, f);
    const int base = Z + s;
    float wa, wb;
    if ((int)f == base) { wa = w0; wb = w1; } else { wa = 0.0f; wb = w0; }  // else: floor == base+1, w1 == 0
    if (base < 0 || base >= limit) wa = 0.0f;
    if (base + 1 < 0 || base + 1 >= limit) wb = 0.0f;
    return make_float2(wa, wb);
}

// Once per solve: for every kept copy the forward tap tables (one float4 per LR column and per LR row), the
// compact per-copy record, and for every (LR tile, copy) the origin of the x box its taps can touch.
// blockIdx.x < tile_blocks: one thread per tile of the copy; then one thread per LR column / row.
__global__ void k_forward_tables(const FwdXf* __restrict__ fwd, const int* __restrict__ src_idx, const ImgParams* __restrict__ ip,
                                 FwdCopy* __restrict__ fcp, float4* __restrict__ fcolw, float4* __restrict__ froww,
                                 BoxDesc* __restrict__ boxd, int N, int h, int w, int H, int W, int ntj, int nti, int box_rows) {
    const int ks = blockIdx.y, b = blockIdx.z;
    if (ks >= ip[b].n_kept) return;
    const size_t slot = (size_t)b * N + ks;
    const FwdXf T = fwd[slot];
    const int sx = (int)floorf(T.tx), sy = (int)floorf(T.ty);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int t1 = ntj * nti;
    if (e == 0) fcp[slot] = FwdCopy{T.r0, T.r1, T.r2, T.r3, T.r4, T.r5, sx, sy, ip[b].stack * N + src_idx[slot], 0, 0, 0};
    if (e < t1) {
        // source bounding box of the tile's p region.  Each rounded op of the coordinate is monotone in qx and in qy,
        // so the literal coordinates of the four corners bound every tap.
        const int ti = e / ntj, tj = e - ti * ntj;
        const int qx_lo = 4 * tj * K1_TJ + 1 + sx, qy_lo = 4 * ti * K1_TI + 1 + sy;
        float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float qx = (float)(qx_lo + ((c & 1) ? K1_SPAN_X : 0)), qy = (float)(qy_lo + ((c & 2) ? K1_SPAN_Y : 0));
            const float cix = affine_coord(T.r0, qx, T.r1, qy, T.r2), ciy = affine_coord(T.r3, qx, T.r4, qy, T.r5);
            xmin = fminf(xmin, cix); xmax = fmaxf(xmax, cix); ymin = fminf(ymin, ciy); ymax = fmaxf(ymax, ciy);
        }
        const int bx0 = (int)floorf(xmin), bx1 = (int)floorf(xmax) + 1, by0 = (int)floorf(ymin), by1 = (int)floorf(ymax) + 1;
        // TMA needs the innermost start coordinate 16-byte aligned (an unaligned start faults with
        // "illegal instruction" on B200: scripts/dev/tma_test3.cu), so the box starts at bx0 rounded down to 4
        const int bx0a = bx0 & ~3;
        const bool empty = (bx1 < 0 || bx0 >= W || by1 < 0 || by0 >= H);   // the rotated image is all zero here
        // the host picks the box variant from the same transforms (build_tables): a box that does not fit is a bug, not data
        if (!empty && (bx1 - bx0a >= K1_XS || by1 - by0 >= box_rows)) __trap();
        boxd[slot * t1 + e] = make_int2(bx0a, empty ? K1_EMPTY : by0);
    }
    const int c = e - t1;
    if (c >= 0 && c < w) {
        const int Z1 = 4 * c + 1;
        const float2 a = translate_taps(Z1, T.tx, sx, W), d = translate_taps(Z1 + 1, T.tx, sx, W);
        fcolw[slot * w + c] = make_float4(a.x, a.y, d.x, d.y);
    } else if (c >= w && c < w + h) {
        const int r = c - w, Z1 = 4 * r + 1;
        const float2 a = translate_taps(Z1, T.ty, sy, H), d = translate_taps(Z1 + 1, T.ty, sy, H);
        froww[slot * h + r] = make_float4(a.x, a.y, d.x, d.y);
    }
}

// The gather of one thread: p = rotate-gather of x at p column `pcn` of the tile, rows 12g..12g+11 (q rows qy0 + {0,1,2,4,5,6,8,9,10,12,13,14}),
// from the source box at shared address `box` whose element (0,0) is image pixel (d.y, d.x).  Two rows travel as the two lanes of packed
// fp32 instructions (asr_common.cuh); the row coordinates are small integers, so qy0 + offset is exact and equals the literal (float)qy.
__device__ __forceinline__ void k1_gather(const FwdCopy& T, BoxDesc d, const float* box, float* prow, float qxf, float qy0) {
    const float ax = fmul(T.r0, qxf), ay = fmul(T.r3, qxf);
    // byte address of tap (y0,x0) = 4*(y0*XS + x0) + cst, box origin and tile address folded into cst
    const int cst = (int)smem_u32(box) - 4 * (d.y * K1_XS + d.x);
    const float cstf = denorm_int(cst);
    const f32x2 cstd = pk(cstf, cstf);
    const f32x2 axp = pk(ax, ax), ayp = pk(ay, ay), r2p = pk(T.r2, T.r2), r5p = pk(T.r5, T.r5);
    const f32x2 r1p = pk(T.r1, T.r1), r4p = pk(T.r4, T.r4), qy0p = pk(qy0, qy0);
    const f32x2 magic2 = pk(kMagic, kMagic), one2 = pk(1.0f, 1.0f);
#ifdef ASR_K1_EXP_NOGATHER   // experiment: descriptor + box load + stores only = the latency floor of one CTA per (tile, copy)
#pragma unroll 1
    for (int m = 0; m < 2; m += 2) {
#else
#pragma unroll
    for (int m = 0; m < 12; m += 2) {
#endif
        const int o0 = 4 * (m / 3) + m % 3, o1 = 4 * ((m + 1) / 3) + (m + 1) % 3;
        const f32x2 qy2 = add2(qy0p, pk((float)o0, (float)o1));
        const f32x2 ix = add2(sum2(axp, mul2(r1p, qy2)), r2p);      // fl(fl(fl(r0*qx) + fl(r1*qy)) + r2)
        const f32x2 iy = add2(sum2(ayp, mul2(r4p, qy2)), r5p);
        const f32x2 fxf = sub2(add2_rd(ix, magic2), magic2), fyf = sub2(add2_rd(iy, magic2), magic2);   // floors
        // (x_ceil - x) == 1 - (x - x_floor) bit for bit unless x in (-1,0), where that weight only ever
        // multiplies the out-of-image tap x_floor = -1, i.e. an exact zero
        const f32x2 wx1 = sub2(ix, fxf), wx0 = sub2(one2, wx1);
        const f32x2 wy1 = sub2(iy, fyf), wy0 = sub2(one2, wy1);
        const f32x2 tp = tap_addr2<K1_XS>(fxf, fyf, cstd);
        const unsigned ta = (unsigned)tp, tb = (unsigned)(tp >> 32);
        const f32x2 o = bilerp2(pk(lds_tap<0>(ta), lds_tap<0>(tb)), pk(lds_tap<4>(ta), lds_tap<4>(tb)),
                                pk(lds_tap<4 * K1_XS>(ta), lds_tap<4 * K1_XS>(tb)),
                                pk(lds_tap<4 * K1_XS + 4>(ta), lds_tap<4 * K1_XS + 4>(tb)), wx0, wx1, wy0, wy1);
        prow[m * K1_PBS] = pk_lo(o);
        prow[(m + 1) * K1_PBS] = pk_hi(o);
    }
}

// One LR cell: translate (2x2 z values from the cell's 3x3 p patch `pr`), resize (literal lerps at 0.5).
__device__ __forceinline__ float k1_cell(const float* pr, float4 wc, float4 wr) {
    float Tx[3][2];
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
        const float p0 = pr[bb * K1_PBS], p1 = pr[bb * K1_PBS + 1], p2 = pr[bb * K1_PBS + 2];
        Tx[bb][0] = fadd(fmul(wc.x, p0), fmul(wc.y, p1));
        Tx[bb][1] = fadd(fmul(wc.z, p1), fmul(wc.w, p2));
    }
    const float tl = fadd(fmul(wr.x, Tx[0][0]), fmul(wr.y, Tx[1][0]));
    const float tr = fadd(fmul(wr.x, Tx[0][1]), fmul(wr.y, Tx[1][1]));
    const float bl = fadd(fmul(wr.z, Tx[1][0]), fmul(wr.w, Tx[2][0]));
    const float br = fadd(fmul(wr.z, Tx[1][1]), fmul(wr.w, Tx[2][1]));
    const float top = fadd(tl, fmul(fsub(tr, tl), 0.5f));
    const float bot = fadd(bl, fmul(fsub(br, bl), 0.5f));
    return fadd(top, fmul(fsub(bot, top), 0.5f));
}

template <int XR>
__global__ void __launch_bounds__(K1_THREADS)
k_forward_residual(const __grid_constant__ CUtensorMap xmap, const float* __restrict__ copies, float* __restrict__ resid,
                   const FwdCopy* __restrict__ fcp, const float4* __restrict__ fcolw, const float4* __restrict__ froww,
                   const BoxDesc* __restrict__ boxd, const ImgParams* __restrict__ ip, int it, int flags, int N, int h, int w, int wp,
                   int ntj, unsigned ntj_magic, int b_base) {
    const int b = blockIdx.z, ks = blockIdx.y;
    // flags: bit 0 = look at ImgParams (the host clears it when every image of the launch is live), bit 1 = let the next kernel in early
    if ((flags & 1) && (ks >= ip[b].n_kept || it >= ip[b].num_iter)) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    float* xt = reinterpret_cast<float*>(smem_raw);                              // [XR][K1_XS], filled by TMA
    float* pb = xt + K1_XS * XR;                                                 // [K1_PR][K1_PBS]
    int* boxs = reinterpret_cast<int*>(pb + K1_PR * K1_PBS);                     // bx0a, by0

    // gather phase (threads 0..191): p column and row group of 12 rows = 4 cell rows; the thread id maps to one LR cell in the second phase.
    // (A conflict-free map -- 24 of 32 lanes active, half a p row per warp -- was measured slower: 24.7 vs 23.9 us, it trades 37 % of the
    //  shared-memory wavefronts for 15 % more instructions and the kernel is as much issue- as shared-memory bound.)
    const int tid = threadIdx.x;
    const int g = tid / K1_PC, pcn = tid - g * K1_PC;
    const int prow0 = 12 * g, qrow0 = 16 * g;
    const bool gathers = tid < K1_GATHER;
    const unsigned slot = (unsigned)b * (unsigned)N + (unsigned)ks;              // B*N*max(h,w,tiles) < 2^32 is checked on the host
    if (tid == 0) {
        const BoxDesc d = boxd[(size_t)slot * gridDim.x + blockIdx.x];
        boxs[0] = d.x; boxs[1] = d.y;
        mbar_init(&bar, 1);
        // x is the previous kernel's output: everything above (and the operand loads of the other threads below) runs while that kernel
        // drains; the next kernel of the chain may be scheduled as soon as every CTA of this grid is past this point
        pdl_wait();
        if (flags & 2) pdl_trigger();   // else the next kernel is let in when this grid's CTAs exit
#ifndef ASR_K1_EXP_NOTMA   // experiment: no box load at all = the compute-only floor of this kernel (profiles/r02_k1_floors.txt)
        if (d.y != K1_EMPTY) tma_load_3d(xt, &xmap, d.x, d.y, b_base + b, &bar, (unsigned)(K1_XS * XR * sizeof(float)));
#endif
    }
    const int ti = (ntj == 1) ? (int)blockIdx.x : (int)__umulhi(blockIdx.x, ntj_magic);          // tile row (2^32/1 does not fit the magic)
    const int tj = (int)blockIdx.x - ti * ntj;                                                    // tile column
    const int j0 = tj * K1_TJ, i0 = ti * K1_TI;
    const FwdCopy T = fcp[slot];
    const int qx_lo = 4 * j0 + 1 + T.sx, qy_lo = 4 * i0 + 1 + T.sy;   // first needed p position
    // this thread's cell of the second phase: request its LR sample and tap weights now, they are consumed at the very end
    const int ci = tid / K1_TJ, cj = tid % K1_TJ;
    const int i = i0 + ci, j = j0 + cj;
    float yk = 0.0f;
    float4 wc = make_float4(0.f, 0.f, 0.f, 0.f), wr = wc;
    const bool live = i < h && j < w;
    if (live) {
        yk = __ldg(copies + (size_t)T.ysrc * (unsigned)(h * w) + (unsigned)(i * w + j));
        wc = __ldg(fcolw + (slot * (unsigned)w + (unsigned)j));
        wr = __ldg(froww + (slot * (unsigned)h + (unsigned)i));
    }
    pdl_wait();        // (returns at once for every thread but the first to ask; the residual stores below must not pass the previous kernel)
    __syncthreads();   // box origin and barrier init visible
    const BoxDesc d = make_int2(boxs[0], boxs[1]);
    const bool empty = d.y == K1_EMPTY;

    if (!empty && gathers) {
#ifndef ASR_K1_EXP_NOTMA
        mbar_wait(&bar, 0);
#endif
        k1_gather(T, d, xt, pb + prow0 * K1_PBS + pcn, (float)(qx_lo + pcn + pcn / 3) /* 4*(pcn/3) + pcn%3 */, (float)(qy_lo + qrow0));
    }
    __syncthreads();

    // ---- one cell per thread: translate, resize, minus y ---------
    if (live) {
        const float D = empty ? 0.0f : k1_cell(pb + 3 * ci * K1_PBS + 3 * cj, wc, wr);
        resid[(size_t)slot * (unsigned)(h * wp) + (unsigned)(i * wp + j)] = fsub(D, yk);
    }
}

// Measured and rejected (round 2, profiles/r02_k1_floors.txt): a pipelined form -- one CTA walks 10-20 consecutive copies of its tile with two
// source boxes and two p buffers in rotation, next box and next operands in flight during the gather, one __syncthreads per copy.  It was
// bit-identical and slower, 26.5 vs 22.6 us: two boxes per CTA leave 3 CTAs = 18 gather warps per SM, and the gather's dependent
// coordinate -> floor -> address -> LDS -> lerp chain needs the 36 of the one-copy kernel more than it needs the hidden TMA latency.
// A second form kept ONE box and one p buffer (still 6 CTAs/SM): an idle thread staged the next copy's record and box origin in shared memory
// during the gather and the next box was requested right after the gather's barrier, so that its flight overlapped the cells.  22.4 vs 22.6 us
// at 64 images, 23.7 vs 23.4 at 8: the six resident CTAs already hide each other's chains; not worth a second kernel.

// ================================================================================================
// K2: gradient + regularisers + optimizer step
// ================================================================================================
// TensorFlow's gradient of the data term for copy k at HR pixel X is
//     v_k(X) = bilinear(u_k, Rinv_k X),   u_k(q) = bilinear(g_hr, q + d_k),
//     g_hr = ResizeBilinearGrad(2*lambda_df*r_k) = 0.25*g on the four positions {4i+1,4i+2}x{4j+1,4j+2}.
// The CTA owns a 64x64 HR tile and loops over the copies with three kinds of warps:
//   * 1 producer warp issues the async staging copies (residual box by TMA, tap-table rows by bulk copies) three copies ahead;
//   * 7 fill warps materialise u_k on the bounding box of Rinv_k(tile) in shared memory, one LR cell ->
//     the 4x4 block of q positions that cell feeds (q+floor(d) in [4c,4c+3]); the block's values follow
//     the literal 2-tap sums, which collapse to w_b*g / (w_a*g + w_b*g) / w_a*g / 0 by phase because the
//     other tap reads an exact zero.  Phase 3 is a whole zero row/column of the tile: those rows are zeroed
//     once and never rewritten.
//   * 8 gather warps read that tile: every thread gathers 16 pixels (8 rows x the columns lane, lane+32,
//     the column pair travelling as the two lanes of packed fp32 instructions) with the op's exact
//     arithmetic and adds them to its accumulators in ascending copy order.
// The two u buffers are handed back and forth like this: "full" is an mbarrier the fill warps arrive on and
// each gather warp tests on its own (the fill runs ahead, so the test normally succeeds at once and the
// gather warps never wait for each other); "empty" is a named barrier the gather warps arrive on without
// blocking and the fill warps sleep on in hardware (no spin, no issue slots).  The fill of copy k+1 thus
// overlaps the gather of copy k and the gather warps carry no bookkeeping at all.
// Bounding boxes and the inverse transforms of a chunk of 128 copies are computed once into shared
// memory (one thread per copy, by the gather side); the roles part for good right after the CTA's set-up.
constexpr int K2_T = 64;               // HR tile width; the tile height TY is 64 for batches that fill the GPU and 32 for one or two images
                                       // (one 512^2 image is only 64 tiles of 64x64: such a solve is latency-bound, 104 -> 79 us per iteration)
#ifndef ASR_K2_GW
#define ASR_K2_GW 8
#endif
constexpr int K2_GW = ASR_K2_GW;       // gather warps; thread owns pixels (lane + 32c, warp + 8r), c<2, r<TY/8
// non-gather warps: 8 in both variants (7 fill + 1 producer): the throughput variant (64-row tiles, two CTAs per SM) and the
// latency variant (32-row tiles, one lone CTA per SM, where the fill's serial latency per copy is what the gather warps wait for)
#ifndef ASR_K2_FW64
#define ASR_K2_FW64 8
#endif
#ifndef ASR_K2_AHEAD
#define ASR_K2_AHEAD 3
#endif
#ifndef ASR_K2_FILLERS      // fill warps of the throughput variant; measured (us per image-iteration): 4: 30.0, 5: 30.5, 6: 30.4, 7: 28.1
#define ASR_K2_FILLERS 7
#endif
#ifndef ASR_K2_PRODUCER
#define ASR_K2_PRODUCER 1
#endif
// non-gather warps of a CTA.  With 8 of them the last one is a PRODUCER: it only issues the async staging copies (the clock64
// trace showed the issuing thread's ~800 clk per copy sitting on the fill's critical path when a fill warp did it), the other
// seven fill.
template <int TY> struct K2Fill {
    static constexpr int warps = TY == 64 ? ASR_K2_FW64 : 8, threads = 32 * (K2_GW + warps), ctas = TY == 64 ? 2 : 1;
    static constexpr bool producer = ASR_K2_PRODUCER && warps == 8;
    static constexpr int fillers = producer ? (TY == 64 ? ASR_K2_FILLERS : warps - 1) : warps;   // warps beyond fillers + producer leave at once
    static constexpr int part = 32 * (K2_GW + fillers + (producer ? 1 : 0));                   // threads that take part in the barriers
};
// The throughput variant (64-row tiles, two CTAs per SM) runs 8 gather warps + 7 fill warps + 1 producer warp with a register split
// between the roles (setmaxnreg, per warpgroup of 4 warps): with 16 warps and two CTAs per SM the launch gives every thread 64
// registers; the fill/producer warpgroups hand theirs back down to 40 and the gather warpgroups grow to 88, the budget their 16
// accumulators + 8-deep unrolled gather needs (8*88 + 8*40 = 16*64).  Measured (r02, us per image-iteration at 250 images):
// 4 fill warps, staging 2 ahead 29.2; 8 fill warps 30.2 (no gain by itself: the loop is issue-bound); 7 fill + producer,
// staging 3 ahead 28.4.  -DASR_K2_FW64=4 -DASR_K2_AHEAD=2 rebuilds the 12-warp variant.
constexpr bool K2_REG_SPLIT = (ASR_K2_FW64 == 8) && (K2_GW == 8);
constexpr int K2_AHEAD = ASR_K2_AHEAD;   // copies the async staging (residual box + tap rows) runs ahead of the fill (at most K2_STAGES - 1)
constexpr int K2_NG = 32 * K2_GW;
constexpr int K2_US = 96;              // u tile stride: 64*sqrt(2)+2+3 < 96, multiple of 32
template <int TY> struct K2Rows { static constexpr int value = TY == 64 ? 96 : 80; };   // u tile rows: sqrt(63^2+(TY-1)^2)+2 -> cells
constexpr int K2_CHUNK = 128;          // copies whose boxes/transforms are staged at once
enum { BAR_EMPTY = 1 };   // named barriers 1,2 (0 is __syncthreads)
struct __align__(16) KBox {
    int cst;        // word offset of the box origin inside a u buffer: -(qy_lo*US + qx_lo), buffer offset excluded
    int cbx0, cby0; // first LR cell of the box
    int ncxy;       // ncx | ncy << 8 | skip << 16   (skip: box does not touch the LR grid, u == 0)
    int qx_lo, qy_lo;
    int inv_ncx;    // ceil(65536 / ncx) for the cell index split
    int pad;
};
// what a gather warp needs of a copy, in one 32-byte record (two LDS.128): the rotate coefficients it multiplies itself, the tap
// address constant of either u buffer already in the denormal-float form tap_addr2 takes, and the skip flag
struct __align__(16) KGather { float b0, b3, b2, b5, cstf0, cstf1; int skip, pad; };
// Per-copy inputs of the fill, staged by async copies two copies ahead: the LR residual box (TMA tensor load,
// cells outside the grid arrive as zeros) and the translate tap tables of the box's cell columns / rows.
constexpr int K2_TPAD = 24;            // tap tables carry 24 cells of halo on both sides (boxes are <= 24 cells)
constexpr int K2_BC = K2_US / 4;       // box cells per axis (24)
constexpr int K2_RW = K2_BC + 4;       // residual box width: the TMA start column is rounded down to a multiple of 4
constexpr int K2_STAGES = 4;
struct __align__(128) K2Stage {
    float r[K2_BC][K2_RW];             // 2688 B
    float4 ctap[K2_BC][2];             // (wa0,wb0,wa1,wb1) (wa2,wb2,wa3,wb3) per cell column, 768 B
    float4 rtap[K2_BC][2];             // same per cell row
};
constexpr unsigned K2_RBOX_BYTES = sizeof(float) * K2_BC * K2_RW, K2_TAP_BYTES = sizeof(float4) * K2_BC * 2;
template <int TY>
constexpr size_t k2_smem() {
    return sizeof(float) * 2 * K2_US * K2Rows<TY>::value + (sizeof(KBox) + sizeof(InvXf) + sizeof(KGather)) * K2_CHUNK +
           sizeof(K2Stage) * K2_STAGES + sizeof(float2) * 2 * TY;
}

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// literal taps of the inverse-translate gather on the window (q+s, q+s+1).  u_k is an image on the
// HR canvas: the rotate-gradient gather zero-fills q outside [0,limit), so those columns/rows get
// zero weights (their u must read as 0 even where q+d lands on a non-zero g_hr).
__device__ __forceinline__ float2 inv_translate_taps(int q, float u, int s, int limit) {
    if (q < 0 || q >= limit) return make_float2(0.0f, 0.0f);
    const float iq = fadd((float)q, u);
    const float f = floorf(iq);
    const float w0 = fsub(fadd(f, 1.0f), iq), w1 = fsub(iq, f);
    return ((int)f == q + s) ? make_float2(w0, w1) : make_float2(0.0f, w0);
}

// Translate tap tables, once per solve: for copy slot k and LR cell c (with K2_TPAD cells of halo) the four
// literal tap pairs of q = 4c - floor(u) + e, e < 4.  They depend on the transform only, not on the iterate.
__global__ void k_tap_tables(const InvXf* __restrict__ inv, float2* __restrict__ tapc, float2* __restrict__ tapr, int N, int h,
                             int w, int H, int W) {
    const int k = blockIdx.y, b = blockIdx.z;
    const int nc = 4 * (w + 2 * K2_TPAD), nr = 4 * (h + 2 * K2_TPAD);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nc + nr) return;
    const InvXf T = inv[(size_t)b * N + k];
    const size_t slot = (size_t)b * N + k;
    if (e < nc) {
        const int s = (int)floorf(T.ux);
        tapc[slot * nc + e] = inv_translate_taps(4 * (e / 4 - K2_TPAD) - s + (e & 3), T.ux, s, W);
    } else {
        const int f = e - nc, s = (int)floorf(T.uy);
        tapr[slot * nr + f] = inv_translate_taps(4 * (f / 4 - K2_TPAD) - s + (f & 3), T.uy, s, H);
    }
}

// One optimizer step on one pixel in TensorFlow's expression order (ApplyAdam[WithAmsgrad], ApplyAdagradV2, ApplyAdadelta,
// ApplyAdaMax, ApplyGradientDescent, ApplyKerasMomentum): slots o0,o1,o2 in, updated slots written to s0,s1,s2[gi].
__device__ __forceinline__ float optimizer_step(const ImgParams& P, const Sched sc, float g, float xi, float o0, float o1, float o2,
                                                size_t gi, float* __restrict__ s0, float* __restrict__ s1, float* __restrict__ s2) {
    float xn;
    switch (P.optimizer) {
    case ASR_OPT_SGD:
        if (P.momentum == 0.0f) {
            xn = fsub(xi, fmul(sc.x, g));
        } else {
            const float a = fsub(fmul(o0, P.momentum), fmul(sc.x, g));
            s0[gi] = a;
            xn = P.nesterov ? fadd(xi, fsub(fmul(a, P.momentum), fmul(sc.x, g))) : fadd(xi, a);
        }
        break;
    case ASR_OPT_ADAGRAD: {
        const float a = fadd(o0, fmul(g, g));
        s0[gi] = a;
        xn = fsub(xi, __fdiv_rn(fmul(g, sc.x), fadd(__fsqrt_rn(a), P.epsilon)));
    } break;
    case ASR_OPT_ADADELTA: {
        const float rho = 0.95f, eps = 1e-7f, omr = fsub(1.0f, rho);
        const float a = fadd(fmul(o0, rho), fmul(fmul(g, g), omr));
        s0[gi] = a;
        const float upd = fmul(fmul(__fsqrt_rn(fadd(o1, eps)), __fdiv_rn(1.0f, __fsqrt_rn(fadd(a, eps)))), g);
        xn = fsub(xi, fmul(upd, sc.x));
        s1[gi] = fadd(fmul(o1, rho), fmul(fmul(upd, upd), omr));
    } break;
    case ASR_OPT_ADAMAX: {
        const float m = fadd(o0, fmul(fsub(g, o0), P.omb1));
        s0[gi] = m;
        const float v = fmaxf(fmul(P.beta_2, o1), fabsf(g));
        s1[gi] = v;
        xn = fsub(xi, fmul(sc.y, __fdiv_rn(m, fadd(v, P.epsilon))));
    } break;
    default: {
        const float m = fadd(o0, fmul(fsub(g, o0), P.omb1));
        const float v = fadd(o1, fmul(fsub(fmul(g, g), o1), P.omb2));
        s0[gi] = m;
        s1[gi] = v;
        float den = v;
        if (P.amsgrad) { den = fmaxf(o2, v); s2[gi] = den; }
        xn = fsub(xi, __fdiv_rn(fmul(m, sc.y), fadd(__fsqrt_rn(den), P.epsilon)));
    } break;
    }
    return xn;
}

#ifdef ASR_K2_TRACE   // development build only (scripts/dev/k2_trace.py): clock64 stamps of one CTA's hand-offs
__device__ long long g_k2_trace[16][128][4];
__device__ long long g_k2_misc[16][4];
#define K2_TR(slot) do { if (blockIdx.x == 37 && blockIdx.y == 1 && lane == 0 && it == 2) g_k2_trace[warp][kc][slot] = clock64(); } while (0)
#define K2_TM(slot) do { if (blockIdx.x == 37 && blockIdx.y == 1 && lane == 0 && it == 2) g_k2_misc[warp][slot] = clock64(); } while (0)
#else
#define K2_TR(slot) do {} while (0)
#define K2_TM(slot) do {} while (0)
#endif

template <bool WRITE_GRAD, bool BTV, int TY>
__global__ void __launch_bounds__(K2Fill<TY>::threads, K2Fill<TY>::ctas)
k_gradient_update(const __grid_constant__ CUtensorMap rmap, const float* __restrict__ x_cur, float* __restrict__ x_next,
                  float* __restrict__ s0, float* __restrict__ s1, float* __restrict__ s2, const float2* __restrict__ tapc,
                  const float2* __restrict__ tapr, const InvXf* __restrict__ inv, const ImgParams* __restrict__ ip,
                  const Sched* __restrict__ sched, int it, int N, int h, int w, int H, int W, int B, int b_base) {
    const int b = blockIdx.y;
    const ImgParams P = ip[b];
    if (it >= P.num_iter) return;
    constexpr int K2_UR = K2Rows<TY>::value, K2_ROWS = TY / K2_GW;
    constexpr int K2_PART = K2Fill<TY>::part;   // threads that take part in the chunk barriers and the "empty" hand-off
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ut = reinterpret_cast<float*>(smem_raw);                       // [2][K2_UR][K2_US]
    K2Stage* stages = reinterpret_cast<K2Stage*>(ut + 2 * K2_US * K2_UR);  // [K2_STAGES]
    KBox* boxes = reinterpret_cast<KBox*>(stages + K2_STAGES);            // [K2_CHUNK]
    InvXf* xfs = reinterpret_cast<InvXf*>(boxes + K2_CHUNK);              // [K2_CHUNK]
    KGather* gth = reinterpret_cast<KGather*>(xfs + K2_CHUNK);            // [K2_CHUNK]
    float2* rowp = reinterpret_cast<float2*>(gth + K2_CHUNK);             // [2][TY] (fl(b1*Y), fl(b4*Y)) of the tile's rows, per u buffer
    __shared__ __align__(8) unsigned long long stage_bar[K2_STAGES], full_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntx = (W + K2_T - 1) / K2_T;
    const int tx0 = (blockIdx.x % ntx) * K2_T, ty0 = (blockIdx.x / ntx) * TY;
    const InvXf* invb = inv + (size_t)b * N;
    const int nk = P.n_kept;
    if (!WRITE_GRAD) {
        // the epilogue's operands (x, optimizer slots) are asked into the L2 now, a whole copy loop before they are needed
        constexpr int LINES = TY * (K2_T * 4 / 128);   // 128-byte lines per array per tile
        for (int i = tid; i < 4 * LINES; i += K2Fill<TY>::threads) {
            const int arr = i / LINES, l = i - arr * LINES, row = l / (K2_T * 4 / 128), seg = l - row * (K2_T * 4 / 128);
            const int X = tx0 + 32 * seg, Y = ty0 + row;
            if (X < W && Y < H) {
                const float* base = arr == 0 ? x_cur : (arr == 1 ? s0 : (arr == 2 ? s1 : s2));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)b * H * W + (size_t)Y * W + X));
            }
        }
    }
    K2_TM(0);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < K2_STAGES; ++i) mbar_init(&stage_bar[i], 3);   // three async copies per stage
        mbar_init(&full_bar[0], K2Fill<TY>::fillers); mbar_init(&full_bar[1], K2Fill<TY>::fillers);   // one arrival per fill warp
    }
    const bool gather_role = warp < K2_GW;

    // tile rows 3 (mod 4) hold phase 3 of every cell: zero for every copy
    for (int i = tid; i < 2 * (K2_UR / 4) * (K2_US / 4) && tid < K2_PART; i += K2_PART) {
        const int row = i / (K2_US / 4), c4 = i - row * (K2_US / 4);
        reinterpret_cast<float4*>(ut + (4 * row + 3) * K2_US)[c4] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // rows of buffer 1 follow buffer 0
    }

    // The roles part here and never meet again except at the chunk barriers (bar 0, every thread, twice per chunk), so that each
    // side can be compiled and run with its own register budget.
    if (!gather_role) {
        if (TY == 64 && K2_REG_SPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // every warp of both warpgroups, same instruction
        if (tid >= K2_PART) return;   // warps beyond the fill warps and the producer only exist to make whole warpgroups
        // The residual is the previous kernel's output and only the staging thread touches it (through TMA): it waits here, while the gather
        // warps build the first chunk's boxes, and lets the next kernel in.  (Kept out of the chunk loop: inside it the staging path lost
        // its uniform-register address arithmetic and K2 ran 3 % slower.)
        if (lane == 0 && warp - K2_GW == (K2Fill<TY>::producer ? K2Fill<TY>::fillers : 0)) {
            pdl_wait();
            pdl_trigger();
        }
        for (int k0 = 0; k0 < nk; k0 += K2_CHUNK) {
            const int nc = min(K2_CHUNK, nk - k0);
            bar_sync(0, K2_PART);   // the previous chunk's boxes are no longer read
            bar_sync(0, K2_PART);   // this chunk's boxes and transforms are in place (written by the gather warps)
            // =================== fill warps ===================
            // Warp fw owns the cell rows fw, fw+4, ... of the box and lane l the cell column l (boxes are at most 24
            // cells wide).  Everything the fill reads was staged by async copies issued two copies earlier by one
            // thread: no address arithmetic, bounds tests or table building is left in these warps.
            const int fw = warp - K2_GW;
            constexpr int K2_FILLERS = K2Fill<TY>::fillers;
            constexpr int ROWS = (K2_UR / 4 + K2_FILLERS - 1) / K2_FILLERS;   // cell rows per fill warp
            const bool is_producer = K2Fill<TY>::producer && fw == K2_FILLERS;   // the warp right after the fill warps
            const size_t slot0 = (size_t)(b_base + b) * N + k0;
            const int ncw = w + 2 * K2_TPAD, nrw = h + 2 * K2_TPAD;
            auto stage_copy = [&](int kq) {   // one thread: residual box + tap rows of copy kq -> stage (k0+kq) % 4
                const KBox bq = boxes[kq];
                K2Stage* S = stages + ((k0 + kq) & (K2_STAGES - 1));
                unsigned long long* bar = &stage_bar[(k0 + kq) & (K2_STAGES - 1)];
                tma_load_3d(&S->r[0][0], &rmap, bq.cbx0 & ~3, bq.cby0, (int)(slot0 + kq), bar, K2_RBOX_BYTES);
                bulk_load(&S->ctap[0][0], tapc + ((slot0 + kq) * ncw + bq.cbx0 + K2_TPAD) * 4, K2_TAP_BYTES, bar);
                bulk_load(&S->rtap[0][0], tapr + ((slot0 + kq) * nrw + bq.cby0 + K2_TPAD) * 4, K2_TAP_BYTES, bar);
            };
            const bool issuer = lane == 0 && (K2Fill<TY>::producer ? is_producer : fw == 0);
            int next_q = 0;   // issuer only: first copy of the chunk whose staging has not been issued yet
            if (issuer) for (; next_q < K2_AHEAD && next_q < nc; ++next_q) stage_copy(next_q);
            for (int kc = 0; kc < nc; ++kc) {
                const int ub = kc & 1, gk = k0 + kc;
                const KBox bc = boxes[kc];
                const bool live = !(bc.ncxy >> 16);
                const int ncx = bc.ncxy & 0xff, ncy = (bc.ncxy >> 8) & 0xff;
                K2_TR(0);
                if (kc >= 2) { if (ub) bar_sync(BAR_EMPTY + 1, K2_PART); else bar_sync(BAR_EMPTY, K2_PART); }   // the gather of copy kc-2 has left this buffer
                // Keep the staging K2_AHEAD copies ahead.  Copy q reuses the stage of copy q-4, which is free once every fill warp
                // has finished q-4: true when q-4 < 0 (the chunk's first use; the previous chunk ended with __syncthreads) or
                // when the barrier above was passed (kc >= 2: every fill warp arrived for kc, i.e. is done with kc-1 >= q-4).
                K2_TR(1);
                if (issuer)
                    while (next_q <= kc + K2_AHEAD && next_q < nc && (next_q < K2_STAGES || (kc >= 2 && next_q - K2_STAGES <= kc - 1)))
                        stage_copy(next_q++);
                if (is_producer) continue;   // the producer warp neither fills nor arrives on "full"
                mbar_wait(&stage_bar[gk & (K2_STAGES - 1)], (gk / K2_STAGES) & 1);
                K2_TR(2);
                if (live && lane < ncx) {
                    const K2Stage* S = stages + (gk & (K2_STAGES - 1));
                    const float4 ca = S->ctap[lane][0], cb = S->ctap[lane][1];     // (wa0,wb0,wa1,wb1) (wa2,wb2,wa3,wb3)
                    const int xo = (bc.cbx0 & 3) + lane;
                    float4* ucol = reinterpret_cast<float4*>(ut + ub * (K2_US * K2_UR)) + lane;
    #pragma unroll
                    for (int j = 0; j < ROWS; ++j) {
                        const int cyi = fw + K2_FILLERS * j;
                        if (cyi < ncy) {
                            const float g = fmul(0.25f, fmul(P.two_ldf, S->r[cyi][xo]));   // g_hr on the cell's 2x2 positions
                            const float t0 = fmul(ca.y, g);                           // phase 0: taps (0, g)
                            const float t1 = fadd(fmul(ca.z, g), fmul(ca.w, g));      // phase 1: taps (g, g)
                            const float t2 = fmul(cb.x, g);                           // phase 2: taps (g, 0); phase 3: (0, 0)
                            const float4 ra = S->rtap[cyi][0], rc = S->rtap[cyi][1];
                            float4* dst = ucol + cyi * K2_US;                         // row 4*cyi
                            dst[0] = make_float4(fmul(ra.y, t0), fmul(ra.y, t1), fmul(ra.y, t2), 0.0f);
                            dst[K2_US / 4] = make_float4(fadd(fmul(ra.z, t0), fmul(ra.w, t0)), fadd(fmul(ra.z, t1), fmul(ra.w, t1)),
                                                         fadd(fmul(ra.z, t2), fmul(ra.w, t2)), 0.0f);
                            dst[2 * (K2_US / 4)] = make_float4(fmul(rc.x, t0), fmul(rc.x, t1), fmul(rc.x, t2), 0.0f);
                        }
                    }
                }
                if (live && tid - K2_NG < TY) {   // row products of the rotate coordinates for the gather warps
                    const float Yf = (float)(ty0 + tid - K2_NG);
                    rowp[ub * TY + tid - K2_NG] = make_float2(fmul(xfs[kc].b1, Yf), fmul(xfs[kc].b4, Yf));
                }
                __syncwarp();
                K2_TR(3);
                if (lane == 0) mbar_arrive(&full_bar[ub]);
            }
        }
        return;
    }
    if (TY == 64 && K2_REG_SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");

    // the two pixels of a row (columns lane, lane+32) travel as the two lanes of packed fp32 registers
    const float X0f = (float)(tx0 + lane), X1f = (float)(tx0 + lane + 32);
    f32x2 accp[K2_ROWS];
#pragma unroll
    for (int r = 0; r < K2_ROWS; ++r) accp[r] = pk(0.0f, 0.0f);
    const f32x2 magic2 = pk(kMagic, kMagic), one2 = pk(1.0f, 1.0f);

    for (int k0 = 0; k0 < nk; k0 += K2_CHUNK) {
        const int nc = min(K2_CHUNK, nk - k0);
        bar_sync(0, K2_PART);   // the previous chunk's boxes are no longer read
        // ---- chunk prologue: bounding box of Rinv(tile) for each copy (one thread per copy).  Each
        //      rounded op of the coordinate is monotone in X and Y: the four corners bound every tap.
        if (tid < nc) {
            const InvXf T = invb[k0 + tid];
            float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
            for (int cnr = 0; cnr < 4; ++cnr) {
                const float X = (float)(tx0 + ((cnr & 1) ? K2_T - 1 : 0)), Y = (float)(ty0 + ((cnr & 2) ? TY - 1 : 0));
                const float cix = affine_coord(T.b0, X, T.b1, Y, T.b2), ciy = affine_coord(T.b3, X, T.b4, Y, T.b5);
                xmin = fminf(xmin, cix); xmax = fmaxf(xmax, cix); ymin = fminf(ymin, ciy); ymax = fmaxf(ymax, ciy);
            }
            const int qx0 = (int)floorf(xmin), qx1 = (int)floorf(xmax) + 1, qy0 = (int)floorf(ymin), qy1 = (int)floorf(ymax) + 1;
            const int sx = (int)floorf(T.ux), sy = (int)floorf(T.uy);
            KBox bx;
            bx.cbx0 = (qx0 + sx) >> LOG2S;
            bx.cby0 = (qy0 + sy) >> LOG2S;
            const int cbx1 = (qx1 + sx) >> LOG2S, cby1 = (qy1 + sy) >> LOG2S;
            const int ncx = cbx1 - bx.cbx0 + 1, ncy = cby1 - bx.cby0 + 1;
            bx.qx_lo = 4 * bx.cbx0 - sx;
            bx.qy_lo = 4 * bx.cby0 - sy;
            const int skip = (cbx1 < 0 || bx.cbx0 >= w || cby1 < 0 || bx.cby0 >= h);
            // a 64x64 tile rotates into at most 64*sqrt(2)+5 < 96 source rows/columns: a box that does not fit is a bug, not data
            if (!skip && (4 * ncx > K2_US || 4 * ncy > K2_UR)) __trap();
            bx.ncxy = (ncx & 0xff) | ((ncy & 0xff) << 8) | (skip << 16);
            if (skip) { bx.cbx0 = 0; bx.cby0 = 0; }   // its (unused) staging copies stay inside the tables
            bx.cst = -(bx.qy_lo * K2_US + bx.qx_lo);
            bx.inv_ncx = (65536 + ncx - 1) / max(ncx, 1);
            bx.pad = 0;
            boxes[tid] = bx;
            xfs[tid] = T;
            // byte address of tap (y0,x0) = 4*(y0*US + x0) + cst: box origin, buffer and tile address folded into cst
            const int cst0 = 4 * bx.cst + (int)smem_u32(ut), cst1 = cst0 + 4 * (K2_US * K2_UR);
            gth[tid] = KGather{T.b0, T.b3, T.b2, T.b5, denorm_int(cst0), denorm_int(cst1), skip, 0};
        }
        bar_sync(0, K2_PART);

        // =================== gather warps ===================
        K2_TM(1);
        for (int kc = 0; kc < nc; ++kc) {
            K2_TR(0);
            mbar_wait(&full_bar[kc & 1], ((k0 + kc) >> 1) & 1);   // u tile of copy kc is complete (no rendezvous among the gather warps)
            K2_TR(1);
            const KGather Gk = gth[kc];
            if (!Gk.skip) {
                float cstf = (kc & 1) ? Gk.cstf1 : Gk.cstf0;
                asm volatile("" : "+f"(cstf));   // opaque and ordered after the wait above: no tap load can be hoisted over it
                const f32x2 cstd = pk(cstf, cstf);
                const f32x2 b2p = pk(Gk.b2, Gk.b2), b5p = pk(Gk.b5, Gk.b5);
                // products stay scalar (a packed product feeding a packed sum would be contracted, asr_common.cuh)
                const f32x2 axp = pk(fmul(Gk.b0, X0f), fmul(Gk.b0, X1f)), ayp = pk(fmul(Gk.b3, X0f), fmul(Gk.b3, X1f));
#pragma unroll
                for (int r = 0; r < K2_ROWS; ++r) {
                    const float2 rp = rowp[(kc & 1) * TY + warp + K2_GW * r];   // row products, built by the fill warps
                    const float bxr = rp.x, byr = rp.y;
                    const f32x2 ix = add2(add2(axp, pk(bxr, bxr)), b2p);
                    const f32x2 iy = add2(add2(ayp, pk(byr, byr)), b5p);
                    // floor on both lanes: fl_rd(v + 1.5*2^23) - 1.5*2^23
                    const f32x2 fxf = sub2(add2_rd(ix, magic2), magic2), fyf = sub2(add2_rd(iy, magic2), magic2);
                    // (x_ceil - x) == 1 - (x - x_floor) bit for bit unless x in (-1,0), where that weight only
                    // multiplies the tap x_floor = -1, which lies outside the canvas and is an exact zero of u
                    const f32x2 wx1 = sub2(ix, fxf), wx0 = sub2(one2, wx1);
                    const f32x2 wy1 = sub2(iy, fyf), wy0 = sub2(one2, wy1);
                    const f32x2 tp = tap_addr2<K2_US>(fxf, fyf, cstd);
                    const unsigned ta = (unsigned)tp, tb = (unsigned)(tp >> 32);
                    accp[r] = add2(accp[r], bilerp2(pk(lds_tap<0>(ta), lds_tap<0>(tb)), pk(lds_tap<4>(ta), lds_tap<4>(tb)),
                                                    pk(lds_tap<4 * K2_US>(ta), lds_tap<4 * K2_US>(tb)),
                                                    pk(lds_tap<4 * K2_US + 4>(ta), lds_tap<4 * K2_US + 4>(tb)), wx0, wx1, wy0, wy1));
                }
            }
            K2_TR(2);
            // hand the buffer back; the last two hand-backs of a chunk have no taker
            if (kc + 2 < nc) { if (kc & 1) bar_arrive(BAR_EMPTY + 1, K2_PART); else bar_arrive(BAR_EMPTY, K2_PART); }
        }
    }
    K2_TM(2);

    // ---- epilogue: TV (tf.image.image_gradients), L2, L1, optimizer ---------------
    // Two rows (four pixels) at a time: every global load of the batch -- x, its four neighbours and the
    // optimizer slots -- is issued before the first use, so a batch costs one memory round trip, not one per pixel.
    const size_t plane = (size_t)H * W;
    const float* xc = x_cur + (size_t)b * plane;
    const Sched sc = sched[(size_t)it * B + b];
    const bool btv = BTV && P.use_btv;
    const bool need0 = !WRITE_GRAD && !(P.optimizer == ASR_OPT_SGD && P.momentum == 0.0f);
    const bool need1 = !WRITE_GRAD && (P.optimizer == ASR_OPT_ADAM || P.optimizer == ASR_OPT_ADADELTA || P.optimizer == ASR_OPT_ADAMAX);
    const bool need2 = !WRITE_GRAD && P.optimizer == ASR_OPT_ADAM && P.amsgrad;
    // The accumulators go through shared memory (each thread reads back only what it wrote: no barrier) so that the batches can be a
    // real loop: fully unrolled, the epilogue was ~2000 instructions per thread executed once per tile, an instruction-cache miss
    // per line on top of its memory round trips.  The staging ring is free by now: the last fill finished before the last gather began.
    float* gs = reinterpret_cast<float*>(stages);
    static_assert(sizeof(K2Stage) * K2_STAGES >= sizeof(float) * TY * K2_T, "accumulator staging does not fit the stage ring");
#pragma unroll
    for (int r = 0; r < K2_ROWS; ++r) {
        gs[(warp + K2_GW * r) * K2_T + lane] = pk_lo(accp[r]);
        gs[(warp + K2_GW * r) * K2_T + lane + 32] = pk_hi(accp[r]);
    }
    constexpr int EB = 2;
#pragma unroll 1
    for (int r0 = 0; r0 < K2_ROWS; r0 += EB) {
        float xi_[EB][2], nu_[EB][2], nl_[EB][2], nd_[EB][2], nr_[EB][2], v0_[EB][2], v1_[EB][2], v2_[EB][2];
#pragma unroll
        for (int rr = 0; rr < EB; ++rr) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int X = tx0 + lane + 32 * c, Y = ty0 + warp + K2_GW * (r0 + rr);
                const bool in = X < W && Y < H;
                const size_t i = (size_t)Y * W + X, gi = (size_t)b * plane + i;
                xi_[rr][c] = in ? xc[i] : 0.0f;
                nu_[rr][c] = (in && !btv && Y > 0) ? xc[i - W] : 0.0f;
                nl_[rr][c] = (in && !btv && X > 0) ? xc[i - 1] : 0.0f;
                nd_[rr][c] = (in && !btv && Y < H - 1) ? xc[i + W] : 0.0f;
                nr_[rr][c] = (in && !btv && X < W - 1) ? xc[i + 1] : 0.0f;
                v0_[rr][c] = (in && need0) ? s0[gi] : 0.0f;
                v1_[rr][c] = (in && need1) ? s1[gi] : 0.0f;
                v2_[rr][c] = (in && need2) ? s2[gi] : 0.0f;
            }
        }
#pragma unroll
        for (int rr = 0; rr < EB; ++rr) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int r = r0 + rr;
                const int X = tx0 + lane + 32 * c, Y = ty0 + warp + K2_GW * r;
                if (X >= W || Y >= H) continue;
                const size_t i = (size_t)Y * W + X;
                const float xi = xi_[rr][c];
                float g = gs[(warp + K2_GW * r) * K2_T + lane + 32 * c];
                if (btv) {
                    // bilateral TV: 15 integer shifts (h in [-2,2], v in [0,2]) by nearest translate with zero fill;
                    // d/dx of w*|x - S(x)| is w*sign(d) here minus the same term pulled back by the inverse shift
                    for (int hh = -2; hh <= 2; ++hh) {
                        for (int vv = 0; vv <= 2; ++vv) {
                            const float lw = P.btv_lw[abs(hh) + vv];
                            const int xs = X - hh, ys = Y - vv, xt = X + hh, yt = Y + vv;
                            const float shifted = (xs >= 0 && xs < W && ys >= 0) ? xc[(size_t)ys * W + xs] : 0.0f;
                            const float gd = fmul(lw, sgn(fsub(xi, shifted)));
                            const float gb = (xt >= 0 && xt < W && yt < H) ? fmul(lw, sgn(fsub(xc[(size_t)yt * W + xt], xi))) : 0.0f;
                            g = fadd(g, fsub(gd, gb));
                        }
                    }
                } else {
                    if (Y > 0) g = fadd(g, fmul(P.lambda_tv, sgn(fsub(xi, nu_[rr][c]))));
                    if (X > 0) g = fadd(g, fmul(P.lambda_tv, sgn(fsub(xi, nl_[rr][c]))));
                    if (Y < H - 1) g = fsub(g, fmul(P.lambda_tv, sgn(fsub(nd_[rr][c], xi))));
                    if (X < W - 1) g = fsub(g, fmul(P.lambda_tv, sgn(fsub(nr_[rr][c], xi))));
                }
                g = fadd(g, fmul(P.lambda_l2, fmul(xi, 2.0f)));
                if (P.lambda_l1 > 0.0f) g = fadd(g, fmul(P.lambda_l1, sgn(xi)));
                const size_t gi = (size_t)b * plane + i;
                if (WRITE_GRAD) { x_next[gi] = g; continue; }
                const float o0 = v0_[rr][c], o1 = v1_[rr][c], o2 = v2_[rr][c];
                const float xn = optimizer_step(P, sc, g, xi, o0, o1, o2, gi, s0, s1, s2);
                x_next[gi] = xn;
            }
        }
    }
    K2_TM(3);
}

// ================================================================================================
// Any even integer output/feature ratio other than 4 (Superresolution's default feature_size (64,64) -> (512,512) is x8,
// superresolution.py:28): the literal operator sequence, one thread per output, no tiling.  Correct for every even ratio
// S >= 2; only the x4 shape of the reference's callers has the tuned kernels above.  Same arithmetic, bit-identical to the oracle.
// ================================================================================================
// bilinear_interpolation() of ImageProjectiveTransformV3 at source coordinate (ix, iy); rd(y, x) supplies the taps (zero fill inside)
template <class Read>
__device__ __forceinline__ float proj_bilinear(float ix, float iy, Read rd) {
    const float fx = floorf(ix), fy = floorf(iy);
    const float cx = fadd(fx, 1.0f), cy = fadd(fy, 1.0f);
    const long x0 = (long)fx, y0 = (long)fy, x1 = (long)cx, y1 = (long)cy;
    const float wx0 = fsub(cx, ix), wx1 = fsub(ix, fx), wy0 = fsub(cy, iy), wy1 = fsub(iy, fy);
    const float top = fadd(fmul(wx0, rd(y0, x0)), fmul(wx1, rd(y0, x1)));
    const float bot = fadd(fmul(wx0, rd(y1, x0)), fmul(wx1, rd(y1, x1)));
    return fadd(fmul(wy0, top), fmul(wy1, bot));
}

// tf.image.resize weights of output index o (half-pixel centres)
__device__ __forceinline__ void resize_taps(int o, float scale, int in_size, int& lo, int& hi, float& lerp) {
    const float src = fsub(fmul(fadd((float)o, 0.5f), scale), 0.5f);
    const float f = floorf(src);
    lo = max((int)f, 0);
    hi = min((int)ceilf(src), in_size - 1);
    lerp = fsub(src, f);
}

__global__ void __launch_bounds__(128)
kg_forward_residual(const float* __restrict__ x_cur, const float* __restrict__ copies, float* __restrict__ resid,
                    const FwdXf* __restrict__ fwd, const int* __restrict__ src_idx, const ImgParams* __restrict__ ip, int it, int N,
                    int h, int w, int wp, int H, int W) {
    const int b = blockIdx.z, ks = blockIdx.y;
    const ImgParams P = ip[b];
    if (ks >= P.n_kept || it >= P.num_iter) return;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= h * w) return;
    const int i = cell / w, j = cell - i * w;
    const FwdXf T = fwd[(size_t)b * N + ks];
    const float* x = x_cur + (size_t)b * H * W;
    auto xread = [&](long y, long xx) -> float { return (y >= 0 && y < H && xx >= 0 && xx < W) ? x[y * W + xx] : 0.0f; };
    auto p = [&](long qy, long qx) -> float {   // rotated image on the canvas, zero outside
        if (qy < 0 || qy >= H || qx < 0 || qx >= W) return 0.0f;
        const float X = (float)qx, Y = (float)qy;
        return proj_bilinear(affine_coord(T.r0, X, T.r1, Y, T.r2), affine_coord(T.r3, X, T.r4, Y, T.r5), xread);
    };
    auto z = [&](int zy, int zx) -> float {     // translated image: (1*X + 0*Y) + t == X + t exactly
        return proj_bilinear(fadd((float)zx, T.tx), fadd((float)zy, T.ty), p);
    };
    int ylo, yhi, xlo, xhi;
    float yl, xl;
    resize_taps(i, (float)H / (float)h, H, ylo, yhi, yl);
    resize_taps(j, (float)W / (float)w, W, xlo, xhi, xl);
    const float tl = z(ylo, xlo), tr = z(ylo, xhi), bl = z(yhi, xlo), br = z(yhi, xhi);
    const float t = fadd(tl, fmul(fsub(tr, tl), xl));
    const float bb = fadd(bl, fmul(fsub(br, bl), xl));
    const float D = fadd(t, fmul(fsub(bb, t), yl));
    const float yk = copies[(((size_t)P.stack * N + src_idx[(size_t)b * N + ks]) * h + i) * w + j];
    resid[(((size_t)b * N + ks) * h + i) * wp + j] = fsub(D, yk);
}

template <bool WRITE_GRAD>
__global__ void __launch_bounds__(128)
kg_gradient_update(const float* __restrict__ x_cur, float* __restrict__ x_next, float* __restrict__ s0, float* __restrict__ s1,
                   float* __restrict__ s2, const float* __restrict__ resid, const InvXf* __restrict__ inv,
                   const ImgParams* __restrict__ ip, const Sched* __restrict__ sched, int it, int N, int h, int w, int wp, int H, int W,
                   int B, int S) {
    const int b = blockIdx.y;
    const ImgParams P = ip[b];
    if (it >= P.num_iter) return;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    const int Y = pix / W, X = pix - Y * W;
    const float Xf = (float)X, Yf = (float)Y;
    const int c_lo = S / 2 - 1;   // the two rows/columns of each SxS block that tf.image.resize samples (weights 0.5 each)
    float acc = 0.0f;
    for (int k = 0; k < P.n_kept; ++k) {
        const InvXf T = inv[(size_t)b * N + k];
        const float* r = resid + ((size_t)b * N + k) * h * wp;
        auto ghr = [&](long y, long xx) -> float {   // ResizeBilinearGrad(2*lambda_df*r): 0.25*g on the sampled positions, else 0
            if (y < 0 || y >= H || xx < 0 || xx >= W) return 0.0f;
            const int cy = (int)y / S, py = (int)y - cy * S, cx = (int)xx / S, px = (int)xx - cx * S;
            if ((py != c_lo && py != c_lo + 1) || (px != c_lo && px != c_lo + 1)) return 0.0f;
            return fmul(fmul(0.5f, fmul(P.two_ldf, r[cy * wp + cx])), 0.5f);
        };
        auto u = [&](long qy, long qx) -> float {    // warp-grad of the translate: the op with the inverted transform
            if (qy < 0 || qy >= H || qx < 0 || qx >= W) return 0.0f;
            return proj_bilinear(fadd((float)qx, T.ux), fadd((float)qy, T.uy), ghr);
        };
        acc = fadd(acc, proj_bilinear(affine_coord(T.b0, Xf, T.b1, Yf, T.b2), affine_coord(T.b3, Xf, T.b4, Yf, T.b5), u));
    }
    // regularisers and the optimizer step, as in k_gradient_update's epilogue
    const size_t plane = (size_t)H * W, i = (size_t)Y * W + X, gi = (size_t)b * plane + i;
    const float* xc = x_cur + (size_t)b * plane;
    const float xi = xc[i];
    float g = acc;
    if (P.use_btv) {
        for (int hh = -2; hh <= 2; ++hh) {
            for (int vv = 0; vv <= 2; ++vv) {
                const float lw = P.btv_lw[abs(hh) + vv];
                const int xs = X - hh, ys = Y - vv, xt = X + hh, yt = Y + vv;
                const float shifted = (xs >= 0 && xs < W && ys >= 0) ? xc[(size_t)ys * W + xs] : 0.0f;
                const float gd = fmul(lw, sgn(fsub(xi, shifted)));
                const float gb = (xt >= 0 && xt < W && yt < H) ? fmul(lw, sgn(fsub(xc[(size_t)yt * W + xt], xi))) : 0.0f;
                g = fadd(g, fsub(gd, gb));
            }
        }
    } else {
        if (Y > 0) g = fadd(g, fmul(P.lambda_tv, sgn(fsub(xi, xc[i - W]))));
        if (X > 0) g = fadd(g, fmul(P.lambda_tv, sgn(fsub(xi, xc[i - 1]))));
        if (Y < H - 1) g = fsub(g, fmul(P.lambda_tv, sgn(fsub(xc[i + W], xi))));
        if (X < W - 1) g = fsub(g, fmul(P.lambda_tv, sgn(fsub(xc[i + 1], xi))));
    }
    g = fadd(g, fmul(P.lambda_l2, fmul(xi, 2.0f)));
    if (P.lambda_l1 > 0.0f) g = fadd(g, fmul(P.lambda_l1, sgn(xi)));
    if (WRITE_GRAD) { x_next[gi] = g; return; }
    const Sched sc = sched[(size_t)it * B + b];
    const bool need0 = !(P.optimizer == ASR_OPT_SGD && P.momentum == 0.0f);
    const bool need1 = P.optimizer == ASR_OPT_ADAM || P.optimizer == ASR_OPT_ADADELTA || P.optimizer == ASR_OPT_ADAMAX;
    const bool need2 = P.optimizer == ASR_OPT_ADAM && P.amsgrad;
    x_next[gi] = optimizer_step(P, sc, g, xi, need0 ? s0[gi] : 0.0f, need1 ? s1[gi] : 0.0f, need2 ? s2[gi] : 0.0f, gi, s0, s1, s2);
}

// ================================================================================================
// loss of one evaluation (superresolution.py:71-98), double accumulators
// ================================================================================================
// at_iter < 0: the last evaluation of every image (x in buffer (num_iter-1)&1); at_iter >= 0: the evaluation of iteration at_iter
// (the verbose trace of superresolution.py:129-130), images that have already finished are skipped.
__global__ void k_loss_terms(const float* __restrict__ xa, const float* __restrict__ xb_, const float* __restrict__ resid,
                             const ImgParams* __restrict__ ip, double* __restrict__ accum, int N, int h, int w, int wp, int H, int W,
                             int at_iter) {
    const int b = blockIdx.y;
    const ImgParams P = ip[b];
    if (at_iter >= P.num_iter) return;
    const int ev = at_iter >= 0 ? at_iter : P.num_iter - 1;
    const float* x = ((ev & 1) ? xb_ : xa) + (size_t)b * H * W;
    const float* r = resid + (size_t)b * N * h * wp;
    const size_t nr = (size_t)P.n_kept * h * wp, nx = (size_t)H * W;
    double df = 0.0, tv = 0.0, l2 = 0.0, l1 = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += (size_t)gridDim.x * blockDim.x) {
        const double v = ((int)(i % wp) < w) ? r[i] : 0.0;   // pitch padding is not part of the residual
        df += v * v;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (size_t)gridDim.x * blockDim.x) {
        const int Y = (int)(i / W), X = (int)(i % W);
        const float xi = x[i];
        if (P.use_btv) {
            for (int hh = -2; hh <= 2; ++hh)
                for (int vv = 0; vv <= 2; ++vv) {
                    const int xs = X - hh, ys = Y - vv;
                    const float shifted = (xs >= 0 && xs < W && ys >= 0) ? x[(size_t)ys * W + xs] : 0.0f;
                    tv += (double)P.btv_w[abs(hh) + vv] * fabs((double)fsub(xi, shifted));
                }
        } else {
            if (Y < H - 1) tv += fabs((double)fsub(x[i + W], xi));
            if (X < W - 1) tv += fabs((double)fsub(x[i + 1], xi));
        }
        l2 += (double)xi * (double)xi;
        l1 += fabs((double)xi);
    }
    df = warp_sum(df); tv = warp_sum(tv); l2 = warp_sum(l2); l1 = warp_sum(l1);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(accum + 4 * b + 0, df);
        atomicAdd(accum + 4 * b + 1, tv);
        atomicAdd(accum + 4 * b + 2, l2);
        atomicAdd(accum + 4 * b + 3, l1);
    }
}

__global__ void k_loss_final(const double* __restrict__ accum, const AsrSolveParams* __restrict__ hp,
                             float* __restrict__ loss, int B, int stride, int at_iter) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const AsrSolveParams p = hp[b];
    if (at_iter >= p.num_iter) return;
    float l = fadd(fmul(p.lambda_df, (float)accum[4 * b]), fmul(p.lambda_tv, (float)accum[4 * b + 1]));
    l = fadd(l, fmul(p.lambda_l2, (float)accum[4 * b + 2]));
    if (p.lambda_l1 > 0.0f) l = fadd(l, fmul(p.lambda_l1, (float)accum[4 * b + 3]));
    loss[(size_t)b * stride] = l;
}

__global__ void k_select_output(const float* __restrict__ xa, const float* __restrict__ xb_, const ImgParams* __restrict__ ip,
                                float* __restrict__ out, size_t plane) {
    const int b = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(((ip[b].num_iter & 1) ? xb_ : xa) + (size_t)b * plane);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)b * plane);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane / 4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void k_fill(float* __restrict__ p, float v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ================================================================================================
// host orchestration
// ================================================================================================
static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
static int pitch4(int w) { return (w + 3) & ~3; }   // residual row pitch: TMA strides must be multiples of 16 bytes

struct Layout {
    size_t xa, xb, s0, s1, s2, resid, tapc, tapr, fwd, inv, src, ip, hp, sched, accum, fcp, fcolw, froww, boxd, total;
};
static int k1_tiles_x(int w) { return (w + K1_TJ - 1) / K1_TJ; }
static int k1_tiles_y(int h) { return (h + K1_TI - 1) / K1_TI; }

static Layout make_layout(int B, int N, int h, int w, int H, int W, int max_iter) {
    Layout L;
    const size_t plane = sizeof(float) * (size_t)B * H * W;
    size_t o = 0;
    L.xa = o; o += align_up(plane);
    L.xb = o; o += align_up(plane);
    L.s0 = o; o += align_up(plane);
    L.s1 = o; o += align_up(plane);
    L.s2 = o; o += align_up(plane);
    L.resid = o; o += align_up(sizeof(float) * (size_t)B * N * h * pitch4(w));
    L.tapc = o; o += align_up(sizeof(float2) * (size_t)B * N * 4 * (w + 2 * K2_TPAD));
    L.tapr = o; o += align_up(sizeof(float2) * (size_t)B * N * 4 * (h + 2 * K2_TPAD));
    L.fwd = o; o += align_up(sizeof(FwdXf) * (size_t)B * N);
    L.inv = o; o += align_up(sizeof(InvXf) * (size_t)B * N);
    L.src = o; o += align_up(sizeof(int) * (size_t)B * N);
    L.ip = o; o += align_up(sizeof(ImgParams) * (size_t)B);
    L.hp = o; o += align_up(sizeof(AsrSolveParams) * (size_t)B);
    L.sched = o; o += align_up(sizeof(Sched) * (size_t)B * (size_t)(max_iter > 0 ? max_iter : 1));
    L.accum = o; o += align_up(sizeof(double) * 4 * (size_t)B);
    L.fcp = o; o += align_up(sizeof(FwdCopy) * (size_t)B * N);
    L.fcolw = o; o += align_up(sizeof(float4) * (size_t)B * N * w);
    L.froww = o; o += align_up(sizeof(float4) * (size_t)B * N * h);
    L.boxd = o; o += align_up(sizeof(BoxDesc) * (size_t)B * N * k1_tiles_x(w) * k1_tiles_y(h));
    L.total = o;
    return L;
}

static int check_shapes(int B, int N, int h, int w, int H, int W) {
    if (B <= 0 || N <= 0 || h <= 0 || w <= 0) return fail(ASR_EINVAL, "B, N, h, w must be positive (got %d %d %d %d)", B, N, h, w);
    if (H <= 0 || W <= 0 || H % h != 0 || W % w != 0 || H / h != W / w || (H / h) % 2 != 0)
        return fail(ASR_EUNSUPPORTED, "output_size must be feature_size times an even integer (got %dx%d -> %dx%d)", h, w, H, W);
    if (B > 65535 || N > 65535) return fail(ASR_EINVAL, "B and N must be <= 65535");
    if (H > 16384 || W > 16384) return fail(ASR_EINVAL, "output larger than 16384 pixels per side (fp32 address arithmetic of the gathers)");
    if ((double)B * N * (double)(h > w ? h : w) * 16.0 >= 4294967296.0 || (double)N * h * w >= 2147483648.0)
        return fail(ASR_EINVAL, "B*N*max(h,w) must stay below 2^28 and N*h*w below 2^31 (32-bit table offsets)");
    return ASR_OK;
}

static int check_params(const AsrSolveParams* p, int n) {
    for (int i = 0; i < n; ++i) {
        if (p[i].optimizer < ASR_OPT_ADAM || p[i].optimizer > ASR_OPT_ADAMAX) return fail(ASR_EINVAL, "unknown optimizer %d", p[i].optimizer);
        if (p[i].num_iter < 0) return fail(ASR_EINVAL, "num_iter < 0");
    }
    return ASR_OK;
}

// ExponentialDecay (optimizer.py:43-52): lr0 * rate^(i/steps), fp32, non-staircase
static float lr_at(const AsrSolveParams& p, int i) {
    if (!p.lr_scheduler) return p.learning_rate;
    const float e = (float)i / p.decay_steps;
    return p.learning_rate * powf(p.decay_rate, e);
}

struct HostTables {
    std::vector<FwdXf> fwd;
    std::vector<InvXf> inv;
    std::vector<int> src;
    std::vector<ImgParams> ip;
    std::vector<AsrSolveParams> hp;
    std::vector<Sched> sched;
    int max_iter = 0, max_kept = 0;
    bool any_btv = false;
    bool small_box = true;   // every copy's K1 source box fits K1_XR_SMALL rows
};

static void build_tables(const AsrSolveParams* params, int n_params, const float* angles, const float* shifts,
                         const uint8_t* keep, const int32_t* stack_index, int B, int N, int H, int W, HostTables& T) {
    T.fwd.assign((size_t)B * N, FwdXf{});
    T.inv.assign((size_t)B * N, InvXf{});
    T.src.assign((size_t)B * N, 0);
    T.ip.resize(B);
    T.hp.resize(B);
    for (int b = 0; b < B; ++b) {
        const AsrSolveParams& p = params[n_params == 1 ? 0 : b];
        T.hp[b] = p;
        const size_t sb = stack_index ? (size_t)stack_index[b] : (size_t)b;   // angles/shifts/copies are per stack
        int kept = 0;
        for (int k = 0; k < N; ++k) {
            if (keep && !keep[(size_t)b * N + k]) continue;
            float rot[8], roti[8], tr[8], tri[8];
            rotate_matrix(angles[sb * N + k], H, W, rot);
            invert_transform(rot, roti);
            tr[0] = 1.0f; tr[1] = 0.0f; tr[2] = -shifts[2 * (sb * N + k)];
            tr[3] = 0.0f; tr[4] = 1.0f; tr[5] = -shifts[2 * (sb * N + k) + 1];
            tr[6] = 0.0f; tr[7] = 0.0f;
            invert_transform(tr, tri);
            const size_t o = (size_t)b * N + kept;
            T.fwd[o] = FwdXf{rot[0], rot[1], rot[2], rot[3], rot[4], rot[5], tr[2], tr[5]};
            // K1 box rows <= 62|t3| + 62|t4| + 3 (corner span of the 63x63 p region, the +1 tap, floor); small margin
            if ((float)K1_SPAN_X * fabsf(rot[3]) + (float)K1_SPAN_Y * fabsf(rot[4]) + 3.05f > (float)K1_XR_SMALL) T.small_box = false;
            T.inv[o] = InvXf{roti[0], roti[1], roti[2], roti[3], roti[4], roti[5], tri[2], tri[5]};
            T.src[o] = k;
            ++kept;
        }
        ImgParams q{};
        q.two_ldf = 2.0f * p.lambda_df;
        q.lambda_tv = p.lambda_tv; q.lambda_l2 = p.lambda_l2; q.lambda_l1 = p.lambda_l1;
        q.omb1 = 1.0f - p.beta_1; q.omb2 = 1.0f - p.beta_2; q.beta_2 = p.beta_2;
        q.epsilon = p.epsilon; q.momentum = p.momentum;
        q.optimizer = p.optimizer; q.amsgrad = p.amsgrad; q.nesterov = p.nesterov;
        q.num_iter = p.num_iter; q.n_kept = kept;
        q.use_btv = p.use_btv ? 1 : 0;
        q.stack = (int)sb;
        if (p.use_btv) T.any_btv = true;
        for (int n = 0; n < 5; ++n) {
            q.btv_w[n] = powf(0.6f, (float)n);
            q.btv_lw[n] = p.lambda_tv * q.btv_w[n];
        }
        T.ip[b] = q;
        if (p.num_iter > T.max_iter) T.max_iter = p.num_iter;
        if (kept > T.max_kept) T.max_kept = kept;
    }
    T.sched.assign((size_t)B * (T.max_iter > 0 ? T.max_iter : 1), make_float2(0.f, 0.f));
    for (int b = 0; b < B; ++b) {
        const AsrSolveParams& p = T.hp[b];
        for (int i = 0; i < p.num_iter; ++i) {
            const float lr = lr_at(p, i);
            const float t = (float)(p.step_offset + (int64_t)i + 1);   // optimizer.iterations + 1
            float y = 0.0f;
            if (p.optimizer == ASR_OPT_ADAM) {
                const float b1p = powf(p.beta_1, t), b2p = powf(p.beta_2, t);
                y = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
            } else if (p.optimizer == ASR_OPT_ADAMAX) {
                const float b1p = powf(p.beta_1, t);
                y = lr / (1.0f - b1p);
            }
            T.sched[(size_t)i * B + b] = make_float2(lr, y);
        }
    }
}

struct Device {
    float *xa, *xb, *s0, *s1, *s2, *resid;
    float2 *tapc, *tapr;
    FwdXf* fwd; InvXf* inv; int* src; ImgParams* ip; AsrSolveParams* hp; Sched* sched; double* accum;
    FwdCopy* fcp; float4* fcolw; float4* froww; BoxDesc* boxd;
};

static Device bind(void* ws, const Layout& L) {
    unsigned char* p = static_cast<unsigned char*>(ws);
    Device D;
    D.xa = (float*)(p + L.xa); D.xb = (float*)(p + L.xb);
    D.s0 = (float*)(p + L.s0); D.s1 = (float*)(p + L.s1); D.s2 = (float*)(p + L.s2);
    D.resid = (float*)(p + L.resid);
    D.tapc = (float2*)(p + L.tapc); D.tapr = (float2*)(p + L.tapr);
    D.fwd = (FwdXf*)(p + L.fwd); D.inv = (InvXf*)(p + L.inv); D.src = (int*)(p + L.src);
    D.ip = (ImgParams*)(p + L.ip); D.hp = (AsrSolveParams*)(p + L.hp); D.sched = (Sched*)(p + L.sched);
    D.accum = (double*)(p + L.accum);
    D.fcp = (FwdCopy*)(p + L.fcp); D.fcolw = (float4*)(p + L.fcolw); D.froww = (float4*)(p + L.froww); D.boxd = (BoxDesc*)(p + L.boxd);
    return D;
}

static int upload(const HostTables& T, const Device& D, cudaStream_t st) {
    ASR_CUDA_TRY(cudaMemcpyAsync(D.fwd, T.fwd.data(), sizeof(FwdXf) * T.fwd.size(), cudaMemcpyHostToDevice, st));
    ASR_CUDA_TRY(cudaMemcpyAsync(D.inv, T.inv.data(), sizeof(InvXf) * T.inv.size(), cudaMemcpyHostToDevice, st));
    ASR_CUDA_TRY(cudaMemcpyAsync(D.src, T.src.data(), sizeof(int) * T.src.size(), cudaMemcpyHostToDevice, st));
    ASR_CUDA_TRY(cudaMemcpyAsync(D.ip, T.ip.data(), sizeof(ImgParams) * T.ip.size(), cudaMemcpyHostToDevice, st));
    ASR_CUDA_TRY(cudaMemcpyAsync(D.hp, T.hp.data(), sizeof(AsrSolveParams) * T.hp.size(), cudaMemcpyHostToDevice, st));
    ASR_CUDA_TRY(cudaMemcpyAsync(D.sched, T.sched.data(), sizeof(Sched) * T.sched.size(), cudaMemcpyHostToDevice, st));
    return ASR_OK;
}

// 3-D tiled tensor map over an x buffer [B][H][W] fp32 with the K1 box; out-of-bounds elements read as zero
static int make_x_map(CUtensorMap* map, const float* base, int B, int H, int W, int box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ASR_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(ASR_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstride[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
    const cuuint32_t box[3] = {K1_XS, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASR_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ASR_OK;
}

// 3-D tiled tensor map over the residuals [B*N][h][pitch] fp32 with the K2 fill box; cells outside the LR grid read as zero
static int make_r_map(CUtensorMap* map, const float* base, int BN, int h, int w) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ASR_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(ASR_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const int wp = pitch4(w);
    const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)BN};
    const cuuint64_t gstride[2] = {(cuuint64_t)wp * sizeof(float), (cuuint64_t)wp * h * sizeof(float)};
    const cuuint32_t box[3] = {K2_RW, K2_BC, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box,
                                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ASR_ECUDA, "cuTensorMapEncodeTiled (residuals) failed with CUresult %d", (int)r);
    return ASR_OK;
}

static unsigned div_magic(int d) { return (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d); }   // __umulhi(n, magic) == n / d for n*d < 2^32

static int configure_kernels() {
    static unsigned long long done = 0;   // one bit per device: the attribute belongs to the (function, device) pair
    if (!first_use_on_device(&done)) return ASR_OK;
    ASR_CUDA_TRY(cudaFuncSetAttribute(k_forward_residual<K1_XR_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem<K1_XR_SMALL>()));
    ASR_CUDA_TRY(cudaFuncSetAttribute(k_forward_residual<K1_XR_SMALL>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // 6 CTAs need 227 of the 228 KB
    ASR_CUDA_TRY(cudaFuncSetAttribute(k_forward_residual<K1_XR_BIG>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ASR_CUDA_TRY(cudaFuncSetAttribute(k_forward_residual<K1_XR_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem<K1_XR_BIG>()));
#define ASR_K2_ATTR(WG, BT, TY) \
    ASR_CUDA_TRY(cudaFuncSetAttribute(k_gradient_update<WG, BT, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(k2_smem<TY>())));
#define ASR_K2_ATTRS(TY) ASR_K2_ATTR(false, false, TY) ASR_K2_ATTR(false, true, TY) ASR_K2_ATTR(true, false, TY) ASR_K2_ATTR(true, true, TY)
    ASR_K2_ATTRS(64) ASR_K2_ATTRS(32)
#undef ASR_K2_ATTRS
#undef ASR_K2_ATTR
    return ASR_OK;
}

// loss of images [b0, b0+nb): the last evaluation (at_iter < 0) into d_loss[b], or the evaluation of iteration at_iter into
// d_loss[b * stride] (d_loss already points at the trace column)
static int launch_loss(const Device& D, float* d_loss, int b0, int nb, int N, int h, int w, int H, int W, cudaStream_t st,
                       int stride = 1, int at_iter = -1) {
    const size_t plane = (size_t)H * W;
    ASR_CUDA_TRY(cudaMemsetAsync(D.accum + 4 * (size_t)b0, 0, sizeof(double) * 4 * nb, st));
    ASR_LAUNCH(k_loss_terms, dim3(64, nb), 256, 0, st, D.xa + b0 * plane, D.xb + b0 * plane, D.resid + (size_t)b0 * N * h * pitch4(w), D.ip + b0,
               D.accum + 4 * (size_t)b0, N, h, w, pitch4(w), H, W, at_iter);
    ASR_LAUNCH(k_loss_final, (nb + 127) / 128, 128, 0, st, D.accum + 4 * (size_t)b0, D.hp + b0, d_loss + (size_t)b0 * stride, nb, stride, at_iter);   // D.hp: one entry per image
    return ASR_OK;
}

}  // namespace asr

using namespace asr;

#ifdef ASR_K2_TRACE
extern "C" int asr_debug_k2_trace(long long* h_trace, long long* h_misc) {
    ASR_CUDA_TRY(cudaDeviceSynchronize());
    ASR_CUDA_TRY(cudaMemcpyFromSymbol(h_trace, g_k2_trace, sizeof(long long) * 16 * 128 * 4));
    ASR_CUDA_TRY(cudaMemcpyFromSymbol(h_misc, g_k2_misc, sizeof(long long) * 16 * 4));
    return ASR_OK;
}
#endif

extern "C" int asr_solve_workspace_bytes(int B, int N, int h, int w, int H, int W, int max_iter, size_t* bytes) {
    if (!bytes) return fail(ASR_ENULL, "bytes is NULL");
    if (int e = check_shapes(B, N, h, w, H, W)) return e;
    *bytes = make_layout(B, N, h, w, H, W, max_iter).total;
    return ASR_OK;
}

// K2 tile height: 64 rows unless the whole solve fits one 64x32 CTA per SM (a 512^2 image is only 64 tiles of 64x64: a single
// image leaves most SMs idle and every CTA latency-bound; measured K2 104 -> 73 us per iteration with 32-row tiles and 8 fill
// warps).  16-row tiles and four u buffers were measured too and never win.  ASR_K2_TY overrides the choice (tests, experiments).
static int k2_tile_height(int n_images, int H, int W) {
    if (const char* e = getenv("ASR_K2_TY")) { const int v = atoi(e); if (v == 64 || v == 32) return v; }
    int n_sm = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles32 = (long long)((W + K2_T - 1) / K2_T) * ((H + 31) / 32);
    return (long long)n_images * tiles32 <= n_sm ? 32 : 64;   // the 32-row variant runs one CTA per SM
}
#define ASR_LAUNCH_K2_TY(WG, BT, TY, pdl, ntiles, nimg, st, ...) \
    ASR_LAUNCH_TIMED_PDL(1, pdl, (k_gradient_update<WG, BT, TY>), dim3(ntiles, nimg), K2Fill<TY>::threads, (k2_smem<TY>()), st, __VA_ARGS__)
#define ASR_LAUNCH_K2(WG, btv, ty, pdl, H, W, nimg, st, ...)                                                          \
    do {                                                                                                              \
        const int ntiles_ = ((W + K2_T - 1) / K2_T) * ((H + (ty) - 1) / (ty));                                        \
        if (btv) {                                                                                                    \
            if ((ty) == 64) ASR_LAUNCH_K2_TY(WG, true, 64, pdl, ntiles_, nimg, st, __VA_ARGS__);                      \
            else ASR_LAUNCH_K2_TY(WG, true, 32, pdl, ntiles_, nimg, st, __VA_ARGS__);                                 \
        } else {                                                                                                      \
            if ((ty) == 64) ASR_LAUNCH_K2_TY(WG, false, 64, pdl, ntiles_, nimg, st, __VA_ARGS__);                     \
            else ASR_LAUNCH_K2_TY(WG, false, 32, pdl, ntiles_, nimg, st, __VA_ARGS__);                                \
        }                                                                                                             \
    } while (0)

#define ASR_LAUNCH_K1(small, pdl, t1, nk, nimg, st, ...)                                                                                \
    do {                                                                                                                                \
        if (small) ASR_LAUNCH_TIMED_PDL(0, pdl, k_forward_residual<K1_XR_SMALL>, dim3(t1, nk, nimg), K1_THREADS, k1_smem<K1_XR_SMALL>(), st, __VA_ARGS__); \
        else ASR_LAUNCH_TIMED_PDL(0, pdl, k_forward_residual<K1_XR_BIG>, dim3(t1, nk, nimg), K1_THREADS, k1_smem<K1_XR_BIG>(), st, __VA_ARGS__);           \
    } while (0)

// Programmatic dependent launch inside the iteration loop (profiles/r02_pdl.txt).  Every K1 launch may overlap the previous K2's tail and
// every K2 launch the previous K1's: the set-up of the first wave (descriptor, record and operand loads; chunk prologue, barrier set-up,
// L2 prefetch) runs while the other kernel drains.  K2 lets K1 in right after its own wait.  K1 does that only in the one-image variant
// (32-row K2 tiles: a K2 CTA needs a whole SM and cannot displace K1's CTAs); otherwise K2 is let in when K1's CTAs exit -- with an early
// trigger the 64-row K2 CTAs of a two-image solve became resident at once and held half of every SM while K1 still had most of its
// grid to run: 136 -> 196 us per iteration.  One image -5.7 % per iteration, 2..8 images -3.6..-1.2 %, 64 images +-0; never slower.
// ASR_PDL (experiments): bit 0 = K1 launches, bit 1 = K2 launches; default 3
static int pdl_mask() {
    const char* e = getenv("ASR_PDL");   // read per solve, like ASR_K2_TY: tests switch it inside one process
    return e ? atoi(e) : 3;
}

static int solve_impl(const AsrSolveParams* params, int n_params, const float* d_copies, const float* h_angles,
                      const float* h_shifts, const uint8_t* h_keep, const int32_t* h_stack_index, int B, int N, int h, int w,
                      int H, int W, float* d_x_out, float* d_loss_out, void* d_workspace, size_t workspace_bytes, void* stream,
                      int loss_every = 0, float* d_loss_trace = nullptr, int trace_cols = 0) {
    if (!params || !d_copies || !h_angles || !h_shifts || !d_x_out || !d_workspace) return fail(ASR_ENULL, "null argument");
    if (n_params != 1 && n_params != B) return fail(ASR_EINVAL, "n_params must be 1 or B");
    if (int e = check_shapes(B, N, h, w, H, W)) return e;
    if (int e = check_params(params, n_params)) return e;
    if (loss_every < 0 || (loss_every > 0 && (!d_loss_trace || trace_cols <= 0))) return fail(ASR_EINVAL, "loss trace needs loss_every > 0, a buffer and trace_cols > 0");
    if (!aligned16(d_copies) || !aligned16(d_x_out) || (reinterpret_cast<uintptr_t>(d_workspace) & 255u))
        return fail(ASR_EINVAL, "d_copies / d_x_out must be 16-byte aligned and d_workspace 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    HostTables T;
    build_tables(params, n_params, h_angles, h_shifts, h_keep, h_stack_index, B, N, H, W, T);
    const Layout L = make_layout(B, N, h, w, H, W, T.max_iter);
    if (workspace_bytes < L.total) return fail(ASR_EWORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, L.total);
    const Device D = bind(d_workspace, L);
    if (int e = configure_kernels()) return e;
    if (int e = upload(T, D, st)) return e;

    const size_t plane = (size_t)H * W;
    ASR_CUDA_TRY(cudaMemsetAsync(D.s0, 0, L.resid - L.s0, st));   // optimizer slots s0,s1,s2 start at zero
    for (int b = 0; b < B; ++b)   // Adagrad slots start at initial_accumulator_value (optimizer.py:25-27)
        if (T.hp[b].optimizer == ASR_OPT_ADAGRAD)
            ASR_LAUNCH(k_fill, 64, 256, 0, st, D.s0 + (size_t)b * plane, T.hp[b].initial_accumulator_value, plane);
    ASR_LAUNCH(k_init_upsample, dim3((W + 31) / 32, (H + 7) / 8, B), dim3(32, 8), 0, st, d_copies, D.ip, D.xa, N, h, w, H, W);

    const int wp = pitch4(w);
    // verbose trace (superresolution.py:129-130): the loss of iteration `it`, evaluated between the forward and the update kernel
    auto trace_loss = [&](int b0, int nb, int it) -> int {
        if (loss_every <= 0 || it % loss_every != 0 || it / loss_every >= trace_cols) return ASR_OK;
        return launch_loss(D, d_loss_trace + it / loss_every, b0, nb, N, h, w, H, W, st, trace_cols, it);
    };
    if (H != 4 * h) {
        // any other even ratio (x2, x6, x8 ...): the literal one-thread-per-output kernels
        const int S = H / h;
        for (int it = 0; it < T.max_iter; ++it) {
            float* xc = (it & 1) ? D.xb : D.xa;
            float* xn = (it & 1) ? D.xa : D.xb;
            ASR_LAUNCH_TIMED(0, kg_forward_residual, dim3((h * w + 127) / 128, T.max_kept, B), 128, 0, st, xc, d_copies, D.resid, D.fwd, D.src,
                             D.ip, it, N, h, w, wp, H, W);
            if (int e = trace_loss(0, B, it)) return e;
            ASR_LAUNCH_TIMED(1, kg_gradient_update<false>, dim3((H * W + 127) / 128, B), 128, 0, st, xc, xn, D.s0, D.s1, D.s2, D.resid, D.inv,
                             D.ip, D.sched, it, N, h, w, wp, H, W, B, S);
        }
        ASR_CUDA_TRY(cudaGetLastError());
        if (d_loss_out) {
            if (int e = launch_loss(D, d_loss_out, 0, B, N, h, w, H, W, st)) return e;
        }
        ASR_LAUNCH(k_select_output, dim3(32, B), 256, 0, st, D.xa, D.xb, D.ip, d_x_out, plane);
        ASR_CUDA_TRY(cudaGetLastError());
        return ASR_OK;
    }
    const int ntj = k1_tiles_x(w), nti = k1_tiles_y(h);
    const int t1 = ntj * nti;
    CUtensorMap map_a, map_b, map_r;
    const int box_rows = T.small_box ? K1_XR_SMALL : K1_XR_BIG;
    if (int e = make_x_map(&map_a, D.xa, B, H, W, box_rows)) return e;
    if (int e = make_x_map(&map_b, D.xb, B, H, W, box_rows)) return e;
    if (int e = make_r_map(&map_r, D.resid, B * N, h, w)) return e;
    // everything that depends on the transforms only, once per solve
    ASR_LAUNCH(k_tap_tables, dim3((4 * (w + h + 4 * K2_TPAD) + 255) / 256, N, B), 256, 0, st, D.inv, D.tapc, D.tapr, N, h, w, H, W);
    ASR_LAUNCH(k_forward_tables, dim3((t1 + w + h + 127) / 128, T.max_kept, B), 128, 0, st, D.fwd, D.src, D.ip, D.fcp, D.fcolw, D.froww,
               D.boxd, N, h, w, H, W, ntj, nti, box_rows);
    // one launch group = the images [b0, b0+nb) that go through every kernel launch together
    struct Group { int b0, nb, iters, min_iters, ty; bool uniform_kept; };
    auto make_group = [&](int b0, int nb) {
        Group G{b0, nb, 0, INT_MAX, k2_tile_height(nb, H, W), true};   // uniform_kept: every image keeps max_kept copies, no K1 CTA is idle
        for (int b = b0; b < b0 + nb; ++b) {
            G.iters = T.hp[b].num_iter > G.iters ? T.hp[b].num_iter : G.iters;
            G.min_iters = T.hp[b].num_iter < G.min_iters ? T.hp[b].num_iter : G.min_iters;
            G.uniform_kept = G.uniform_kept && T.ip[b].n_kept == T.max_kept;
        }
        return G;
    };
    // pdl: the launch directly follows the other solve kernel of the same group in the stream (no loss trace, no profile mark in between)
    auto launch_k1 = [&](const Group& G, int it, cudaStream_t s, bool pdl) -> int {
        const size_t ro = (size_t)G.b0 * N * h * wp, so = (size_t)G.b0 * N;
        ASR_LAUNCH_K1(T.small_box, pdl, t1, T.max_kept, G.nb, s, (it & 1) ? map_b : map_a, d_copies, D.resid + ro, D.fcp + so, D.fcolw + so * w,
                      D.froww + so * h, D.boxd + so * t1, D.ip + G.b0, it, ((!G.uniform_kept || it >= G.min_iters) ? 1 : 0) | (G.ty == 32 ? 2 : 0), N, h, w, wp, ntj,
                      div_magic(ntj), G.b0);
        return ASR_OK;
    };
    auto launch_k2 = [&](const Group& G, int it, cudaStream_t s, bool pdl) -> int {
        const size_t po = (size_t)G.b0 * plane;
        float* xc = ((it & 1) ? D.xb : D.xa) + po;
        float* xn = ((it & 1) ? D.xa : D.xb) + po;
        ASR_LAUNCH_K2(false, T.any_btv, G.ty, pdl, H, W, G.nb, s, map_r, xc, xn, D.s0 + po, D.s1 + po, D.s2 + po, D.tapc, D.tapr,
                      D.inv + (size_t)G.b0 * N, D.ip + G.b0, D.sched + G.b0, it, N, h, w, H, W, B, G.b0);
        return ASR_OK;
    };
    int group = params[0].images_in_flight > 0 ? params[0].images_in_flight : B;
    // (Two halves of a group on two streams, half an iteration apart, so that K1 of one half shares the SMs with K2 of the other:
    //  measured +1.3 % with the kernels as they are and -16 % when K2 is held to one CTA per SM to make room -- both kernels are
    //  issue-heavy, there is little idle capacity to trade.  scripts/dev/dual_stream_experiment.sh, DESIGN.md.)
    const int pdl = profile_enabled() ? 0 : pdl_mask();   // profile marks are event records between the launches
    for (int b0 = 0; b0 < B; b0 += group) {
        const int nb = (B - b0 < group) ? B - b0 : group;
        const Group G = make_group(b0, nb);
        for (int it = 0; it < G.iters; ++it) {
            const bool traced = loss_every > 0 && it % loss_every == 0;   // a loss kernel sits between K1 and K2 of this iteration
            if (int e = launch_k1(G, it, st, (pdl & 1) && it > 0)) return e;
            if (int e = trace_loss(b0, nb, it)) return e;
            if (int e = launch_k2(G, it, st, (pdl & 2) && !traced)) return e;
        }
    }
    ASR_CUDA_TRY(cudaGetLastError());
    if (d_loss_out) {
        if (int e = launch_loss(D, d_loss_out, 0, B, N, h, w, H, W, st)) return e;
    }
    ASR_LAUNCH(k_select_output, dim3(32, B), 256, 0, st, D.xa, D.xb, D.ip, d_x_out, plane);
    ASR_CUDA_TRY(cudaGetLastError());
    return ASR_OK;
}

extern "C" int asr_solve_batched(const AsrSolveParams* params, int n_params, const float* d_copies,
                                 const float* h_angles, const float* h_shifts, const uint8_t* h_keep, int B, int N,
                                 int h, int w, int H, int W, float* d_x_out, float* d_loss_out, void* d_workspace,
                                 size_t workspace_bytes, void* stream) {
    return solve_impl(params, n_params, d_copies, h_angles, h_shifts, h_keep, nullptr, B, N, h, w, H, W, d_x_out, d_loss_out,
                      d_workspace, workspace_bytes, stream);
}

extern "C" int asr_solve_batched_traced(const AsrSolveParams* params, int n_params, const float* d_copies, const float* h_angles,
                                        const float* h_shifts, const uint8_t* h_keep, int B, int N, int h, int w, int H, int W,
                                        float* d_x_out, float* d_loss_out, int loss_every, float* d_loss_trace, int trace_cols,
                                        void* d_workspace, size_t workspace_bytes, void* stream) {
    return solve_impl(params, n_params, d_copies, h_angles, h_shifts, h_keep, nullptr, B, N, h, w, H, W, d_x_out, d_loss_out,
                      d_workspace, workspace_bytes, stream, loss_every, d_loss_trace, trace_cols);
}

extern "C" int asr_solve_sweep(const AsrSolveParams* params, int n_points, const float* d_copies, const float* h_angles,
                               const float* h_shifts, const int32_t* h_stack_index, int n_stacks, int N, int h, int w, int H,
                               int W, float* d_x_out, float* d_loss_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!h_stack_index) return fail(ASR_ENULL, "h_stack_index is NULL");
    for (int i = 0; i < n_points; ++i)
        if (h_stack_index[i] < 0 || h_stack_index[i] >= n_stacks) return fail(ASR_EINVAL, "stack index %d of point %d outside [0,%d)", h_stack_index[i], i, n_stacks);
    return solve_impl(params, n_points, d_copies, h_angles, h_shifts, nullptr, h_stack_index, n_points, N, h, w, H, W, d_x_out,
                      d_loss_out, d_workspace, workspace_bytes, stream);
}

extern "C" int asr_loss_grad_batched(const AsrSolveParams* params, int n_params, const float* d_x, const float* d_copies,
                                     const float* h_angles, const float* h_shifts, const uint8_t* h_keep, int B, int N,
                                     int h, int w, int H, int W, float* d_resid, float* d_grad, float* d_loss_out,
                                     void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!params || !d_x || !d_copies || !h_angles || !h_shifts || !d_workspace) return fail(ASR_ENULL, "null argument");
    if (n_params != 1 && n_params != B) return fail(ASR_EINVAL, "n_params must be 1 or B");
    if (int e = check_shapes(B, N, h, w, H, W)) return e;
    if (int e = check_params(params, n_params)) return e;
    if (!aligned16(d_copies) || !aligned16(d_x) || (reinterpret_cast<uintptr_t>(d_workspace) & 255u))
        return fail(ASR_EINVAL, "d_copies / d_x must be 16-byte aligned and d_workspace 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    std::vector<AsrSolveParams> one(params, params + n_params);
    for (auto& p : one) p.num_iter = 1;   // a single evaluation at the supplied x
    HostTables T;
    build_tables(one.data(), n_params, h_angles, h_shifts, h_keep, nullptr, B, N, H, W, T);
    const Layout L = make_layout(B, N, h, w, H, W, 1);
    if (workspace_bytes < L.total) return fail(ASR_EWORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, L.total);
    const Device D = bind(d_workspace, L);
    if (int e = configure_kernels()) return e;
    if (int e = upload(T, D, st)) return e;

    const size_t plane = (size_t)H * W;
    ASR_CUDA_TRY(cudaMemcpyAsync(D.xa, d_x, sizeof(float) * B * plane, cudaMemcpyDeviceToDevice, st));
    const int wp = pitch4(w);
    if (H != 4 * h) {
        ASR_LAUNCH_TIMED(0, kg_forward_residual, dim3((h * w + 127) / 128, T.max_kept, B), 128, 0, st, D.xa, d_copies, D.resid, D.fwd, D.src,
                         D.ip, 0, N, h, w, wp, H, W);
        ASR_LAUNCH_TIMED(1, kg_gradient_update<true>, dim3((H * W + 127) / 128, B), 128, 0, st, D.xa, D.xb, D.s0, D.s1, D.s2, D.resid, D.inv,
                         D.ip, D.sched, 0, N, h, w, wp, H, W, B, H / h);
    } else {
    const int ntj = k1_tiles_x(w), nti = k1_tiles_y(h);
    const int t1 = ntj * nti;
    CUtensorMap map_a, map_r;
    const int box_rows = T.small_box ? K1_XR_SMALL : K1_XR_BIG;
    if (int e = make_x_map(&map_a, D.xa, B, H, W, box_rows)) return e;
    if (int e = make_r_map(&map_r, D.resid, B * N, h, w)) return e;
    ASR_LAUNCH(k_tap_tables, dim3((4 * (w + h + 4 * K2_TPAD) + 255) / 256, N, B), 256, 0, st, D.inv, D.tapc, D.tapr, N, h, w, H, W);
    ASR_LAUNCH(k_forward_tables, dim3((t1 + w + h + 127) / 128, T.max_kept, B), 128, 0, st, D.fwd, D.src, D.ip, D.fcp, D.fcolw, D.froww,
               D.boxd, N, h, w, H, W, ntj, nti, box_rows);
    ASR_LAUNCH_K1(T.small_box, false, t1, T.max_kept, B, st, map_a, d_copies, D.resid, D.fcp, D.fcolw, D.froww, D.boxd, D.ip, 0, 1, N, h, w, wp, ntj,
                  div_magic(ntj), 0);
    ASR_LAUNCH_K2(true, T.any_btv, k2_tile_height(B, H, W), false, H, W, B, st, map_r, D.xa, D.xb, D.s0, D.s1, D.s2, D.tapc, D.tapr, D.inv, D.ip,
                  D.sched, 0, N, h, w, H, W, B, 0);
    }
    ASR_CUDA_TRY(cudaGetLastError());
    if (d_grad) ASR_CUDA_TRY(cudaMemcpyAsync(d_grad, D.xb, sizeof(float) * B * plane, cudaMemcpyDeviceToDevice, st));
    if (d_resid) {
        // workspace residuals are indexed by kept slot; scatter back to copy order (dropped copies -> 0)
        ASR_CUDA_TRY(cudaMemsetAsync(d_resid, 0, sizeof(float) * (size_t)B * N * h * w, st));
        for (int b = 0; b < B; ++b)
            for (int s = 0; s < T.ip[b].n_kept; ++s)
                ASR_CUDA_TRY(cudaMemcpy2DAsync(d_resid + ((size_t)b * N + T.src[(size_t)b * N + s]) * h * w, sizeof(float) * w,
                                               D.resid + ((size_t)b * N + s) * h * wp, sizeof(float) * wp, sizeof(float) * w, h,
                                               cudaMemcpyDeviceToDevice, st));
    }
    if (d_loss_out) {
        if (int e = launch_loss(D, d_loss_out, 0, B, N, h, w, H, W, st)) return e;
    }
    return ASR_OK;
}
