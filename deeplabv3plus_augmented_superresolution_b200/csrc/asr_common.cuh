// asr_common.cuh -- shared device/host helpers for libasr (sm_100a only).
//
// Numerical contract (DESIGN.md "Exactness"): every kernel on the solve path evaluates the
// TensorFlow operators in the same fp32 order as the literal op sequence, with NO fused
// multiply-add, so that results are bit-identical to the un-fused IEEE evaluation.  All arithmetic
// that feeds a result therefore goes through the explicit round-to-nearest intrinsics below; the
// library is also compiled with --fmad=false as a second line of defence.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/asr.h"

namespace asr {

// ---- error plumbing ---------------------------------------------------------------------------
extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);

#define ASR_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return ::asr::fail(ASR_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// Attributes set with cudaFuncSetAttribute belong to the (function, device) pair, so "done once" is tracked per
// device: returns true the first time it is called with `mask` on the current device (bit = device ordinal;
// ordinals >= 64 simply repeat the cheap call every time).  A benign race at worst repeats the call.
bool first_use_on_device(unsigned long long* mask);
// device pointers handed to the library must be 16-byte aligned (float4 / TMA / 128-bit stores); see include/asr.h
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// stream-ordered scratch that is released on every return path
struct AsyncScratch {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t s) { st = s; return cudaMallocAsync(&p, bytes, s); }
    ~AsyncScratch() { if (p) cudaFreeAsync(p, st); }
};

// ---- launch accounting / optional per-kernel timing (asr_kernel_launches, asr_profile_*) -------
// Every kernel launch of the library goes through ASR_LAUNCH so that bench.py can report how many of
// our kernels ran in its timed region; with profiling on, the two solve kernels are bracketed by
// CUDA events on the launching stream (summed by asr_profile_read after a stream sync).
void count_launch();
bool profile_enabled();    // asr_profile_enable(1): event records sit between the launches, so the solve launches without PDL
void profile_mark(int slot, cudaStream_t st, bool begin);   // slot 0 = forward residual, 1 = gradient/update
#define ASR_LAUNCH(kernel, grid, block, smem, st, ...)          \
    do {                                                        \
        ::asr::count_launch();                                  \
        kernel<<<grid, block, smem, st>>>(__VA_ARGS__);         \
    } while (0)
#define ASR_LAUNCH_TIMED(slot, kernel, grid, block, smem, st, ...) \
    do {                                                           \
        ::asr::profile_mark(slot, st, true);                       \
        ASR_LAUNCH(kernel, grid, block, smem, st, __VA_ARGS__);    \
        ::asr::profile_mark(slot, st, false);                      \
    } while (0)

// Programmatic dependent launch (the solve's K1 -> K2 -> K1 chain): with `pdl` set the kernel may be scheduled while the previous kernel
// of the stream is still draining; it runs its own set-up and blocks in pdl_wait() until that kernel has completed and its writes are
// visible.  Without the attribute (or after a non-kernel stream operation) pdl_wait() / pdl_trigger() are no-ops.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    count_launch();
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define ASR_LAUNCH_TIMED_PDL(slot, pdl, kernel, grid, block, smem, st, ...)          \
    do {                                                                             \
        ::asr::profile_mark(slot, st, true);                                         \
        ::asr::launch_pdl(kernel, grid, block, smem, st, pdl, __VA_ARGS__);          \
        ::asr::profile_mark(slot, st, false);                                        \
    } while (0)

// ---- exact fp32 building blocks ---------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fadd_rn(a, -b); }

// floor(v) for |v| < 2^22 with two full-rate FADDs and no conversion-pipe instruction:
// v + 1.5*2^23 rounded toward -inf lands on the integer grid; its low mantissa bits are the integer.
constexpr float kMagic = 12582912.0f;       // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;      // __float_as_int(kMagic)

struct Floor {
    float f;   // floorf(v) as float
    int raw;   // kMagicBits + (int)floorf(v)
};
__device__ __forceinline__ Floor floor_magic(float v) {
    float t = __fadd_rd(v, kMagic);
    Floor r;
    r.raw = __float_as_int(t);
    r.f = __fadd_rn(t, -kMagic);
    return r;
}

// ImageProjectiveTransformV3 source coordinate: (t0*x + t1*y) + t2 with separately rounded products.
__device__ __forceinline__ float affine_coord(float t0, float x, float t1, float y, float t2) {
    return fadd(fadd(fmul(t0, x), fmul(t1, y)), t2);
}

// bilinear_interpolation() of the op: (x_ceil-x)*a + (x-x_floor)*b per row, then the same in y.
__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11,
                                        float wx0, float wx1, float wy0, float wy1) {
    float top = fadd(fmul(wx0, v00), fmul(wx1, v01));
    float bot = fadd(fmul(wx0, v10), fmul(wx1, v11));
    return fadd(fmul(wy0, top), fmul(wy1, bot));
}

// ---- packed fp32 (sm_100a FADD2 / FMUL2) ---------------------------------------------------------
// Two independent IEEE fp32 lanes in one 64-bit register pair.  Every packed op rounds each lane exactly
// like its scalar counterpart (no fusion), so packing two pixels into one instruction changes nothing
// in the result -- it halves the issue slots the arithmetic needs (measured: scripts/dev/mb_packed.cu,
// FADD2 = 2 clk/warp on the FMA pipes while LDS / integer instructions issue in its shadow).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float pk_lo(f32x2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float pk_hi(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b) { f32x2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// fma.rn.f32x2 is used for two things only, neither of which fuses a product of the op sequence into a sum:
// (a) an exact packed ADD written as m*1.0 + n (see sum2), (b) shared-memory address arithmetic on small integers.
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// (1.0f, 1.0f) in constant memory: a value ptxas cannot see, so the fma below is not simplified back into an add
static __constant__ f32x2 c_one2 = 0x3f8000003f800000ull;
// a + b on both lanes where a and/or b are packed PRODUCTS.  ptxas (12.9) contracts a packed multiply feeding a packed
// add into one FFMA2 even when both carry an explicit .rn (it never does so for scalar ops), which would round the
// product and the sum once instead of twice.  An FFMA2 cannot absorb a second multiply, so the sum is issued as
// fma(a, 1.0, b): a*1.0 is exact and the single rounding of a*1.0 + b is the rounding of the IEEE add a + b
// (signed zeros included), i.e. the result is bit-identical to add.rn.  One issue slot per pixel pair instead of two
// scalar FADDs, same FMA-pipe time.
__device__ __forceinline__ f32x2 sum2(f32x2 a, f32x2 b) { return fma2(a, c_one2, b); }
// bilerp() on two pixels at once: 6 FMUL2 + 3 exact packed sums
__device__ __forceinline__ f32x2 bilerp2(f32x2 v00, f32x2 v01, f32x2 v10, f32x2 v11, f32x2 wx0, f32x2 wx1, f32x2 wy0, f32x2 wy1) {
    const f32x2 top = sum2(mul2(wx0, v00), mul2(wx1, v01));
    const f32x2 bot = sum2(mul2(wx0, v10), mul2(wx1, v11));
    return sum2(mul2(wy0, top), mul2(wy1, bot));
}

// shared-memory address of tap (y0,x0): 4*(raw_y*STRIDE + raw_x) + cst4, where cst4 folds the magic bias of the
// raw floor bits, the box origin and the 32-bit shared address of the tile (mod 2^32).  One shift-add + one
// multiply-add per pixel; the four taps are then immediate offsets of one register.
template <int STRIDE>
__device__ __forceinline__ unsigned tap_offset(unsigned raw_x, unsigned raw_y, unsigned cst4) {
    unsigned o;
    asm("{\n.reg .u32 t;\n"
        "shl.b32 t, %1, 2;\n add.u32 t, t, %3;\n"
        "mad.lo.u32 %0, %2, %4, t;\n}"
        : "=r"(o) : "r"(raw_x), "r"(raw_y), "r"(cst4), "n"(4 * STRIDE));
    return o;
}
// The same address for two pixels at once on the FMA pipe (one issue slot per pixel instead of two integer ones):
// 4*STRIDE*fy + 4*fx + cst evaluated in units of 2^-149, i.e. on fp32 DENORMALS, whose bit pattern is the integer
// itself.  fx, fy are the float floors the gather has anyway; every product and partial sum is an integer below
// 2^23 in those units, so both fmas are exact (this is address arithmetic, not part of the op sequence; without
// -ftz the FMA pipe handles denormals at full rate).  Needs |coordinate| * 4 * STRIDE < 2^23: check_shapes limits
// the canvas to 16384 pixels per side.
__device__ __forceinline__ float denorm_int(int c) { return (c >= 0) ? __uint_as_float((unsigned)c) : -__uint_as_float((unsigned)(-c)); }
template <int STRIDE>
__device__ __forceinline__ f32x2 tap_addr2(f32x2 fxf, f32x2 fyf, f32x2 cst_d) {
    const float k4 = __uint_as_float(4u), ks = __uint_as_float(4u * STRIDE);
    return fma2(fyf, pk(ks, ks), fma2(fxf, pk(k4, k4), cst_d));
}
// ld.shared of one tap: 32-bit shared address + immediate byte offset
template <int OFF>
__device__ __forceinline__ float lds_tap(unsigned addr) {
    float v;
    // volatile: keeps the load between the barrier operations (also volatile asm) that hand the tile over
    asm volatile("ld.shared.f32 %0, [%1 + %2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.0f) ? 1.0f : ((v < 0.0f) ? -1.0f : 0.0f); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mbarrier + TMA (bulk async copy) primitives ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// programmatic dependent launch, device side (see launch_pdl)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one 3-D tiled tensor load global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar,
                                            unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

// contiguous bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- per-copy transforms (host-built, see asr_host.cu) ----------------------------------
// Forward: rotate output->input coefficients t0..t5 and the translate offsets (-dx,-dy).
struct __align__(16) FwdXf { float r0, r1, r2, r3, r4, r5, tx, ty; };
// TensorFlow's gradient of the warp op: the same op with the numerically inverted transform.
struct __align__(16) InvXf { float b0, b1, b2, b3, b4, b5, ux, uy; };

// host side, mirrors tfa.image.angles_to_projective_transforms / translations_to_projective_transforms
void rotate_matrix(float angle, int H, int W, float t[8]);
void invert_transform(const float t[8], float tinv[8]);

}  // namespace asr
