// Microbenchmark 2: packed fp32 with vector-register operands, and mixes with LDS whose results are consumed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o mb_packed2 mb_packed2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
typedef unsigned long long u64;
#define FADD_RR(x, y)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define FMUL_RR(x, y)  asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(y))
#define FADD2_RR(x, y) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y))
#define FMUL2_RR(x, y) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y))
#define FADD2_RRR(d, x, y) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y))
#define LDS(v, a)   asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a))
#define LEA(o, a, b) asm volatile("{.reg .u32 t; shl.b32 t, %2, 7; add.u32 %0, %1, t;}" : "=r"(o) : "r"(a), "r"(b))

template <int MODE, int NW>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, float cf, int ci) {
    __shared__ float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)(i & 7);
    __syncthreads();
    float f[8], g[8]; u64 p[8], q[8]; unsigned n[8];
    for (int i = 0; i < 8; ++i) {
        f[i] = threadIdx.x + i; g[i] = 1.0f + 1e-7f * (threadIdx.x + i); n[i] = (threadIdx.x * 4 + 128 * i) & 8191;
        p[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f);
        q[i] = ((u64)__float_as_uint(g[i]) << 32) | __float_as_uint(g[i]);
    }
    unsigned sa = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD_RR(f[i], g[i]); }
        if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD2_RR(p[i], q[i]); }
        if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FMUL2_RR(p[i], q[i]); }
        if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD2_RRR(p[i], p[(i + 1) & 7], q[i]); }
        if (MODE == 4) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2_RR(p[i], q[i]); FADD_RR(f[i], g[i]); } }   // 2:1 clk packed:scalar
        if (MODE == 5) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { float v; LDS(v, sa + n[i]); FADD_RR(f[i], v); } }     // LDS consumed by FADD
        if (MODE == 6) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { float v; LDS(v, sa + n[i]); FADD_RR(f[i], v); FADD2_RR(p[i], q[i]); } }
        if (MODE == 7) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { float v; unsigned o; LEA(o, sa + n[i], ci); LDS(v, o); FADD_RR(f[i], v); FADD2_RR(p[i], q[i]); FADD2_RR(q[i], p[(i+3)&7]); } }
        if (MODE == 8) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { float v, w2; LDS(v, sa + n[i]); LDS(w2, sa + n[i] + 4); FADD_RR(f[i], v); FADD_RR(g[i], w2); FADD2_RR(p[i], q[i]); FMUL2_RR(q[i], p[(i+3)&7]); } }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + g[i] + (float)n[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)q[i]);
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static const char* names[] = {"FADD r,r", "FADD2 r,r", "FMUL2 r,r", "FADD2 d,a,b", "FADD2+FADD", "LDS->FADD", "LDS->FADD +FADD2", "LEA,LDS->FADD,2xFADD2", "2LDS,2FADD,FADD2,FMUL2"};
static const int per[] = {8, 8, 8, 8, 16, 16, 24, 40, 48};
static const int fpclk[] = {8, 16, 16, 16, 24, 8, 24, 40, 48};
template <int M, int NW> void run(float* out, long long* cyc, int nsm) {
    int iters = 2048;
    k<M, NW><<<nsm, 32 * NW>>>(out, cyc, 16, 1.0001f, 0);
    k<M, NW><<<nsm, 32 * NW>>>(out, cyc, iters, 1.0001f, 0);
    cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, nsm * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double c = (double)h[nsm / 2];
    double wi = (double)iters * per[M] * (NW / 4.0);
    printf("%-26s warps/SMSP=%d  IPC %5.3f   FP-pipe util %5.3f\n", names[M], NW / 4, wi / c, (double)iters * fpclk[M] * (NW / 4.0) / c);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int nsm = pr.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, nsm * 8);
    run<0, 32>(out, cyc, nsm); run<1, 32>(out, cyc, nsm); run<2, 32>(out, cyc, nsm); run<3, 32>(out, cyc, nsm); run<4, 32>(out, cyc, nsm);
    run<5, 32>(out, cyc, nsm); run<6, 32>(out, cyc, nsm); run<7, 32>(out, cyc, nsm); run<8, 32>(out, cyc, nsm);
    run<1, 16>(out, cyc, nsm); run<4, 16>(out, cyc, nsm); run<6, 16>(out, cyc, nsm); run<7, 16>(out, cyc, nsm); run<8, 16>(out, cyc, nsm);
    run<1, 8>(out, cyc, nsm); run<7, 8>(out, cyc, nsm); run<8, 8>(out, cyc, nsm);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
