// Microbenchmark: issue/pipe throughput of scalar vs packed (f32x2) fp32 ops on sm_100a, alone and mixed with
// integer and shared-memory instructions.  One CTA of 1024 threads per SM; prints warp-instructions per clock per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -O3 -o mb_packed mb_packed.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
typedef unsigned long long u64;
#define FADD(x, c)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c))
#define FMUL(x, c)  asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(c))
#define FADDRM(x, c) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(x) : "f"(c))
#define FADD2(x, c) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c))
#define FMUL2(x, c) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c))
#define FADD2RM(x, c) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(x) : "l"(c))
#define FFMA2(x, c) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x) : "l"(c))
#define FFMA(x, c)  asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x) : "f"(c))
#define IADD(i, c)  asm volatile("add.s32 %0, %0, %1;" : "+r"(i) : "r"(c))
#define LOP(i, c)   asm volatile("xor.b32 %0, %0, %1;" : "+r"(i) : "r"(c))
#define IMAD(i, c)  asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(i) : "r"(c))
#define LDS(v, a)   asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a))
#define LDS64(v, a)   asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a))

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, float cf, int ci) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 1024) sm[i] = (float)i;
    __syncthreads();
    float f[8]; u64 p[8]; int n[8]; float ld[4] = {0, 0, 0, 0}; u64 ld2[2] = {0, 0};
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x + i; n[i] = threadIdx.x * i; p[i] = ((u64)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 1.f); }
    u64 c2 = ((u64)__float_as_uint(cf) << 32) | __float_as_uint(cf);
    unsigned sa = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4 + (threadIdx.x >> 5) * 128;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD(f[i], cf); }
        if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD2(p[i], c2); }
        if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FMUL2(p[i], c2); }
        if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FADD2RM(p[i], c2); }
        if (MODE == 4) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD(f[i], cf); IADD(n[i], ci); } }
        if (MODE == 5) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); IADD(n[i], ci); } }
        if (MODE == 6) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); IADD(n[i], ci); LOP(n[i], ci); } }
        if (MODE == 7) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); FMUL(f[i], cf); } }
        if (MODE == 8) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); LDS(ld[i & 3], sa + 4 * i); } }
        if (MODE == 9) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD(f[i], cf); LDS(ld[i & 3], sa + 4 * i); } }
        if (MODE == 10) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); IMAD(n[i], ci); } }
        if (MODE == 11) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD(f[i], cf); FMUL(f[(i + 4) & 7], cf); } }
        if (MODE == 12) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FFMA2(p[i], c2); }
        if (MODE == 13) { _Pragma("unroll") for (int i = 0; i < 8; ++i) FFMA(f[i], cf); }
        if (MODE == 14) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { LDS(ld[i & 3], sa + 4 * i); } }
        if (MODE == 15) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); IADD(n[i], ci); LDS(ld[i & 3], sa + 4 * i);} }
        if (MODE == 16) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD2(p[i], c2); FADD2(p[(i+4)&7], c2); IADD(n[i], ci); LDS(ld[i & 3], sa + 4 * i);} }
        if (MODE == 17) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { IADD(n[i], ci); } }
        if (MODE == 18) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { IMAD(n[i], ci); } }
        if (MODE == 19) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { LDS64(ld2[i & 1], sa + 8 * i); } }
        if (MODE == 20) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { FADD(f[i], cf); IMAD(n[i], ci); } }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += f[i] + (float)n[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    s += ld[0] + ld[1] + ld[2] + ld[3] + __uint_as_float((unsigned)ld2[0]) + __uint_as_float((unsigned)ld2[1]);
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static const char* names[] = {"FADD", "FADD2", "FMUL2", "FADD2.RM", "FADD+IADD", "FADD2+IADD", "FADD2+IADD+LOP", "FADD2+FMUL", "FADD2+LDS",
                              "FADD+LDS", "FADD2+IMAD", "FADD+FMUL", "FFMA2", "FFMA", "LDS", "FADD2+IADD+LDS", "2FADD2+IADD+LDS", "IADD", "IMAD", "LDS64", "FADD+IMAD"};
static const int per[] = {8, 8, 8, 8, 16, 16, 24, 16, 16, 16, 16, 16, 8, 8, 8, 24, 32, 8, 8, 8, 16};
template <int M> void run(float* out, long long* cyc, int nsm) {
    int iters = 4096;
    k<M><<<nsm, 1024>>>(out, cyc, 16, 1.0001f, 3);
    k<M><<<nsm, 1024>>>(out, cyc, iters, 1.0001f, 3);
    cudaDeviceSynchronize();
    std::vector<long long> h(nsm);
    cudaMemcpy(h.data(), cyc, nsm * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    double c = (double)h[nsm / 2];
    double wi = (double)iters * per[M] * 8;   // warp-instr per SMSP (8 warps per SMSP)
    printf("%-18s  %6.3f warp-instr/clk/SMSP   (%d instr per iter, median %0.f clk)\n", names[M], wi / c, per[M], c);
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int nsm = pr.multiProcessorCount;
    float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, nsm * 8);
    printf("%s, %d SMs\n", pr.name, nsm);
    run<0>(out, cyc, nsm); run<1>(out, cyc, nsm); run<2>(out, cyc, nsm); run<3>(out, cyc, nsm); run<4>(out, cyc, nsm); run<5>(out, cyc, nsm);
    run<6>(out, cyc, nsm); run<7>(out, cyc, nsm); run<8>(out, cyc, nsm); run<9>(out, cyc, nsm); run<10>(out, cyc, nsm); run<11>(out, cyc, nsm);
    run<12>(out, cyc, nsm); run<13>(out, cyc, nsm); run<14>(out, cyc, nsm); run<15>(out, cyc, nsm); run<16>(out, cyc, nsm); run<17>(out, cyc, nsm);
    run<18>(out, cyc, nsm); run<19>(out, cyc, nsm); run<20>(out, cyc, nsm);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
