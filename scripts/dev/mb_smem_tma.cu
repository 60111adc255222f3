// Do asynchronous bulk copies into shared memory (the TMA write path) take shared-memory bandwidth away from LDS?
// One CTA per SM, 512 consumer threads: every thread runs a loop of conflict-free LDS.32 (one wavefront per warp instruction).  In the second
// run a producer thread keeps 16 KB cp.async.bulk copies (global, L2-resident -> another shared region) in flight the whole time.
// Prints clocks per LDS warp-instruction with and without the copies, and the bytes the copies moved per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_smem_tma mb_smem_tma.cu && ./mb_smem_tma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <bool COPIES>
__global__ void __launch_bounds__(544) k(const float* __restrict__ src, float* out, long long* clocks, long long* copied, int iters) {
    extern __shared__ __align__(128) float sm[];          // [0, 8192) floats: LDS region; [8192, 8192 + 2*4096): copy targets
    __shared__ __align__(8) unsigned long long bar[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < 8192; i += blockDim.x) sm[i] = 0.f;
    if (tid == 0) {
        reinterpret_cast<volatile int*>(&sm[8192 + 2 * 4096])[0] = 0;   // stop flag
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= 512) {                                      // warp 16: producer
        long long n = 0;
        if (COPIES && tid == 512) {
            const float* s = src + (size_t)blockIdx.x * 8192;
            unsigned phase[2] = {0, 0};
            volatile int* stop = reinterpret_cast<volatile int*>(&sm[8192 + 2 * 4096]);
            for (int b = 0; b < 2; ++b) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + 8192 + b * 4096)), "l"(s + b * 4096), "r"(16384), "r"(smem_u32(&bar[b])) : "memory");
            }
            int b = 0;
            while (!*stop) {
                unsigned ok = 0;
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(phase[b]) : "memory");
                if (!ok) continue;
                phase[b] ^= 1; ++n;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(sm + 8192 + b * 4096)), "l"(s + b * 4096), "r"(16384), "r"(smem_u32(&bar[b])) : "memory");
                b ^= 1;
            }
            // drain the two copies still in flight before the CTA (and its shared memory) goes away
            for (int d = 0; d < 2; ++d) {
                unsigned ok = 0;
                while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(ok) : "r"(smem_u32(&bar[d])), "r"(phase[d]) : "memory");
            }
            copied[blockIdx.x] = n;
        }
        return;
    }
    // consumers: 16 warps, 8 independent conflict-free LDS.32 in flight per thread; each iteration's address comes from
    // the previous iteration's first load so nothing can be hoisted, and the sum is a tree so the FADD chain stays short
    const long long t0 = clock64();
    float acc = 0.f;
    const unsigned base = smem_u32(sm) + 4 * tid;
    float v0 = 0.f;
    for (int i = 0; i < iters; ++i) {
        const unsigned a = base + __float_as_uint(v0);       // the region holds zeros: the address is the same at run time, unknown at compile time
        float v1, v2, v3, v4, v5, v6, v7;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+2048];" : "=f"(v1) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+4096];" : "=f"(v2) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+6144];" : "=f"(v3) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+8192];" : "=f"(v4) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+10240];" : "=f"(v5) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+12288];" : "=f"(v6) : "r"(a) : "memory");
        asm volatile("ld.shared.f32 %0, [%1+14336];" : "=f"(v7) : "r"(a) : "memory");
        acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    }
    const long long t1 = clock64();
    __syncwarp();
    asm volatile("bar.sync 1, 512;");
    if (tid == 0) { reinterpret_cast<volatile int*>(&sm[8192 + 2 * 4096])[0] = 1; clocks[blockIdx.x] = t1 - t0; }
    if (acc == 12345.678f) out[0] = acc;
}

int main() {
    const int nsm = 148, iters = 20000;
    float *src, *out; long long *clocks, *copied;
    cudaMalloc(&src, sizeof(float) * 8192 * nsm); cudaMemset(src, 0, sizeof(float) * 8192 * nsm);
    cudaMalloc(&out, 4); cudaMalloc(&clocks, 8 * nsm); cudaMalloc(&copied, 8 * nsm); cudaMemset(copied, 0, 8 * nsm);
    const size_t smem = sizeof(float) * (8192 + 2 * 4096 + 32);
    cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long hc[nsm], hn[nsm];
    for (int pass = 0; pass < 2; ++pass) {
        for (int with = 0; with < 2; ++with) {
            cudaMemset(src + 0, 0, 4);
            if (with) k<true><<<nsm, 544, smem>>>(src, out, clocks, copied, iters); else k<false><<<nsm, 544, smem>>>(src, out, clocks, copied, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hc, clocks, 8 * nsm, cudaMemcpyDeviceToHost); cudaMemcpy(hn, copied, 8 * nsm, cudaMemcpyDeviceToHost);
            double c = 0, n = 0; for (int i = 0; i < nsm; ++i) { c += hc[i]; n += hn[i]; }
            c /= nsm; n /= nsm;
            const double lds_instr = 16.0 * 8 * iters;     // warp-level LDS per SM
            if (pass) printf("%s bulk copies: %.0f clk for %.0f LDS warp-instructions per SM = %.3f clk each (16 warps, 1 wavefront each)\n",
                             with ? "with   " : "without", c, lds_instr, c / lds_instr);
            if (pass && with) printf("         the copies moved %.1f bytes per clock into shared memory = %.3f 128-byte wavefront-equivalents per clock\n",
                                     n * 16384 / c, n * 16384 / c / 128);
        }
    }
    return 0;
}
