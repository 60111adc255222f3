// standalone check of the K1 TMA box load
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#ifndef XS_
#define XS_ 96
#endif
#ifndef XR_
#define XR_ 92
#endif
constexpr int XS = XS_, XR = XR_;
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int cx, int cy, int cz) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    float* xt = (float*)raw;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(XS * XR * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(xt)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < XS * XR; i += blockDim.x) out[i] = xt[i];
}
int main(int argc, char** argv) {
    int B = 2, H = 128, W = 128;
    if (argc > 1) { H = W = atoi(argv[1]); }
    std::vector<float> h((size_t)B * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, XS * XR * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B}; cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {XS, XR, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d (q=%d)\n", (int)r, (int)q);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, XS * XR * 4);
    int cx = -5, cy = 100, cz = 1;
    k<<<1, 128, XS * XR * 4>>>(map, o, cx, cy, cz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> res(XS * XR);
    cudaMemcpy(res.data(), o, XS * XR * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < XR; ++y) for (int x = 0; x < XS; ++x) {
        int gx = cx + x, gy = cy + y;
        float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)cz * H + gy) * W + gx] : 0.f;
        if (res[y * XS + x] != want) ++bad;
    }
    printf("mismatches %d\n", bad);
    return 0;
}
