#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int XS = 96, XR = 92;
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int cx, int cy, int cz) {
    __shared__ alignas(128) float xt[XS * XR];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_3d_global_to_shared(&xt, &map, cx, cy, cz, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(xt));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < XS * XR; i += blockDim.x) out[i] = xt[i];
}
int main(int argc, char** argv) {
    int B = 2, H = 512, W = 512;
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
    int cx = -5, cy = 100, cz = 1;
    k<<<1, 128>>>(map, o, cx, cy, cz);
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
