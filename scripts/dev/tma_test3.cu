#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
// (a) bulk copy without descriptor
__global__ void ka(const float* src, float* out) {
    __shared__ alignas(128) float buf[1024];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cuda::memcpy_async(buf, src, cuda::aligned_size_t<16>(sizeof(buf)), bar);
        token = bar.arrive();
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}
// (b) 2D tensor map
__global__ void kb(const __grid_constant__ CUtensorMap map, float* out, int cx, int cy) {
    __shared__ alignas(128) float xt[32 * 32];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&xt, &map, cx, cy, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(xt));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) out[i] = xt[i];
}
int main(int argc, char** argv) {
    int which = atoi(argv[1]);
    int H = 512, W = 512;
    std::vector<float> h((size_t)H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 1024 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    if (which == 0) {
        ka<<<1, 128>>>(d, o);
        printf("bulk kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    } else {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t gs[1] = {(cuuint64_t)W * 4};
        cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d\n", (int)r);
        const unsigned char* mb = (const unsigned char*)&map;
        for (int i = 0; i < 128; ++i) printf("%02x%s", mb[i], (i % 32 == 31) ? "\n" : "");
        kb<<<1, 128>>>(map, o, atoi(argv[2]), atoi(argv[3]));
        printf("tensor kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    }
    std::vector<float> res(1024);
    cudaMemcpy(res.data(), o, 4096, cudaMemcpyDeviceToHost);
    printf("res[0..3] %g %g %g %g\n", res[0], res[1], res[2], res[3]);
    return 0;
}
