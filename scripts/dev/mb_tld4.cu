// Microbenchmark: throughput of texture gather (tld4) from a CUDA array vs the shared-memory 4-tap gather.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_tld4 mb_tld4.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__global__ void k_gather(cudaTextureObject_t tex, float* out, int iters, float c, float s) {
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * (blockDim.x / 32) + (threadIdx.x >> 5);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const float fx = floorf(c * X + s * (Y + it) + 3.0f), fy = floorf(-s * X + c * (Y + it) + 2.0f);
        const float4 g = tex2Dgather<float4>(tex, fx + 1.0f, fy + 1.0f, 0);
        acc += g.x + 2.f * g.y + 3.f * g.z + 4.f * g.w;
    }
    out[(size_t)Y * 512 + X] = acc;
}
__global__ void k_check(cudaTextureObject_t tex, float* out) {   // which texel lands in which component
    const float4 g = tex2Dgather<float4>(tex, 10.0f + 1.0f, 20.0f + 1.0f, 0);
    out[0] = g.x; out[1] = g.y; out[2] = g.z; out[3] = g.w;
    const float4 b = tex2Dgather<float4>(tex, 0.0f, 0.0f, 0);       // footprint (-1,-1)..(0,0): border zeros + texel (0,0)
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
int main() {
    const int W = 512, H = 512;
    std::vector<float> h(W * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[y * W + x] = 1000.f * y + x + 1.f;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaArray_t arr; cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather);
    cudaMemcpy2DToArray(arr, 0, 0, h.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float* out; cudaMalloc(&out, W * H * 4);
    k_check<<<1, 1>>>(tex, out);
    float c8[8]; cudaMemcpy(c8, out, 32, cudaMemcpyDeviceToHost);
    printf("gather at floor=(10,20): x=%.0f y=%.0f z=%.0f w=%.0f   (texel(x,y) = 1000*y + x + 1)\n", c8[0], c8[1], c8[2], c8[3]);
    printf("gather at corner (0,0):  x=%.0f y=%.0f z=%.0f w=%.0f\n", c8[4], c8[5], c8[6], c8[7]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 256;
        cudaEventRecord(e0);
        k_gather<<<dim3(16, 64), 256>>>(tex, out, iters, 0.989f, 0.149f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double n = 512.0 * 512.0 * iters;
        printf("tld4: %.1f M gathers in %.3f ms -> %.2f G gathers/s, %.2f lanes/clk/SM at 1.9 GHz\n", n / 1e6, ms, n / ms / 1e6, n / (ms * 1e-3) / 148 / 1.9e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
