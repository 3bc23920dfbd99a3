// microbench.cu — measures the peaks the raycast kernels are judged against on this B200:
// texture fetch rates (tex3D point rgba16f, tex3D linear u8/f16/f32), LDG gather rates out of
// L1 / L2, L2 and HBM streaming bandwidth. Prints one JSON object; bench.py reads the committed copy
// (profiles/microbench_r01.json) for its roofline denominators.
//
// Pattern: every warp fetches 32 neighbouring texels (8x4 footprint in x,y) at a z that advances
// per iteration inside a footprint of F^3 texels — the coherent access a warp of rays makes.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/microbench bench/microbench.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int ITERS = 256;
constexpr int UNROLL = 8;

// F is a power of two: masks only, so the address arithmetic stays far below the fetch cost.
__device__ __forceinline__ void coords(int it, int F, int& x, int& y, int& z) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    const int bx = (warp * 7 + it * 3) & (F / 8 - 1), by = (warp * 5 + it) & (F / 4 - 1);
    x = bx * 8 + (lane & 7);
    y = by * 4 + (lane >> 3);
    z = (warp * 11 + it * 13) & (F - 1);
}

template <class T> __global__ void __launch_bounds__(256) tex_point_kernel(cudaTextureObject_t tex, int F, float* sink) {
    float acc = 0.f;
    for (int it = 0; it < ITERS; it += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int x, y, z;
            coords(it + u, F, x, y, z);
            const T v = tex3D<T>(tex, x + 0.5f, y + 0.5f, z + 0.5f);
            if constexpr (sizeof(T) == 16) acc += v.x + v.w;
            else acc += *reinterpret_cast<const float*>(&v);
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) tex_linear_kernel(cudaTextureObject_t tex, int F, float* sink) {
    float acc = 0.f;
    for (int it = 0; it < ITERS; it += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int x, y, z;
            coords(it + u, F, x, y, z);
            acc += tex3D<float>(tex, x + 0.37f, y + 0.61f, z + 0.83f);
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

// tld4 on a 2-D layered texture: 4 texels of one bilinear footprint per instruction (the GATHER layout)
__global__ void __launch_bounds__(256) tld4_kernel(cudaTextureObject_t tex, int F, float* sink) {
    float acc = 0.f;
    for (int it = 0; it < ITERS; it += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int x, y, z;
            coords(it + u, F, x, y, z);
            float4 r;
            asm("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(tex), "r"(z), "f"(x + 1.0f), "f"(y + 1.0f));
            acc += r.x + r.y + r.z + r.w;
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

template <class V> __global__ void __launch_bounds__(256) ldg_gather_kernel(const V* __restrict__ buf, int F, float* sink) {
    float acc = 0.f;
    for (int it = 0; it < ITERS; it += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int x, y, z;
            coords(it + u, F, x, y, z);
            const V v = __ldg(buf + ((size_t)z * F + y) * F + x);
            acc += __uint_as_float(*reinterpret_cast<const unsigned*>(&v));
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) stream_read_kernel(const uint4* __restrict__ buf, size_t n16, int reps, float* sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldg(buf + i);
            acc += v.x ^ v.w;
        }
    if (acc == 0x12345679u) sink[0] = (float)acc;
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) b[i] = a[i];
}

template <class F> float time_ms(F&& launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    return best;
}

cudaTextureObject_t make_tex(cudaArray_t* arr_out, cudaChannelFormatDesc d, int F, size_t elem, bool linear, bool norm) {
    cudaArray_t arr;
    CK(cudaMalloc3DArray(&arr, &d, make_cudaExtent(F, F, F)));
    std::vector<unsigned char> h((size_t)F * F * F * elem);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 2654435761u >> 13) & (elem == 2 || elem == 8 ? 0x3b : 0xff);
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(h.data(), (size_t)F * elem, F, F);
    p.dstArray = arr;
    p.extent = make_cudaExtent(F, F, F);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = norm ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    cudaTextureObject_t t;
    CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    *arr_out = arr;
    return t;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    float* sink;
    CK(cudaMalloc(&sink, 4));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    const double fetches = (double)blocks * threads * ITERS;
    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"clock_khz_max\": %d,\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize, clock_khz);

    // footprints: 16^3 (L1-resident for all texel sizes), 64^3, 256^3 (L2-resident)
    const int Fs[3] = {16, 64, 256};
    for (int fi = 0; fi < 3; ++fi) {
        const int F = Fs[fi];
        cudaArray_t arr;
        {
            cudaTextureObject_t t = make_tex(&arr, cudaCreateChannelDescHalf4(), F, 8, false, false);
            float ms = time_ms([&] { tex_point_kernel<float4><<<blocks, threads>>>(t, F, sink); });
            printf(" \"tex3d_point_rgba16f_F%d_gfetch_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(arr));
        }
        {  // rgba8 unorm texels, point filter: the LAYOUT_QUAD fetch (pre-gathered xy quads, two fetches per sample)
            cudaTextureObject_t t = make_tex(&arr, cudaCreateChannelDesc<uchar4>(), F, 4, false, true);
            float ms = time_ms([&] { tex_point_kernel<float4><<<blocks, threads>>>(t, F, sink); });
            printf(" \"tex3d_point_rgba8_F%d_gfetch_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(arr));
        }
        {
            cudaTextureObject_t t = make_tex(&arr, cudaCreateChannelDesc<unsigned char>(), F, 1, true, true);
            float ms = time_ms([&] { tex_linear_kernel<<<blocks, threads>>>(t, F, sink); });
            printf(" \"tex3d_linear_u8_F%d_gfetch_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(arr));
        }
        {
            cudaTextureObject_t t = make_tex(&arr, cudaCreateChannelDescHalf(), F, 2, true, false);
            float ms = time_ms([&] { tex_linear_kernel<<<blocks, threads>>>(t, F, sink); });
            printf(" \"tex3d_linear_f16_F%d_gfetch_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(arr));
        }
        {
            cudaTextureObject_t t = make_tex(&arr, cudaCreateChannelDesc<float>(), F, 4, true, false);
            float ms = time_ms([&] { tex_linear_kernel<<<blocks, threads>>>(t, F, sink); });
            printf(" \"tex3d_linear_f32_F%d_gfetch_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(arr));
        }
        for (int fmt = 0; fmt < 3; ++fmt) {  // tld4 on layered u8 / f16 / f32
            const size_t elem = fmt == 0 ? 1 : (fmt == 1 ? 2 : 4);
            cudaChannelFormatDesc d = fmt == 0 ? cudaCreateChannelDesc<unsigned char>() : (fmt == 1 ? cudaCreateChannelDescHalf() : cudaCreateChannelDesc<float>());
            cudaArray_t larr;
            CK(cudaMalloc3DArray(&larr, &d, make_cudaExtent(F, F, F), cudaArrayLayered));
            std::vector<unsigned char> h((size_t)F * F * F * elem, 0x3b);
            cudaMemcpy3DParms p{};
            p.srcPtr = make_cudaPitchedPtr(h.data(), (size_t)F * elem, F, F);
            p.dstArray = larr;
            p.extent = make_cudaExtent(F, F, F);
            p.kind = cudaMemcpyHostToDevice;
            CK(cudaMemcpy3D(&p));
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = larr;
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = fmt == 0 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
            cudaTextureObject_t t;
            CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
            float ms = time_ms([&] { tld4_kernel<<<blocks, threads>>>(t, F, sink); });
            printf(" \"tld4_a2d_%s_F%d_ginstr_s\": %.2f,\n", fmt == 0 ? "u8" : (fmt == 1 ? "f16" : "f32"), F, fetches / ms * 1e-6);
            CK(cudaDestroyTextureObject(t));
            CK(cudaFreeArray(larr));
        }
        {
            void* buf;
            CK(cudaMalloc(&buf, (size_t)F * F * F * 16));
            CK(cudaMemset(buf, 1, (size_t)F * F * F * 16));
            float ms = time_ms([&] { ldg_gather_kernel<uint4><<<blocks, threads>>>((const uint4*)buf, F, sink); });
            printf(" \"ldg128_gather_F%d_gload_s\": %.2f, \"ldg128_gather_F%d_GBs\": %.1f,\n", F, fetches / ms * 1e-6, F, fetches * 16 / ms * 1e-6);
            ms = time_ms([&] { ldg_gather_kernel<uint2><<<blocks, threads>>>((const uint2*)buf, F, sink); });
            printf(" \"ldg64_gather_F%d_gload_s\": %.2f, \"ldg64_gather_F%d_GBs\": %.1f,\n", F, fetches / ms * 1e-6, F, fetches * 8 / ms * 1e-6);
            ms = time_ms([&] { ldg_gather_kernel<unsigned char><<<blocks, threads>>>((const unsigned char*)buf, F, sink); });
            printf(" \"ldg8_gather_F%d_gload_s\": %.2f,\n", F, fetches / ms * 1e-6);
            CK(cudaFree(buf));
        }
    }
    // streaming: 32 MiB (L2-resident, repeated) and 4 GiB (HBM)
    {
        const size_t n16 = ((size_t)32 << 20) / 16;
        uint4* buf;
        CK(cudaMalloc(&buf, n16 * 16));
        CK(cudaMemset(buf, 1, n16 * 16));
        const int reps = 32;
        float ms = time_ms([&] { stream_read_kernel<<<blocks, threads>>>(buf, n16, reps, sink); });
        printf(" \"l2_stream_read_32MiB_GBs\": %.1f,\n", (double)n16 * 16 * reps / ms * 1e-6);
        CK(cudaFree(buf));
    }
    {
        const size_t n16 = ((size_t)2 << 30) / 16;
        uint4 *a, *b;
        CK(cudaMalloc(&a, n16 * 16));
        CK(cudaMalloc(&b, n16 * 16));
        CK(cudaMemset(a, 1, n16 * 16));
        float ms = time_ms([&] { copy_kernel<<<blocks * 4, threads>>>(a, b, n16); });
        printf(" \"hbm_copy_2GiB_GBs\": %.1f,\n", (double)n16 * 32 / ms * 1e-6);
        ms = time_ms([&] { stream_read_kernel<<<blocks * 4, threads>>>(a, n16, 1, sink); });
        printf(" \"hbm_stream_read_2GiB_GBs\": %.1f,\n", (double)n16 * 16 / ms * 1e-6);
        CK(cudaFree(a));
        CK(cudaFree(b));
    }
    printf(" \"iters_per_thread\": %d, \"threads\": %d\n}\n", ITERS, blocks * threads);
    return 0;
}
