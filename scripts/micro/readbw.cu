// Micro-benchmark: how fast can 64 x 85 planes of 6400 floats (the level-0 tensor of the bench workload,
// 139 MB) be streamed with (a) linear float4 loads, (b) the decode kernel's access pattern with LDG.128,
// (c) the same pattern with cp.async.bulk (TMA) into shared memory?  Developer tool, not part of the library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float4 ld_stream_v4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__global__ void __launch_bounds__(256) k_linear(const float4* __restrict__ in, long long n4, float* out) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_stream_v4(reinterpret_cast<const float*>(in + i + u * stride));
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { float4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) out[0] = acc;
}

// decode pattern: warp -> 128 consecutive cells of one image, loops the 85 planes, 8 in flight
template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_planes(const float* __restrict__ in, int B, int C, int HW, float* out) {
    const int chunks = HW / 128;
    const int unit = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int b = unit / chunks, ch = unit % chunks;
    if (b >= B) return;
    const float* p = in + ((long long)b * C) * HW + ch * 128 + (threadIdx.x & 31) * 4;
    float best = -1e30f;
    int c = 0;
    for (; c + 8 <= C; c += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ld_stream_v4(p + (long long)(c + u) * HW);
#pragma unroll
        for (int u = 0; u < 8; ++u) best = fmaxf(best, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
    }
    for (; c < C; ++c) { float4 v = ld_stream_v4(p + (long long)c * HW); best = fmaxf(best, v.x + v.y + v.z + v.w); }
    if (best == 12345.678f) out[0] = best;
}

// TMA bulk pattern: CTA -> CELLS consecutive cells of one image; producer thread streams planes in stages of
// PL planes (PL x CELLS x 4 bytes), consumers read shared memory.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}\n"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)));
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int CELLS, int PL, int STAGES>
__global__ void __launch_bounds__(CELLS / 4 + 32) k_tma(const float* __restrict__ in, int B, int C, int HW, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tiles = reinterpret_cast<float*>(smem);                       // STAGES x PL x CELLS floats
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int nthr_cons = CELLS / 4;
    const int chunks = HW / CELLS;
    const int b = blockIdx.x / chunks, ch = blockIdx.x % chunks;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nthr_cons); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nstage = (C + PL - 1) / PL;
    const float* base = in + ((long long)b * C) * HW + ch * CELLS;
    if (tid >= nthr_cons) {
        if (tid == nthr_cons) {          // producer: one elected thread
            for (int it = 0; it < nstage; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
                const int p0 = it * PL, np = min(PL, C - p0);
                mbar_expect(&full[s], (unsigned)(np * CELLS * 4));
                for (int p = 0; p < np; ++p)
                    bulk_g2s(tiles + ((long long)s * PL + p) * CELLS, base + (long long)(p0 + p) * HW, CELLS * 4, &full[s]);
            }
        }
    } else {
        float best = -1e30f;
        for (int it = 0; it < nstage; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const int p0 = it * PL, np = min(PL, C - p0);
            const float4* t4 = reinterpret_cast<const float4*>(tiles + (long long)s * PL * CELLS) + tid;
#pragma unroll 4
            for (int p = 0; p < np; ++p) {
                const float4 v = t4[p * (CELLS / 4)];
                best = fmaxf(best, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
            }
            mbar_arrive(&empty[s]);
        }
        if (best == 12345.678f) out[0] = best;
    }
}

template <typename F>
static float time_it(F f, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f(i);
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f(i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main() {
    const int B = 64, C = 85, HW = 6400;
    const long long n = (long long)B * C * HW;
    const int NBUF = 3;                              // rotate: > L2
    float* in[NBUF]; float* out;
    for (int i = 0; i < NBUF; ++i) { cudaMalloc(&in[i], n * 4); cudaMemset(in[i], 0, n * 4); }
    cudaMalloc(&out, 4);
    const double gb = n * 4 / 1e9;
    float ms = time_it([&](int i) { k_linear<<<148 * 8, 256>>>((const float4*)in[i % NBUF], n / 4, out); }, 30);
    printf("linear LDG.128           %7.1f us  %7.1f GB/s\n", ms * 1e3, gb / (ms * 1e-3));
    ms = time_it([&](int i) { k_planes<4, 8><<<B * (HW / 128) / 4, 128>>>(in[i % NBUF], B, C, HW, out); }, 30);
    printf("planes LDG 4w x8 CTA/SM  %7.1f us  %7.1f GB/s\n", ms * 1e3, gb / (ms * 1e-3));
    ms = time_it([&](int i) { k_planes<8, 4><<<B * (HW / 128) / 8, 256>>>(in[i % NBUF], B, C, HW, out); }, 30);
    printf("planes LDG 8w x4 CTA/SM  %7.1f us  %7.1f GB/s\n", ms * 1e3, gb / (ms * 1e-3));
#define RUN_TMA(CELLS, PL, ST) { \
        const size_t sm = (size_t)ST * PL * CELLS * 4; \
        cudaFuncSetAttribute(k_tma<CELLS, PL, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        ms = time_it([&](int i) { k_tma<CELLS, PL, ST><<<B * (HW / CELLS), CELLS / 4 + 32, sm>>>(in[i % NBUF], B, C, HW, out); }, 30); \
        cudaError_t e = cudaGetLastError(); \
        printf("TMA cells %4d planes/stage %2d stages %d smem %3zu KB  %7.1f us  %7.1f GB/s %s\n", CELLS, PL, ST, sm / 1024, ms * 1e3, gb / (ms * 1e-3), e ? cudaGetErrorString(e) : ""); }
    RUN_TMA(512, 8, 4)
    RUN_TMA(512, 16, 3)
    RUN_TMA(256, 16, 3)
    RUN_TMA(256, 8, 4)
    RUN_TMA(640, 8, 4)
    RUN_TMA(1280, 4, 4)
    cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
