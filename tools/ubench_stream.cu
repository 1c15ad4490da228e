// ubench_stream.cu -- how fast can one CTA per SM stream global memory into shared memory on
// B200?  Diagnostic micro-benchmark behind the tile/ring sizing of the fused lasso kernel
// (DESIGN.md "stream rate"); not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_stream ubench_stream.cu
//   ./ubench_stream
//
// mode 0: TMA bulk copies (cp.async.bulk) issued by ONE thread into an S-slot mbarrier ring,
//         16 consumer warps wait / touch / release each tile (the fused kernel's mechanism)
// mode 1: same, but every bulk copy is split in `split` pieces issued by different lanes
// mode 2: plain vectorised loads (ld.global.nc.v4) by 512 threads, unroll 8
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int NW = 16;

// every CTA streams `tiles` tiles of `tile_bytes` from its own region of `region_bytes`
// (wrapping), so region_bytes*grid is the footprint
__global__ void __launch_bounds__(NW * 32 + 32, 1)
stream_tma(const unsigned char *base, size_t region_bytes, int tile_bytes, int S, long long tiles, int split,
           int touch, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + 64;
    unsigned char *ring = smem + 1024;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const unsigned char *src0 = base + (size_t)blockIdx.x * region_bytes;
    const long long tiles_per_region = region_bytes / tile_bytes;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (wid == NW) {
        if (split <= 1) {
            if (lane == 0) {
                int slot = 0; uint32_t ph = 1;
                for (long long k = 0; k < tiles; ++k) {
                    while (!mbar_try_wait(empty + slot, ph)) {}
                    mbar_expect_tx(full + slot, (uint32_t)tile_bytes);
                    tma_bulk_g2s(ring + (size_t)slot * tile_bytes, src0 + (size_t)(k % tiles_per_region) * tile_bytes,
                                 (uint32_t)tile_bytes, full + slot);
                    if (++slot == S) { slot = 0; ph ^= 1; }
                }
            }
        } else {
            int slot = 0; uint32_t ph = 1;
            const int piece = tile_bytes / split;
            for (long long k = 0; k < tiles; ++k) {
                while (!mbar_try_wait(empty + slot, ph)) {}
                if (lane == 0) mbar_expect_tx(full + slot, (uint32_t)tile_bytes);
                __syncwarp();
                if (lane < split)
                    tma_bulk_g2s(ring + (size_t)slot * tile_bytes + (size_t)lane * piece,
                                 src0 + (size_t)(k % tiles_per_region) * tile_bytes + (size_t)lane * piece,
                                 (uint32_t)piece, full + slot);
                if (++slot == S) { slot = 0; ph ^= 1; }
            }
        }
    } else {
        int slot = 0; uint32_t ph = 0;
        float acc = 0.f;
        for (long long k = 0; k < tiles; ++k) {
            if (touch >= 2) {
                if (lane == 0) while (!mbar_try_wait(full + slot, ph)) {}
                __syncwarp();
            } else {
                while (!mbar_try_wait(full + slot, ph)) {}
            }
            if (touch & 1) {
                const float4 *t = reinterpret_cast<const float4 *>(ring + (size_t)slot * tile_bytes);
                for (int i = tid; i < tile_bytes / 16; i += NW * 32) { const float4 v = t[i]; acc += v.x + v.y + v.z + v.w; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);
            if (++slot == S) { slot = 0; ph ^= 1; }
        }
        if (acc == 123.456f) sink[0] = acc;
    }
}

__global__ void __launch_bounds__(512, 1)
stream_ldg(const float4 *base, size_t region_vec, long long vec_total, float *sink) {
    const float4 *src = base + (size_t)blockIdx.x * region_vec;
    float acc = 0.f;
    for (long long i0 = 0; i0 < vec_total; i0 += 512 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const size_t idx = (size_t)((i0 + u * 512 + threadIdx.x) % region_vec);
            v[u] = __ldg(src + idx);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t buf_bytes = (size_t)4 << 30;
    unsigned char *buf;
    float *sink;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(buf, 1, buf_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(stream_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    printf("SMs=%d\n", sms);
    const size_t per_cta_bytes = (size_t)96 << 20;   // streamed per CTA per launch
    struct Foot { const char *name; size_t bytes; } foots[] = {{"HBM(4GB)", buf_bytes}, {"L2(40MB)", (size_t)40 << 20}};
    for (auto &f : foots) {
        const size_t region = (f.bytes / sms) / 65536 * 65536;
        // mode 2
        {
            const long long vec_total = per_cta_bytes / 16;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                stream_ldg<<<sms, 512>>>((const float4 *)buf, region / 16, vec_total, sink);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("%-9s LDG.128 x8 unroll, 512 thr/SM            : %8.1f GB/s\n", f.name, (double)per_cta_bytes * sms / ms / 1e6);
        }
        const int tiles_b[] = {4096, 16384, 32768};
        for (int tb : tiles_b) {
            for (int ring_kb : {64, 192}) {
                const int S = ring_kb * 1024 / tb;
                if (S < 2 || S > 64) continue;
                for (int split : {1, 4}) {
                    for (int touch : {0, 1, 2, 3}) {
                        if (split == 4 && touch != 0) continue;
                        const long long tiles = per_cta_bytes / tb;
                        float ms = 0;
                        for (int rep = 0; rep < 2; ++rep) {
                            CK(cudaEventRecord(e0));
                            stream_tma<<<sms, NW * 32 + 32, 1024 + (size_t)S * tb>>>(buf, region, tb, S, tiles, split, touch, sink);
                            CK(cudaEventRecord(e1));
                            CK(cudaEventSynchronize(e1));
                            CK(cudaGetLastError());
                        }
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        printf("%-9s TMA tile=%5d S=%2d (ring %3d KB) split=%d touch=%d : %8.1f GB/s  (%.2f us/tile)\n", f.name, tb, S,
                               ring_kb, split, touch, (double)per_cta_bytes * sms / ms / 1e6, ms * 1e3 / tiles);
                    }
                }
            }
        }
    }
    return 0;
}
