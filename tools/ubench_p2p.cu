// ubench_p2p.cu -- one-way latency of a tagged-word exchange between two GPUs over NVLink
// (diagnostic, not part of the library): GPU 0 stores word i into GPU 1's memory and polls its own
// memory until GPU 1 has answered with word i; round trip / 2 = "store here -> seen by a poll there".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_p2p.cu -o tools/ubench_p2p
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ST, int LD>
__global__ void pingpong(unsigned long long *remote, unsigned long long *local, int n, int first, unsigned long long *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long t0 = clock64();
    for (int i = 1; i <= n; ++i) {
        const unsigned long long v = ((unsigned long long)i << 32) | (unsigned)i;
        if (!first) {                       // the answering side waits first
            unsigned long long w;
            do {
                if (LD == 0) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 2) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 3) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
            } while (w != v);
        }
        if (ST == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(remote), "l"(v) : "memory");
        if (ST == 1) asm volatile("st.global.u64 [%0], %1;" ::"l"(remote), "l"(v) : "memory");
        if (ST == 2) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(v) : "memory");
        if (ST == 3) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(v) : "memory");
        if (ST == 4) asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(remote), "l"(v) : "memory");
        if (first) {
            unsigned long long w;
            do {
                if (LD == 0) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 1) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 2) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
                if (LD == 3) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(local) : "memory");
            } while (w != v);
        }
    }
    out[0] = clock64() - t0;
}

typedef void (*kfn)(unsigned long long *, unsigned long long *, int, int, unsigned long long *);

int main() {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
    unsigned long long *buf[2], *out[2];
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d], 4096));
        CK(cudaMalloc(&out[d], 64));
        CK(cudaStreamCreate(&st[d]));
    }
    const int n = 2000;
    const char *stn[] = {"st.volatile", "st (weak)", "st.relaxed.sys", "st.release.sys", "st.wt"};
    const char *ldn[] = {"ld.cg", "ld.volatile", "ld.relaxed.sys", "ld.acquire.sys"};
    kfn fns[5][4] = {{pingpong<0, 0>, pingpong<0, 1>, pingpong<0, 2>, pingpong<0, 3>},
                     {pingpong<1, 0>, pingpong<1, 1>, pingpong<1, 2>, pingpong<1, 3>},
                     {pingpong<2, 0>, pingpong<2, 1>, pingpong<2, 2>, pingpong<2, 3>},
                     {pingpong<3, 0>, pingpong<3, 1>, pingpong<3, 2>, pingpong<3, 3>},
                     {pingpong<4, 0>, pingpong<4, 1>, pingpong<4, 2>, pingpong<4, 3>}};
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    for (int s = 0; s < 5; ++s)
        for (int l = 0; l < 4; ++l) {
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(buf[d], 0, 4096)); CK(cudaDeviceSynchronize()); }
            for (int d = 1; d >= 0; --d) {          // the answering side (GPU 1) first
                CK(cudaSetDevice(d));
                fns[s][l]<<<1, 32, 0, st[d]>>>(buf[1 - d], buf[d], n, d == 0, out[d]);
            }
            unsigned long long cyc = 0;
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamSynchronize(st[d])); }
            CK(cudaSetDevice(0));
            CK(cudaMemcpy(&cyc, out[0], 8, cudaMemcpyDeviceToHost));
            printf("%-16s + %-16s : round trip %.3f us, one way %.3f us\n", stn[s], ldn[l],
                   (double)cyc / n / (clk * 1e-3), (double)cyc / n / (clk * 1e-3) / 2);
        }
    return 0;
}
