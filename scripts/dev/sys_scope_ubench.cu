// What system-scope memory operations cost on a B200 (one thread, %globaltimer around 64 repetitions each): the numbers behind the
// choice of fences in the fused exchange (finalize.cuh push_publish, xchg.cu).   nvcc -arch=sm_100a -o sys_scope_ubench sys_scope_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k(unsigned *loc, unsigned *peer, unsigned long long *out) {
    const int R = 64;
    unsigned v = 0; unsigned long long t0, t1;
    t0 = gt(); for (int i = 0; i < R; ++i) __threadfence(); t1 = gt(); out[0] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) __threadfence_system(); t1 = gt(); out[1] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { loc[i] = i; __threadfence_system(); } t1 = gt(); out[2] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(loc + i), "r"(i) : "memory"); } t1 = gt(); out[3] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { unsigned x; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(loc + i) : "memory"); v += x; } t1 = gt(); out[4] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { unsigned x; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(x) : "l"(loc + i) : "memory"); v += x; } t1 = gt(); out[5] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { unsigned x; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(x) : "l"(loc + i) : "memory"); v += x; } t1 = gt(); out[6] = (t1 - t0) / R;
    t0 = gt(); for (int i = 0; i < R; ++i) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(loc + i), "r"(i) : "memory"); } t1 = gt(); out[7] = (t1 - t0) / R;
    if (peer) {
        t0 = gt(); for (int i = 0; i < R; ++i) { peer[i] = i; __threadfence_system(); } t1 = gt(); out[8] = (t1 - t0) / R;
        t0 = gt(); for (int i = 0; i < R; ++i) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer + i), "r"(i) : "memory"); } t1 = gt(); out[9] = (t1 - t0) / R;
        t0 = gt(); for (int i = 0; i < R; ++i) { peer[i + 64] = i; } asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer), "r"(1) : "memory"); t1 = gt(); out[10] = (t1 - t0);
    }
    out[15] = v;
}
int main() {
    int n = 0; cudaGetDeviceCount(&n);
    unsigned *loc, *peer = nullptr; unsigned long long *out;
    cudaSetDevice(0); cudaMalloc(&loc, 4096); cudaMemset(loc, 0, 4096); cudaMallocManaged(&out, 128);
    if (n > 1) { cudaSetDevice(1); cudaMalloc(&peer, 4096); cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0); }
    for (int rep = 0; rep < 3; ++rep) { k<<<1, 1>>>(loc, peer, out); cudaDeviceSynchronize(); }
    const char *nm[] = {"__threadfence()", "__threadfence_system(), nothing outstanding", "local store + __threadfence_system()", "st.release.sys (local)",
                        "ld.acquire.sys (local)", "ld.volatile (local)", "ld.relaxed.sys (local)", "st.release.gpu (local)",
                        "peer store + __threadfence_system()", "st.release.sys (peer)", "64 peer stores + one st.release.sys (total)"};
    for (int i = 0; i < (peer ? 11 : 8); ++i) printf("%-50s %6llu ns\n", nm[i], out[i]);
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
}
