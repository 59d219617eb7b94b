// pcie_probe.cu - host<->device copy rates of the box (pinned memory), contiguous and pitched, one and both directions.
// Build: nvcc -O2 -o tools/pcie_probe tools/pcie_probe.cu      (measurement aid, not part of the library)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t n = (size_t)2 << 30;
    char *h1, *h2, *d1, *d2;
    double t0 = now();
    CK(cudaHostAlloc(&h1, n, cudaHostAllocDefault)); CK(cudaHostAlloc(&h2, n, cudaHostAllocDefault));
    printf("cudaHostAlloc 2 x 2 GiB: %.1f ms\n", (now() - t0) * 1e3);
    memset(h1, 1, n); memset(h2, 2, n);
    t0 = now();
    CK(cudaMalloc(&d1, n)); CK(cudaMalloc(&d2, n));
    printf("cudaMalloc 2 x 2 GiB: %.1f ms\n", (now() - t0) * 1e3);
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    for (int rep = 0; rep < 2; rep++) {
        t0 = now(); CK(cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, s1)); CK(cudaStreamSynchronize(s1));
        printf("H2D contiguous: %.1f GB/s\n", n / (now() - t0) / 1e9);
        t0 = now(); CK(cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s2));
        printf("D2H contiguous: %.1f GB/s\n", n / (now() - t0) / 1e9);
        t0 = now();
        CK(cudaMemcpyAsync(d1, h1, n, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(h2, d2, n, cudaMemcpyDeviceToHost, s2));
        CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        printf("H2D + D2H together: %.1f GB/s each\n", n / (now() - t0) / 1e9);
    }
    // pitched: rows of `w` bytes out of a host pitch of 64800*4 (one cell chunk of a [T, C] f32 array)
    const size_t pitch = 64800 * 4;
    for (size_t w : {(size_t)2048 * 4, (size_t)4096 * 4, (size_t)8192 * 4, (size_t)16384 * 4}) {
        const size_t rows = n / pitch;
        t0 = now(); CK(cudaMemcpy2DAsync(d1, w, h1, pitch, w, rows, cudaMemcpyHostToDevice, s1)); CK(cudaStreamSynchronize(s1));
        printf("H2D pitched rows of %zu B: %.1f GB/s\n", w, w * rows / (now() - t0) / 1e9);
    }
    for (size_t w : {(size_t)2048 * 2, (size_t)4096 * 2, (size_t)16384 * 2}) {
        const size_t hp = 64800 * 2, rows = n / hp;
        t0 = now(); CK(cudaMemcpy2DAsync(h2, hp, d2, w, w, rows, cudaMemcpyDeviceToHost, s2)); CK(cudaStreamSynchronize(s2));
        printf("D2H pitched rows of %zu B: %.1f GB/s\n", w, w * rows / (now() - t0) / 1e9);
    }
    t0 = now(); CK(cudaFree(d1)); CK(cudaFree(d2));
    printf("cudaFree 2 x 2 GiB: %.1f ms\n", (now() - t0) * 1e3);
    // pageable
    char *p = (char *)malloc(n / 4); memset(p, 3, n / 4);
    CK(cudaMalloc(&d1, n));
    t0 = now(); CK(cudaMemcpy(d1, p, n / 4, cudaMemcpyHostToDevice));
    printf("H2D pageable: %.1f GB/s\n", n / 4 / (now() - t0) / 1e9);
    return 0;
}
