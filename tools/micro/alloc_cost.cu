// Micro-benchmark: what device / pinned-host allocations cost on this box (decides the cold-path policy of solver.cu).
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    double t0 = now();
    cudaFree(0);
    printf("context %.1f ms\n", (now() - t0) * 1e3);
    for (size_t mb : {1, 16, 64, 256, 1024, 4096}) {
        void *p;
        t0 = now();
        cudaMalloc(&p, mb << 20);
        double t1 = now();
        cudaFree(p);
        printf("cudaMalloc %5zu MiB: %.3f ms   cudaFree %.3f ms\n", mb, (t1 - t0) * 1e3, (now() - t1) * 1e3);
    }
    for (size_t kb : {4, 256, 1024, 4096, 32768, 131072}) {
        void *p;
        t0 = now();
        cudaMallocHost(&p, kb << 10);
        double t1 = now();
        cudaFreeHost(p);
        printf("cudaMallocHost %7zu KiB: %.3f ms   free %.3f ms\n", kb, (t1 - t0) * 1e3, (now() - t1) * 1e3);
    }
    for (size_t mb : {16, 128}) {
        for (int threads : {1, 4, 8}) {
            size_t bytes = mb << 20;
            t0 = now();
            char *p = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            madvise(p, bytes, MADV_HUGEPAGE);
            std::vector<std::thread> pool;
            size_t per = bytes / threads;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([=] {
                    if (madvise(p + t * per, per, MADV_POPULATE_WRITE) != 0)
                        for (size_t o = 0; o < per; o += 4096) p[t * per + o] = 0;
                });
            for (auto &t : pool) t.join();
            double t1 = now();
            cudaHostRegister(p, bytes, cudaHostRegisterDefault);
            double t2 = now();
            cudaHostUnregister(p);
            munmap(p, bytes);
            printf("mmap+populate %4zu MiB with %d threads: %.3f ms   cudaHostRegister %.3f ms\n", mb, threads, (t1 - t0) * 1e3, (t2 - t1) * 1e3);
        }
    }
    // D2H into pinned vs staged
    size_t bytes = (size_t)128 << 20;
    void *d, *h;
    cudaMalloc(&d, bytes);
    cudaMallocHost(&h, bytes);
    for (int i = 0; i < 2; i++) {
        t0 = now();
        cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
        printf("D2H 128 MiB pinned: %.3f ms (%.1f GB/s)\n", (now() - t0) * 1e3, bytes / (now() - t0) / 1e9);
    }
    char *pg = (char *)malloc(bytes);
    memset(pg, 1, bytes);
    t0 = now();
    memcpy(pg, h, bytes);
    printf("memcpy 128 MiB warm, 1 thread: %.3f ms (%.1f GB/s)\n", (now() - t0) * 1e3, bytes / (now() - t0) / 1e9);
    t0 = now();
    cudaMemcpy(pg, d, bytes, cudaMemcpyDeviceToHost);
    printf("D2H 128 MiB pageable (driver staging): %.3f ms (%.1f GB/s)\n", (now() - t0) * 1e3, bytes / (now() - t0) / 1e9);
    return 0;
}
