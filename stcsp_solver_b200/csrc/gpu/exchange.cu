// The meeting point of the ranks of a sharded solve ("counts first, then payload ... or
// device-initiated puts"), entirely on the devices.
//
// Every rank owns an XBlock in its own HBM that all peers have mapped (cudaDeviceEnablePeerAccess inside one process,
// CUDA IPC across processes; NVLink 5 / NVSwitch either way).  One exchange = an all-gather of one header row per rank
// plus a barrier:
//     1. copy my row (and my arena directory) into hdr[epoch & 1][my rank] of EVERY peer's block   (remote stores)
//     2. __threadfence_system, then raise flag[my rank] = epoch in every peer's block               (remote stores)
//     3. wait until flag[q] >= epoch for every q in my own block                                    (local polling)
// After the kernel the host of rank r reads all rows from its own block.  Rows are double-buffered by epoch parity: a
// rank can be at most one exchange ahead of the slowest rank (it cannot pass exchange e + 1 before everybody has
// raised the flag for e + 1, which they do only after they have read the rows of e).
//
// The payload itself -- the routed leaf records -- never goes through this block: each rank groups its records by owner
// in an outbox in its own HBM, publishes where it is (header row), and the owners' ingest kernels read the records
// straight out of the producers' outboxes (kernels.cu, PULL mode of ingest_kernel).
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace stcsp {

namespace {

__global__ void __launch_bounds__(256) exchange_kernel(const XPeers peers, const long long *my_row, const ArenaDir *my_dir,
                                                       int rank, int world, unsigned long long epoch, long long timeout_ns,
                                                       int *status) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int par = (int)(epoch & 1ull);
    // 1. my row and my directory into every block (my own included)
    const long long *dir_words = reinterpret_cast<const long long *>(my_dir);
    const int n_dir = (int)(sizeof(ArenaDir) / sizeof(long long));
    for (int q = 0; q < world; q++) {
        XBlock *b = peers.block[q];
        for (int w = tid; w < kHdrWords; w += nt) b->hdr[par][rank][w] = my_row[w];
        long long *dst = reinterpret_cast<long long *>(&b->dir[rank]);
        for (int w = tid; w < n_dir; w += nt) dst[w] = dir_words[w];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere
    if (tid < world) {
        volatile unsigned long long *f = &peers.block[tid]->flag[rank];
        *f = epoch;
    }
    __threadfence_system();
    // 3. wait for everybody (bounded: a peer that died must not hang this GPU)
    if (tid < world) {
        volatile unsigned long long *f = &peers.block[rank]->flag[tid];
        unsigned long long t0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*f < epoch) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if ((long long)(t - t0) > timeout_ns) {
                atomicExch(status, 1);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace

void preload_exchange_kernels() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, exchange_kernel);
    cudaGetLastError();
}

void launch_exchange(const XPeers &peers, const long long *my_row, const ArenaDir *my_dir, int rank, int world,
                     unsigned long long epoch, double timeout_s, int *status, cudaStream_t stream) {
    exchange_kernel<<<1, 256, 0, stream>>>(peers, my_row, my_dir, rank, world, epoch, (long long)(timeout_s * 1e9), status);
}

}  // namespace stcsp
