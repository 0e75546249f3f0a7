// Device-side finishing of the automaton (single-rank path): group the edges by source state, apply the fail
// rule as a greatest fixpoint, and compact -- so that the host receives the final arrays in one copy.
//
//   reference                                              here
//   vertexAddEdge (src/graph.cpp:33-38): per-vertex lists   counting sort of the edge store by source id
//   fail marking  (src/solveralgorithm.cpp:865-868,904-910) trim_step_kernel iterated to a fixpoint: a state is
//       failed iff it has no edge into a non-failed state;   failed without live out-edges; edges into failed
//       no edge into a failed state is ever kept             states die, which may fail their sources in turn
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include "kernels.cuh"

namespace stcsp {

namespace {

__global__ void __launch_bounds__(256) edge_count_kernel(const int32_t *src, long long n, int32_t *deg) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        atomicAdd(&deg[src[e]], 1);
}

// one warp per edge: claim a slot in the source's range, copy (src, dst, label)
__global__ void __launch_bounds__(256) edge_scatter_kernel(const int32_t *src, const int32_t *dst, const int32_t *label,
                                                           long long n, int V, const int32_t *first, int32_t *fill,
                                                           int32_t *osrc, int32_t *odst, int32_t *olabel) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_id; e < n; e += total_warps) {
        const int s = src[e];
        int pos = 0;
        if (lane == 0) pos = first[s] + atomicAdd(&fill[s], 1);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (lane == 0) { osrc[pos] = s; odst[pos] = dst[e]; }
        for (int v = lane; v < V; v += 32) olabel[(long long)pos * V + v] = label[e * V + v];
    }
}

__global__ void __launch_bounds__(256) trim_init_kernel(const int32_t *deg, long long n_states, int32_t *outdeg,
                                                        uint8_t *failed, int32_t *changed) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_states; s += (long long)gridDim.x * blockDim.x) {
        const int d = deg[s];
        outdeg[s] = d;
        failed[s] = d == 0;
        if (d == 0) *changed = 1;
    }
}

// one pass: every live edge into a failed state dies; a source that loses its last live edge fails
__global__ void __launch_bounds__(256) trim_step_kernel(const int32_t *src, const int32_t *dst, long long n, uint8_t *alive,
                                                        int32_t *outdeg, uint8_t *failed, int32_t *changed, int32_t *dead) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        if (!alive[e] || !failed[dst[e]]) continue;
        alive[e] = 0;
        atomicAdd(dead, 1);
        if (atomicSub(&outdeg[src[e]], 1) == 1) {
            failed[src[e]] = 1;
            *changed = 1;
        }
    }
}

__global__ void __launch_bounds__(256) alive_to_int_kernel(const uint8_t *alive, long long n, int32_t *flag) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        flag[e] = alive[e];
}

__global__ void __launch_bounds__(256) edge_compact_kernel(const int32_t *src, const int32_t *dst, const int32_t *label,
                                                           const uint8_t *alive, const int32_t *pos, long long n, int V,
                                                           int32_t *osrc, int32_t *odst, int32_t *olabel) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp_id; e < n; e += total_warps) {
        if (!alive[e]) continue;
        const long long p = pos[e];
        if (lane == 0) { osrc[p] = src[e]; odst[p] = dst[e]; }
        for (int v = lane; v < V; v += 32) olabel[p * V + v] = label[e * V + v];
    }
}

// state rows from state keys: cset = max(0, key[0]) (the root's key starts with -1), sig = key[1..]
__global__ void __launch_bounds__(256) state_rows_kernel(const int32_t *keys, long long n_states, int KW, int32_t *cset,
                                                         int32_t *sig) {
    const long long total = n_states * KW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long s = i / KW;
        const int j = (int)(i % KW);
        const int32_t v = keys[i];
        if (j == 0) cset[s] = v < 0 ? 0 : v;
        else sig[s * (KW - 1) + (j - 1)] = v;
    }
}

// Small automata: everything above in ONE CTA (no launches or host round trips between the steps).
//   deg/first/cursor/outdeg: n_states + 1 ints each; failed: n_states bytes; alive: n_edges bytes
// result[0] = edges that died in the fail-rule fixpoint (the host compacts only then).
__global__ void __launch_bounds__(1024) finish_small_kernel(const int32_t *src, const int32_t *dst, const int32_t *label,
                                                           int n_edges, int V, int n_states, int32_t *deg, int32_t *first,
                                                           int32_t *cursor, int32_t *outdeg, uint8_t *failed, uint8_t *alive,
                                                           int32_t *osrc, int32_t *odst, int32_t *olabel, const int32_t *keys,
                                                           int KW, int32_t *cset, int32_t *sig, int do_trim, int32_t *result) {
    __shared__ int32_t partial[1024];
    __shared__ int s_changed, s_dead;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int s = tid; s <= n_states; s += nt) deg[s] = 0;
    if (tid == 0) { s_changed = 0; s_dead = 0; }
    __syncthreads();
    for (int e = tid; e < n_edges; e += nt) atomicAdd(&deg[src[e]], 1);
    __syncthreads();
    // exclusive scan of deg[0 .. n_states]: contiguous chunk per thread, block scan of the chunk sums
    const int n = n_states + 1, chunk = (n + nt - 1) / nt, lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += deg[i];
    partial[tid] = sum;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {
        const int v = tid >= off ? partial[tid - off] : 0;
        __syncthreads();
        partial[tid] += v;
        __syncthreads();
    }
    int run = partial[tid] - sum;
    for (int i = lo; i < hi; i++) {
        const int d = deg[i];
        first[i] = run;
        cursor[i] = run;
        outdeg[i] = d;
        if (i < n_states) failed[i] = do_trim && d == 0;
        run += d;
    }
    __syncthreads();
    // group by source: one warp per edge
    for (int e = warp; e < n_edges; e += nwarps) {
        const int s = src[e];
        int pos = 0;
        if (lane == 0) pos = atomicAdd(&cursor[s], 1);
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (lane == 0) { osrc[pos] = s; odst[pos] = dst[e]; alive[pos] = 1; }
        for (int v = lane; v < V; v += 32) olabel[(long long)pos * V + v] = label[(long long)e * V + v];
    }
    // state rows
    for (int i = tid; i < n_states * KW; i += nt) {
        const int s = i / KW, j = i % KW, v = keys[i];
        if (j == 0) cset[s] = v < 0 ? 0 : v;
        else sig[s * (KW - 1) + (j - 1)] = v;
    }
    __syncthreads();
    // fail rule to a fixpoint
    if (do_trim) {
        for (;;) {
            for (int e = tid; e < n_edges; e += nt) {
                if (!alive[e] || !failed[odst[e]]) continue;
                alive[e] = 0;
                atomicAdd(&s_dead, 1);
                if (atomicSub(&outdeg[osrc[e]], 1) == 1) {
                    failed[osrc[e]] = 1;
                    s_changed = 1;
                }
            }
            __syncthreads();
            const int again = s_changed;
            __syncthreads();
            if (!again) break;
            if (tid == 0) s_changed = 0;
            __syncthreads();
        }
    }
    if (tid == 0) result[0] = s_dead;
}

// ---- multi-rank merge: global state ids (local * W + rank) -> dense ids in ascending global order --------------
struct RankCounts {
    int W;
    long long n[16];        // states owned by each rank
    long long off[16];      // first row of each rank in the concatenated key array
};

__device__ __forceinline__ int dense_id(const RankCounts &rc, long long g) {
    const long long l = g / rc.W;
    const int r = (int)(g % rc.W);
    long long d = 0;
    for (int q = 0; q < rc.W; q++) d += (rc.n[q] < l ? rc.n[q] : l) + (q < r && rc.n[q] > l ? 1 : 0);
    return (int)d;
}

__global__ void __launch_bounds__(256) remap_ids_kernel(int32_t *ids, long long n, RankCounts rc) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ids[i] = dense_id(rc, ids[i]);
}

// keys_in: rank 0's rows, then rank 1's ... ; keys_out: rows in dense order
__global__ void __launch_bounds__(256) place_keys_kernel(const int32_t *keys_in, int32_t *keys_out, long long total_rows, int KW,
                                                         RankCounts rc) {
    const long long total = total_rows * KW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / KW;
        const int j = (int)(i % KW);
        int r = 0;
        while (r + 1 < rc.W && row >= rc.off[r + 1]) r++;
        const long long l = row - rc.off[r];
        keys_out[(long long)dense_id(rc, l * rc.W + r) * KW + j] = keys_in[i];
    }
}

int grid_for(long long n, int per_block, int sm_count) {
    long long g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > (long long)sm_count * 8) g = (long long)sm_count * 8;
    return (int)g;
}

}  // namespace

size_t scan_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    return bytes;
}

void launch_exclusive_scan(void *temp, size_t temp_bytes, const int32_t *in, int32_t *out, long long n, cudaStream_t stream) {
    cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, stream);
}

void launch_edge_count(const int32_t *src, long long n, int32_t *deg, int sm_count, cudaStream_t stream) {
    if (n > 0) edge_count_kernel<<<grid_for(n, 256, sm_count), 256, 0, stream>>>(src, n, deg);
}

void launch_edge_scatter(const int32_t *src, const int32_t *dst, const int32_t *label, long long n, int V, const int32_t *first,
                         int32_t *fill, int32_t *osrc, int32_t *odst, int32_t *olabel, int sm_count, cudaStream_t stream) {
    if (n > 0)
        edge_scatter_kernel<<<grid_for(n, 8, sm_count), 256, 0, stream>>>(src, dst, label, n, V, first, fill, osrc, odst, olabel);
}

void launch_trim_init(const int32_t *deg, long long n_states, int32_t *outdeg, uint8_t *failed, int32_t *changed, int sm_count,
                      cudaStream_t stream) {
    if (n_states > 0) trim_init_kernel<<<grid_for(n_states, 256, sm_count), 256, 0, stream>>>(deg, n_states, outdeg, failed, changed);
}

void launch_trim_step(const int32_t *src, const int32_t *dst, long long n, uint8_t *alive, int32_t *outdeg, uint8_t *failed,
                      int32_t *changed, int32_t *dead, int sm_count, cudaStream_t stream) {
    if (n > 0) trim_step_kernel<<<grid_for(n, 256, sm_count), 256, 0, stream>>>(src, dst, n, alive, outdeg, failed, changed, dead);
}

void launch_alive_to_int(const uint8_t *alive, long long n, int32_t *flag, int sm_count, cudaStream_t stream) {
    if (n > 0) alive_to_int_kernel<<<grid_for(n, 256, sm_count), 256, 0, stream>>>(alive, n, flag);
}

void launch_edge_compact(const int32_t *src, const int32_t *dst, const int32_t *label, const uint8_t *alive, const int32_t *pos,
                         long long n, int V, int32_t *osrc, int32_t *odst, int32_t *olabel, int sm_count, cudaStream_t stream) {
    if (n > 0)
        edge_compact_kernel<<<grid_for(n, 8, sm_count), 256, 0, stream>>>(src, dst, label, alive, pos, n, V, osrc, odst, olabel);
}

void launch_finish_small(const int32_t *src, const int32_t *dst, const int32_t *label, int n_edges, int V, int n_states,
                         int32_t *deg, int32_t *first, int32_t *cursor, int32_t *outdeg, uint8_t *failed, uint8_t *alive,
                         int32_t *osrc, int32_t *odst, int32_t *olabel, const int32_t *keys, int KW, int32_t *cset, int32_t *sig,
                         int do_trim, int32_t *result, cudaStream_t stream) {
    finish_small_kernel<<<1, 1024, 0, stream>>>(src, dst, label, n_edges, V, n_states, deg, first, cursor, outdeg, failed, alive,
                                                osrc, odst, olabel, keys, KW, cset, sig, do_trim, result);
}

static RankCounts make_counts(int world, const long long *n_states) {
    RankCounts rc{};
    rc.W = world;
    long long off = 0;
    for (int r = 0; r < world; r++) {
        rc.n[r] = n_states[r];
        rc.off[r] = off;
        off += n_states[r];
    }
    return rc;
}

void launch_remap_ids(int32_t *ids, long long n, int world, const long long *n_states, int sm_count, cudaStream_t stream) {
    if (n > 0) remap_ids_kernel<<<grid_for(n, 256, sm_count), 256, 0, stream>>>(ids, n, make_counts(world, n_states));
}

void launch_place_keys(const int32_t *keys_in, int32_t *keys_out, long long total_rows, int KW, int world, const long long *n_states,
                       int sm_count, cudaStream_t stream) {
    if (total_rows > 0)
        place_keys_kernel<<<grid_for(total_rows * KW, 256, sm_count), 256, 0, stream>>>(keys_in, keys_out, total_rows, KW,
                                                                                        make_counts(world, n_states));
}

// ---- post-processing fixpoints on the device ---------------------------------------------------
// liveness        reference graphTraverse        src/graph.cpp:357-418  least fixpoint: valid = a final state can be reached
// adversarial -a  reference adversarialTraverse  src/graph.cpp:304-355  greatest fixpoint: every value of variable #5 has an edge
// adversarial -z  reference adversarialTraverse2 src/graph.cpp:247-302  greatest fixpoint: some value of #6 answers every value of #5
// All three run over the edge list that survived the fail rule, one thread per edge / per state, without a CSR (the value sets of
// a state are 64-bit masks collected with atomicOr: domains have at most 64 values); the host iterates each until *changed stays 0.
// `valid` only grows (liveness) or only shrinks (-a, -z), so stale reads inside a sweep only delay the fixpoint by a sweep.
// final = every until flag the reference looks at is set (it looks at numUntil = DISTINCT right-hand variables, :372-374).
__global__ void __launch_bounds__(256) liveness_init_kernel(const int32_t *keys, long long ns, int KW, int n_sig, int n_flags,
                                                            int root_final, uint8_t *fin, uint8_t *valid) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (long long)gridDim.x * blockDim.x) {
        bool f = true;
        if (s == 0) f = root_final != 0;
        else
            for (int c = 0; f && c < n_flags; c++) f = keys[s * KW + 1 + n_sig + c] == 1;
        fin[s] = f;
        valid[s] = f;
    }
}

__global__ void __launch_bounds__(256) liveness_step_kernel(const int32_t *src, const int32_t *dst, long long ne, uint8_t *valid,
                                                            int32_t *changed) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        const int s = src[e], d = dst[e];
        if (s != d && valid[d] && !valid[s]) {
            valid[s] = 1;
            *changed = 1;
        }
    }
}

// -a sweep, edge half: have[src] |= bit(value of the opponent's variable) for every edge into a valid state
__global__ void __launch_bounds__(256) adv1_edge_kernel(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne,
                                                        int V, int var, int lb, const uint8_t *valid,
                                                        unsigned long long *have) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        const int s = src[e];
        if (!valid[s] || !valid[dst[e]]) continue;
        const int b = label[e * V + var] - lb;
        if (b >= 0 && b < 64) atomicOr(have + s, 1ull << b);
    }
}

// -a sweep, state half: a valid state stays valid iff every value lb..ub was seen (reference checkVertexOutEdge, :304-326)
__global__ void __launch_bounds__(256) adv1_state_kernel(long long ns, unsigned long long full, unsigned long long *have,
                                                         uint8_t *valid, int32_t *changed) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (long long)gridDim.x * blockDim.x) {
        const unsigned long long h = have[s];
        have[s] = 0;                                 // ready for the next sweep
        if (valid[s] && h != full) {
            valid[s] = 0;
            *changed = 1;
        }
    }
}

// -z sweep, edge half: cover[src][avatar value] |= bit(opponent value) for every edge into a valid state
__global__ void __launch_bounds__(256) adv2_edge_kernel(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne,
                                                        int V, int op, int op_lb, int ava, int ava_lb, int A,
                                                        const uint8_t *valid, unsigned long long *cover) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        const int s = src[e];
        if (!valid[s] || !valid[dst[e]]) continue;
        const int a = label[e * V + ava] - ava_lb, b = label[e * V + op] - op_lb;
        if (a >= 0 && a < A && b >= 0 && b < 64) atomicOr(cover + (long long)s * A + a, 1ull << b);
    }
}

// -z sweep, state half: valid iff some avatar value saw every opponent value (reference checkVertexOutEdge2, :247-273).
// `clear` = 0 on the last sweep keeps the masks for adv2_kill_kernel.
__global__ void __launch_bounds__(256) adv2_state_kernel(long long ns, int A, unsigned long long full,
                                                         const unsigned long long *cover, uint8_t *valid, int32_t *changed) {
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += (long long)gridDim.x * blockDim.x) {
        if (!valid[s]) continue;
        bool ok = false;
        for (int a = 0; a < A && !ok; a++) ok = cover[s * A + a] == full;
        if (!ok) {
            valid[s] = 0;
            *changed = 1;
        }
    }
}

// What survives: both ends valid, and (after -z, cover != NULL) the avatar value of the edge answers every opponent value
// (the reference deletes the other edges of a valid vertex, :262-271, and every edge into an invalid vertex, :290-300 / :343-353).
__global__ void __launch_bounds__(256) post_alive_kernel(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne,
                                                         int V, int ava, int ava_lb, int A, unsigned long long full,
                                                         const unsigned long long *cover, const uint8_t *valid, uint8_t *alive) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        const int s = src[e];
        bool ok = valid[s] && valid[dst[e]];
        if (ok && cover) {
            const int a = label[e * V + ava] - ava_lb;
            ok = a >= 0 && a < A && cover[(long long)s * A + a] == full;
        }
        alive[e] = ok;
    }
}

void launch_liveness_init(const int32_t *keys, long long ns, int KW, int n_sig, int n_flags, int root_final, uint8_t *fin,
                          uint8_t *valid, int sm_count, cudaStream_t stream) {
    if (ns > 0) liveness_init_kernel<<<grid_for(ns, 256, sm_count), 256, 0, stream>>>(keys, ns, KW, n_sig, n_flags, root_final, fin, valid);
}

void launch_liveness_step(const int32_t *src, const int32_t *dst, long long ne, uint8_t *valid, int32_t *changed, int sm_count,
                          cudaStream_t stream) {
    if (ne > 0) liveness_step_kernel<<<grid_for(ne, 256, sm_count), 256, 0, stream>>>(src, dst, ne, valid, changed);
}

void launch_adv1_sweep(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int var, int lb,
                       unsigned long long full, long long ns, unsigned long long *have, uint8_t *valid, int32_t *changed,
                       int sm_count, cudaStream_t stream) {
    if (ne > 0) adv1_edge_kernel<<<grid_for(ne, 256, sm_count), 256, 0, stream>>>(src, dst, label, ne, V, var, lb, valid, have);
    if (ns > 0) adv1_state_kernel<<<grid_for(ns, 256, sm_count), 256, 0, stream>>>(ns, full, have, valid, changed);
}

void launch_adv2_sweep(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int op, int op_lb, int ava,
                       int ava_lb, int A, unsigned long long full, long long ns, unsigned long long *cover, uint8_t *valid,
                       int32_t *changed, int sm_count, cudaStream_t stream) {
    if (ne > 0)
        adv2_edge_kernel<<<grid_for(ne, 256, sm_count), 256, 0, stream>>>(src, dst, label, ne, V, op, op_lb, ava, ava_lb, A, valid, cover);
    if (ns > 0) adv2_state_kernel<<<grid_for(ns, 256, sm_count), 256, 0, stream>>>(ns, A, full, cover, valid, changed);
}

void launch_post_alive(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int ava, int ava_lb, int A,
                       unsigned long long full, const unsigned long long *cover, const uint8_t *valid, uint8_t *alive, int sm_count,
                       cudaStream_t stream) {
    if (ne > 0)
        post_alive_kernel<<<grid_for(ne, 256, sm_count), 256, 0, stream>>>(src, dst, label, ne, V, ava, ava_lb, A, full, cover, valid, alive);
}

void preload_automaton_kernels(cudaStream_t stream) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, edge_count_kernel);
    cudaFuncGetAttributes(&fa, edge_scatter_kernel);
    cudaFuncGetAttributes(&fa, trim_init_kernel);
    cudaFuncGetAttributes(&fa, trim_step_kernel);
    cudaFuncGetAttributes(&fa, alive_to_int_kernel);
    cudaFuncGetAttributes(&fa, edge_compact_kernel);
    cudaFuncGetAttributes(&fa, finish_small_kernel);
    cudaFuncGetAttributes(&fa, place_keys_kernel);
    cudaFuncGetAttributes(&fa, remap_ids_kernel);
    cudaFuncGetAttributes(&fa, state_rows_kernel);
    cudaFuncGetAttributes(&fa, liveness_init_kernel);
    cudaFuncGetAttributes(&fa, liveness_step_kernel);
    cudaFuncGetAttributes(&fa, adv1_edge_kernel);
    cudaFuncGetAttributes(&fa, adv1_state_kernel);
    cudaFuncGetAttributes(&fa, adv2_edge_kernel);
    cudaFuncGetAttributes(&fa, adv2_state_kernel);
    cudaFuncGetAttributes(&fa, post_alive_kernel);
    // the library kernels behind cub::DeviceScan: run one tiny scan
    int32_t *buf = nullptr;
    const size_t tmp = scan_temp_bytes(64);
    if (cudaMalloc(&buf, 2 * 64 * sizeof(int32_t) + tmp + 256) == cudaSuccess) {
        cudaMemsetAsync(buf, 0, 2 * 64 * sizeof(int32_t), stream);
        launch_exclusive_scan(buf + 128, tmp, buf, buf + 64, 64, stream);
        cudaStreamSynchronize(stream);
        cudaFree(buf);
    }
    cudaGetLastError();
}

void launch_state_rows(const int32_t *keys, long long n_states, int KW, int32_t *cset, int32_t *sig, int sm_count,
                       cudaStream_t stream) {
    if (n_states > 0)
        state_rows_kernel<<<grid_for(n_states * KW, 256, sm_count), 256, 0, stream>>>(keys, n_states, KW, cset, sig);
}

}  // namespace stcsp
