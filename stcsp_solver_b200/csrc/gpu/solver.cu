// Host driver of the GPU search: owns the device pools, runs the frontier waves, and implements the
// C ABI of include/stcsp_b200.h (stcsp_gpu_solve, the step-wise session, assemble, trim).
//
// Replaces the recursion of solverSolve / solverSolveRe (reference src/solveralgorithm.cpp:733-1005):
// instead of a depth-first walk with an undo trail, every search node is a self-contained domain
// block in HBM and one WAVE propagates and branches all of them at once (expand_kernel), routes the
// complete assignments found (route_kernel) and merges them into the automaton (ingest_kernel).
// There is no CPU solver in this file: without a CUDA device every entry point fails with
// STCSP_ERR_CUDA.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../host/error.h"
#include "../host/trace_ranges.h"
#include "kernels.cuh"
#include "stcsp_b200.h"
#include "stcsp_host.h"

namespace stcsp {
namespace {

struct Failure : std::runtime_error {
    int code;
    Failure(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct PeerFailure : Failure {      // learnt from a peer's header row: no need to tell the others
    PeerFailure(int c, const std::string &m) : Failure(c, m) {}
};
struct ModelMismatch : Failure {    // the ranks of a group started from different resident copies of the model (seen by all of them)
    ModelMismatch() : Failure(STCSP_ERR_INVALID, "the ranks of the group hold different resident copies of this model") {}
};

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            throw Failure(STCSP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
    } while (0)

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// STCSP_TRACE_GROUP=1: a timestamped line on stderr at every step of a sharded solve (who waits where)
bool group_trace_on() {
    static const bool on = getenv("STCSP_TRACE_GROUP") != nullptr;
    return on;
}
#define GTRACE(...)                                                     \
    do {                                                                \
        if (group_trace_on()) {                                         \
            fprintf(stderr, "[group %.6f] ", now_s());                  \
            fprintf(stderr, __VA_ARGS__);                               \
            fprintf(stderr, "\n");                                      \
        }                                                               \
    } while (0)

// Process-wide cache of device blocks (power-of-two size classes, per device), carved out of a few large ARENAS.
// cudaMalloc / cudaFree cost 0.3-0.6 ms each on B200 and the stream-ordered pool showed 5-350 ms stalls when it had to
// grow, so (1) blocks released by one solve are kept and handed to the next -- in steady state a solve allocates nothing --
// and (2) a block that is not cached is a bump allocation from an arena, so that the FIRST solve of a process pays for
// one cudaMalloc instead of thirty (round 1: 10 ms of first-solve wall time on the headline instance were cudaMalloc).
struct DeviceCache {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void *>> free_blocks;
    struct Arena {
        int device;
        char *base;
        size_t size, used;
        bool has_handle = false;            // exported to peers (CUDA IPC): never freed again -- freeing memory that another
                                            // process still has mapped is undefined
        cudaIpcMemHandle_t handle{};
        long long serial;
        Arena(int d, char *b, size_t s, size_t u) : device(d), base(b), size(s), used(u) {
            static std::atomic<long long> next{1};
            serial = next.fetch_add(1);
        }
    };
    std::vector<Arena> arenas;
    std::map<int, long long> outstanding;           // blocks handed out per device
    // Which arena of `device` holds p (by its serial number), and where in it.
    bool locate(int device, const void *p, long long &serial, long long &offset) {
        std::lock_guard<std::mutex> g(mu);
        for (const Arena &a : arenas) {
            if (a.device != device) continue;
            if ((const char *)p >= a.base && (const char *)p < a.base + a.size) {
                serial = a.serial;
                offset = (const char *)p - a.base;
                return true;
            }
        }
        return false;
    }
    // The arenas of `device` with their CUDA IPC handles (for peers in other processes).
    void directory(int device, ArenaDir &d) {
        std::lock_guard<std::mutex> g(mu);
        d.n = 0;
        for (Arena &a : arenas) {
            if (a.device != device) continue;
            if (d.n >= kMaxArenas) throw Failure(STCSP_ERR_CAPACITY, "more device arenas than a rank can publish to its peers");
            if (!a.has_handle) {
                cudaError_t e = cudaIpcGetMemHandle(&a.handle, a.base);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    memset(&a.handle, 0, sizeof a.handle);      // peers in this process never need it
                }
                a.has_handle = true;
            }
            d.serial[d.n] = a.serial;
            d.size[d.n] = (long long)a.size;
            static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
            memcpy(d.handle[d.n], &a.handle, 64);
            d.n++;
        }
    }
    static constexpr size_t kFirstArena = (size_t)256 << 20;
    static size_t size_class(size_t bytes) {
        size_t c = 4096;
        while (c < bytes) c <<= 1;
        return c;
    }
    void *acquire(int device, size_t bytes) {       // bytes must be a size class
        std::lock_guard<std::mutex> g(mu);
        outstanding[device]++;
        auto it = free_blocks.find({device, bytes});
        if (it != free_blocks.end() && !it->second.empty()) {
            void *p = it->second.back();
            it->second.pop_back();
            return p;
        }
        size_t biggest = 0;
        for (Arena &a : arenas) {
            if (a.device != device) continue;
            biggest = std::max(biggest, a.size);
            if (a.size - a.used >= bytes) {
                void *p = a.base + a.used;
                a.used += bytes;                    // size classes are multiples of 4096: every block stays 4 KiB-aligned
                return p;
            }
        }
        // a new arena: at least twice the largest so far (the pools of one solve double as they grow)
        size_t want = std::max(bytes, biggest ? biggest * 2 : kFirstArena);
        void *p = nullptr;
        GTRACE("device %d: cudaMalloc of a %zu MiB arena ...", device, want >> 20);
        cudaError_t e = cudaMalloc(&p, want);
        GTRACE("device %d: ... arena allocated", device);
        if (e != cudaSuccess && want > bytes) {     // not that much room left: just what is needed
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            outstanding[device]--;
            throw Failure(STCSP_ERR_CAPACITY, "device allocation of " + std::to_string(bytes) + " bytes failed: " +
                                                  cudaGetErrorString(e));
        }
        arenas.push_back(Arena(device, (char *)p, want, bytes));
        return p;
    }
    void give_back(int device, size_t bytes, void *p) {
        std::lock_guard<std::mutex> g(mu);
        outstanding[device]--;
        free_blocks[{device, bytes}].push_back(p);
    }
    // Give the memory back to the driver.  Blocks are slices of arenas, so this only happens when nothing is in use.
    void trim(int device) {
        std::lock_guard<std::mutex> g(mu);
        if (outstanding[device] > 0) return;
        for (auto &kv : free_blocks) {
            if (kv.first.first != device) continue;
            std::vector<void *> keep;               // blocks inside exported arenas stay usable
            for (void *p : kv.second)
                for (const Arena &a : arenas)
                    if (a.device == device && a.has_handle && (char *)p >= a.base && (char *)p < a.base + a.size) keep.push_back(p);
            kv.second.swap(keep);
        }
        for (size_t i = 0; i < arenas.size();) {
            if (arenas[i].device == device && !arenas[i].has_handle) {
                cudaFree(arenas[i].base);
                arenas.erase(arenas.begin() + (long)i);
            } else {
                i++;
            }
        }
    }
};
DeviceCache &device_cache() {
    static DeviceCache *c = new DeviceCache();      // leaked on purpose (no CUDA calls during static destruction)
    return *c;
}

template <class T>
struct DBuf {
    T *p = nullptr;
    size_t cap = 0;     // elements
    size_t bytes = 0;   // size class of the block
    int device = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    // The caller guarantees that no work using the block is in flight (the session synchronises its stream first).
    void release() {
        if (p) device_cache().give_back(device, bytes, p);
        p = nullptr;
        cap = bytes = 0;
    }
    // at least `need` elements; the first `keep` elements survive a reallocation
    void swap(DBuf &o) {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
        std::swap(bytes, o.bytes);
        std::swap(device, o.device);
    }
    bool reserve(size_t need, size_t keep, cudaStream_t st) {
        if (need <= cap) return false;
        const size_t nbytes = DeviceCache::size_class(std::max(need, 2 * cap) * sizeof(T));
        int dev = 0;
        CK(cudaGetDevice(&dev));
        T *np = (T *)device_cache().acquire(dev, nbytes);
        if (p) {
            if (keep) CK(cudaMemcpyAsync(np, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));          // nothing in flight may still touch the old block
            device_cache().give_back(device, bytes, p);
        }
        p = np;
        cap = nbytes / sizeof(T);
        bytes = nbytes;
        device = dev;
        return true;
    }
};

// Everything a solve derives from the model alone -- compiled constraint sets, relation tables, the constraint-set
// transition map, and their device copies -- stays resident between solves of the same model (take / put: one session
// at a time owns an entry; a concurrent solve of the same model simply builds its own).
struct ModelState {
    SetTable sets;
    // constraint-set transitions resolved so far
    std::vector<CapEntry> capmap;
    std::vector<int32_t> capvals;
    long long capmap_used = 0;
    std::map<std::vector<int32_t>, int32_t> cap_lookup;
    // device copies
    DBuf<int32_t> d_lb, d_width, d_sigvars, d_scope, d_stride, d_aux, d_arr_off, d_arr_val, d_capvals;
    DBuf<unsigned long long> d_tables;
    DBuf<DevSet> d_sets;
    DBuf<DevCon> d_cons;
    DBuf<DevProp> d_props;
    DBuf<Instr> d_code;
    DBuf<uint32_t> d_wake;
    DBuf<CapEntry> d_capmap;
    long long tables_built = 0;     // u64 words of the table pool already filled
    bool uploaded = false;
    // pool sizes the last solve of this model ended with: the next one starts there and never has to grow
    long long hint_frontier = 0, hint_states = 0, hint_edges = 0, hint_table = 0;
    long long hint_max_wave = 0;        // widest wave of the last solve (0: unknown)
    long long hint_out_states = 0, hint_out_edges = 0;     // the automaton the last solve returned (0: unknown)
};

struct ModelCache {
    static constexpr size_t kCapacity = 24;
    std::mutex mu;
    std::map<std::string, std::unique_ptr<ModelState>> entries;
    std::vector<std::string> order;
    std::unique_ptr<ModelState> take(const std::string &key) {
        std::lock_guard<std::mutex> g(mu);
        auto it = entries.find(key);
        if (it == entries.end()) return nullptr;
        std::unique_ptr<ModelState> e = std::move(it->second);
        entries.erase(it);
        order.erase(std::remove(order.begin(), order.end(), key), order.end());
        return e;
    }
    void put(const std::string &key, std::unique_ptr<ModelState> e) {
        std::lock_guard<std::mutex> g(mu);
        if (entries.count(key)) return;
        entries[key] = std::move(e);
        order.push_back(key);
        while (order.size() > kCapacity) {  // oldest out (its blocks go back to the device cache) ...
            // ... together with the copies the other ranks of its group keep in this process (stcsp_gpu_solve_multi): the
            // ranks of a group must all have a resident copy of a model, or none
            const std::string victim = order.front();
            const size_t cut = victim.find("|group");
            const std::string base = cut == std::string::npos ? victim : victim.substr(0, cut);
            for (size_t i = 0; i < order.size();) {
                const bool same = order[i] == victim ||
                                  (cut != std::string::npos && order[i].compare(0, base.size(), base) == 0 &&
                                   order[i].find("|group", base.size()) == base.size());
                if (same) {
                    entries.erase(order[i]);
                    order.erase(order.begin() + (long)i);
                } else {
                    i++;
                }
            }
        }
    }
};
ModelCache &model_cache() {
    static ModelCache *c = new ModelCache();
    return *c;
}

std::string model_key(const stcsp_problem_t &p, int device) {
    std::string k;
    auto add = [&](const void *ptr, size_t n) { k.append((const char *)ptr, n); };
    add(&device, sizeof device);
    add(&p.prefix_k, sizeof p.prefix_k);
    add(&p.n_vars, sizeof p.n_vars);
    add(p.var_lb, sizeof(int32_t) * p.n_vars);
    add(p.var_ub, sizeof(int32_t) * p.n_vars);
    add(&p.n_arrays, sizeof p.n_arrays);
    if (p.n_arrays > 0) {
        add(p.arr_offsets, sizeof(int32_t) * (p.n_arrays + 1));
        add(p.arr_values, sizeof(int32_t) * p.arr_offsets[p.n_arrays]);
    }
    add(&p.n_constraints, sizeof p.n_constraints);
    add(p.con_offsets, sizeof(int32_t) * (p.n_constraints + 1));
    add(p.con_tokens, sizeof(stcsp_tok_t) * p.con_offsets[p.n_constraints]);
    return k;
}

// A stream with its timing events, pooled per device: creating them costs ~50 us per solve otherwise.
struct ExecContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int sm_count = 0, coop = 0;
};
struct ExecPool {
    std::mutex mu;
    std::vector<ExecContext> idle;
    ExecContext acquire(int device) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < idle.size(); i++)
                if (idle[i].device == device) {
                    ExecContext c = idle[i];
                    idle.erase(idle.begin() + (long)i);
                    return c;
                }
        }
        ExecContext c;
        c.device = device;
        CK(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, device));
        if (cudaDeviceGetAttribute(&c.coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) c.coop = 0;
        CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) CK(cudaEventCreate(&c.ev[i]));
        return c;
    }
    void give_back(const ExecContext &c) {
        std::lock_guard<std::mutex> g(mu);
        idle.push_back(c);
    }
};
ExecPool &exec_pool() {
    static ExecPool *p = new ExecPool();
    return *p;
}

// Host memory the device writes into: always PINNED, but never through cudaMallocHost.
//
// Measured on this pool's B200 hosts (tools/micro/alloc_cost.cu): cudaMallocHost costs 2-3 ms even for 4 KiB and
// 0.45 ms per MiB beyond (128 MiB: 60 ms); an anonymous mapping faulted in by four threads costs 0.035 ms per MiB and
// cudaHostRegister of the populated range 0.03 ms per MiB (128 MiB: 4.5 + 4.2 ms).  Round 1 pinned the 89 MiB automaton of
// partialorder_14 with cudaMallocHost: 80 ms of a 96 ms first solve.  So a block is mmap + parallel populate +
// cudaHostRegister; small blocks (<= 256 KiB: counters, the automaton of a small instance) are slices of an arena made
// the same way, so that the first solve of a process registers host memory once.  Released blocks are cached by size
// class like the device blocks.  If registration fails the block stays pageable and is filled through two pinned
// staging chunks (download()).
struct HostCache {
    enum Kind : int { PINNED = 0, PAGEABLE = 1 };
    std::mutex mu;
    std::map<size_t, std::vector<void *>> free_blocks;        // registered, by size class
    struct Arena {
        char *base;
        size_t size, used;
        char *dev_base = nullptr;       // where the device sees `base` (mapped registration), asked for once per arena
        bool dev_asked = false;
    };
    std::vector<Arena> arenas;
    // Device-side address of a pinned block.  Blocks carved out of an arena share its mapping, so the driver is asked once
    // per arena, not seven times per solve (cudaHostGetDevicePointer is a driver call of about a microsecond).
    void *device_ptr(void *h) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (Arena &a : arenas)
                if ((char *)h >= a.base && (char *)h < a.base + a.size) {
                    if (!a.dev_asked) {
                        void *d = nullptr;
                        a.dev_asked = true;
                        if (cudaHostGetDevicePointer(&d, a.base, 0) == cudaSuccess) a.dev_base = (char *)d;
                        else cudaGetLastError();
                    }
                    return a.dev_base ? a.dev_base + ((char *)h - a.base) : nullptr;
                }
        }
        void *d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return d;
    }
    static constexpr size_t kSmall = (size_t)256 << 10;
    static constexpr size_t kFirstArena = (size_t)1 << 20;
    void *acquire(size_t bytes, int &kind) {
        kind = PINNED;
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_blocks.find(bytes);
            if (it != free_blocks.end() && !it->second.empty()) {
                void *p = it->second.back();
                it->second.pop_back();
                return p;
            }
            if (bytes > kSmall && big_state.load(std::memory_order_acquire) == 2) {
                // the arena prepared in the background (prepare_big): large blocks are carved out of it, page-aligned
                const size_t need = (bytes + 4095) & ~(size_t)4095;
                if (big.size - big.used >= need) {
                    void *p = big.base + big.used;
                    big.used += need;
                    return p;
                }
            }
            if (bytes <= kSmall) {
                for (Arena &a : arenas)
                    if (a.size - a.used >= bytes) {
                        void *p = a.base + a.used;
                        a.used += bytes;
                        return p;
                    }
                const size_t want = arenas.empty() ? kFirstArena : std::min<size_t>(arenas.back().size * 2, (size_t)64 << 20);
                int k = PINNED;
                void *p = map_and_register(want, k);
                if (k == PINNED) {
                    arenas.push_back(Arena{(char *)p, want, bytes});
                    return p;
                }
                munmap(p, want);
            }
        }
        return map_and_register(bytes, kind);
    }
    // ---- a large pinned arena, prepared ahead of time (stcsp_gpu_warmup) ---------------------------------------------
    // Pinning fresh host memory costs ~4-7 GB/s here (the OS zeroes the pages, the driver registers them): 15 of the 22 ms of
    // a FIRST partialorder_14 solve in a warm process went into the 91 MB of its result.  A process that will solve more than
    // once can pin one arena up front; from then on a result block of a size class not seen before is a pointer bump.
    // Blocks given back are kept by size class like any other; what does not fit goes the old way.  (Doing this behind the
    // caller's back in a background thread was tried and dropped: cudaHostRegister of 512 MB holds the driver for the better
    // part of a second, and a solve that runs meanwhile stalls -- partialorder_14 6 ms -> 99 ms.)
    Arena big{nullptr, 0, 0};
    std::atomic<int> big_state{0};             // 0 none, 1 being prepared, 2 ready
    void prepare_big(size_t bytes) {
        int expect = 0;
        if (bytes == 0 || !big_state.compare_exchange_strong(expect, 1)) return;       // (one arena per process; 1 = being prepared)
        bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        int kind = PINNED;
        void *p = nullptr;
        try {
            p = map_and_register(bytes, kind);
        } catch (...) {
            big_state.store(0);
            throw;
        }
        if (kind != PINNED) {
            munmap(p, bytes);
            big_state.store(0);
            throw Failure(STCSP_ERR_CUDA, "cannot pin the host arena");
        }
        std::lock_guard<std::mutex> g(mu);
        big = Arena{(char *)p, bytes, 0};
        arenas.push_back(Arena{(char *)p, bytes, bytes});      // (known to release_all; nothing small is carved out of it)
        big_state.store(2, std::memory_order_release);
    }
    void give_back(void *p, size_t bytes, int kind) {
        if (kind == PAGEABLE) { munmap(p, bytes); return; }
        std::lock_guard<std::mutex> g(mu);
        free_blocks[bytes].push_back(p);
    }
    // Anonymous mapping (huge pages requested), faulted in by up to four threads, then registered with the driver.
    static void *map_and_register(size_t bytes, int &kind) {
        void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
        madvise(p, bytes, MADV_HUGEPAGE);
#endif
        auto populate = [p](size_t lo, size_t hi) {
#ifdef MADV_POPULATE_WRITE
            if (madvise((char *)p + lo, hi - lo, MADV_POPULATE_WRITE) == 0) return;
#endif
            for (size_t o = lo; o < hi; o += 4096) ((volatile char *)p)[o] = 0;
        };
        const size_t slice = (size_t)4 << 20;
        const size_t n_threads = std::min<size_t>(4, bytes / slice);
        if (n_threads >= 2) {
            std::vector<std::thread> pool;
            const size_t per = ((bytes / n_threads) + 4095) & ~(size_t)4095;
            for (size_t t = 1; t < n_threads; t++) {
                const size_t lo = t * per, hi = t + 1 == n_threads ? bytes : std::min(bytes, lo + per);
                if (lo < hi) pool.emplace_back(populate, lo, hi);
            }
            populate(0, std::min(bytes, per));
            for (std::thread &t : pool) t.join();
        } else {
            populate(0, bytes);
        }
        kind = PINNED;
        if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
            cudaGetLastError();
            kind = PAGEABLE;
        }
        return p;
    }
    void release_all() {
        std::lock_guard<std::mutex> g(mu);
        for (auto &kv : free_blocks) {
            for (void *p : kv.second) {
                bool in_arena = false;
                for (const Arena &a : arenas) in_arena |= (char *)p >= a.base && (char *)p < a.base + a.size;
                if (!in_arena) {
                    cudaHostUnregister(p);
                    munmap(p, kv.first);
                }
            }
            kv.second.clear();
        }
        // (arena slices may still be referenced by live automata: the arenas stay)
    }
};
HostCache &host_cache() {
    static HostCache *c = new HostCache();      // leaked on purpose: outlives the CUDA context teardown order
    return *c;
}

// The counter read-back block of a session: a small pinned block like any other.
struct PinnedCache {
    static constexpr size_t kBytes = 4096;
    unsigned long long *acquire() {
        static_assert(C_COUNT * sizeof(unsigned long long) + 256 <= kBytes, "counter block");
        int kind = 0;
        void *p = host_cache().acquire(kBytes, kind);
        if (kind != HostCache::PINNED) {
            host_cache().give_back(p, kBytes, kind);
            throw Failure(STCSP_ERR_CUDA, "cannot pin host memory for the counter read-back");
        }
        return (unsigned long long *)p;
    }
    void give_back(unsigned long long *b) { host_cache().give_back(b, kBytes, HostCache::PINNED); }
};
PinnedCache &pinned_cache() {
    static PinnedCache *c = new PinnedCache();
    return *c;
}

struct Store {          // backing storage of a library-owned stcsp_automaton_t
    virtual ~Store() {}
};

struct AutoStore : Store {      // host vectors (assembled / multi-rank automata)
    std::vector<int32_t> sig_vars, state_sig, state_cset, edge_src, edge_dst, edge_label;
    std::vector<uint8_t> state_failed;
};

void bind_store(stcsp_automaton_t *a, AutoStore *st) {
    a->impl = static_cast<Store *>(st);
    a->sig_vars = st->sig_vars.data();
    a->state_sig = st->state_sig.data();
    a->state_cset = st->state_cset.data();
    a->state_failed = st->state_failed.data();
    a->edge_src = st->edge_src.data();
    a->edge_dst = st->edge_dst.data();
    a->edge_label = st->edge_label.data();
}

struct HostBlock {
    void *p = nullptr;
    size_t bytes = 0;
    int kind = HostCache::PINNED;
    bool pinned() const { return kind == HostCache::PINNED; }
    void alloc(size_t need) {
        bytes = DeviceCache::size_class(std::max<size_t>(need, 1));
        p = host_cache().acquire(bytes, kind);
    }
    ~HostBlock() { if (p) host_cache().give_back(p, bytes, kind); }
};

// dst <- src with up to four threads (one thread copies ~10 GB/s, a PCIe 5 link delivers ~50 GB/s)
void parallel_copy(void *dst, const void *src, size_t n) {
    const size_t piece = (size_t)8 << 20;
    if (n < 2 * piece) { memcpy(dst, src, n); return; }
    const size_t n_threads = std::min<size_t>(4, n / piece);
    const size_t per = (n + n_threads - 1) / n_threads;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < n_threads; t++) {
        const size_t lo = t * per, hi = std::min(n, lo + per);
        if (lo < hi) pool.emplace_back([=] { memcpy((char *)dst + lo, (const char *)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(n, per));
    for (std::thread &t : pool) t.join();
}

struct PinnedStore : Store {    // single-rank automata finished on the device
    HostBlock sig_vars, state_sig, state_cset, state_failed, edge_src, edge_dst, edge_label, state_final, state_valid, edge_alive;
};

}  // namespace
}  // namespace stcsp

using namespace stcsp;

struct stcsp_session {
    std::unique_ptr<ModelState> model;
    stcsp_options_t opt{};
    int rank = 0, world = 1, device = 0, sm_count = 148, coop_launch = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    DevModel dm{};
    DBuf<int32_t> d_jobs;
    std::string cache_key, group_tag;
    DBuf<SearchCtl> d_ctl;
    // finishing scratch (also handed to the search kernel, which finishes small automata itself)
    DBuf<int32_t> fb_deg, fb_first, fb_fill, fb_outdeg, fb_src, fb_dst, fb_label, fb_cset, fb_sig, fb_flags;
    DBuf<uint8_t> fb_failed, fb_alive, fb_scan, fb_fin, fb_valid, fb_palive;
    DBuf<unsigned long long> fb_cover;
    bool finish_in_kernel = false, finish_trim = true;     // set by stcsp_gpu_solve (single rank)
    bool time_expand = false;                               // expand launches of run_persistent's wide waves are timed
    static constexpr long long kWideWaveNodes = 32768;     // default of stcsp_options_t::wide_wave_nodes
    static constexpr long long kNarrowInstanceWave = 2048; // widest wave of an instance that runs on one CTA per SM
    static constexpr long long kFirstFrontier = 32768, kFirstStates = 65536, kFirstEdges = 262144;    // pool sizes of a first solve
    bool wide_grid = false;         // the narrow (one CTA per SM) grid was given up for this solve
    bool search_complete = false;                           // the wave loop ran to the end (max_wave is the instance's)
    bool poisoned = false;          // a pool overflowed in mid-wave (table slots tombstoned, cursors past their capacity):
                                    // unreachable as long as every path pre-reserves n_states + n, fatal for the session if not
    bool deg_ok = false;            // fb_deg has counted every edge of this solve so far (FinishArgs::deg_counted)
    const int32_t *deg_ptr = nullptr, *cur_ptr = nullptr;
    long long deg_cs = 0;
    bool sa_first_launch_of_solve() const { return root_pending; }
    bool prefinished = false;       // the search kernel already grouped + trimmed (fb_* hold the result)
    long long prefinished_dead = 0;
    // PUSH (FinishArgs::h_*): pinned output buffers handed to the search kernel, which writes a small automaton into them
    unsigned long long *host_mirror = nullptr;     // device-side address of h_counters (mapped pinned memory), or null
    PinnedStore *push_store = nullptr;
    long long push_cap_states = 0, push_cap_edges = 0;
    bool pushed = false;            // ... and did (no device-to-host copy is needed)
    bool root_in_kernel = false;    // set by stcsp_gpu_solve before init(): the persistent path creates the root on the device
    bool root_pending = false;      // ... and has not done so yet
    bool counters_fresh = false;    // every counter set is zero (init() has just cleared them)
    long long out_states = 0, out_edges = 0;        // size of the automaton this solve returned
    SearchCtl *h_ctl = nullptr;     // pinned, behind h_counters
    int search_grid = 0;
    // search pools
    DBuf<int32_t> frontier[2];
    int cur = 0;
    DBuf<int32_t> leaves, unresolved, gathered, table, state_key, edge_src, edge_dst, edge_label;
    DBuf<unsigned long long> counters;
    DBuf<long long> d_offsets;
    unsigned long long *h_counters = nullptr;       // pinned
    long long table_size = 0;
    long long n_in = 0, n_out = 0, n_leaves = 0, n_unres = 0, n_states = 0, n_edges = 0;
    long long max_wave = 0;             // widest wave so far
    int expand_grid_max = 148;
    std::vector<int32_t> pending;                   // flat requests [cid, values[V]]
    // statistics
    long long t_nodes = 0, t_fails = 0, t_tuples = 0, t_revisions = 0, t_leaves = 0, t_dominance = 0, t_waves = 0,
              t_launches = 0, t_expand_launches = 0, h2d = 0, d2h = 0;
    double expand_ms = 0, t_create = 0;
    bool timing_open = false;

    ~stcsp_session() {
        if (stream) cudaStreamSynchronize(stream);
        delete push_store;
        if (h_counters) pinned_cache().give_back(h_counters);
        // Only a solve that ran to the end hands its model back: after a failure (an exception between the host and the
        // device update of the transition map, a timeout, a CUDA error) the host and device copies may disagree.
        if (model && search_complete && !poisoned && !cache_key.empty() && model->uploaded && !model->sets.dirty() &&
            model->sets.table_jobs.empty() && model->tables_built == model->sets.table_words && dm.node_words > 0) {
            model->hint_frontier = (long long)std::min(frontier[0].cap, frontier[1].cap) / dm.node_words;
            model->hint_states = (long long)state_key.cap / dm.key_words;
            model->hint_edges = (long long)edge_src.cap;
            model->hint_table = table_size;
            if (search_complete) model->hint_max_wave = max_wave;
            model->hint_out_states = out_states;
            model->hint_out_edges = out_edges;
            model_cache().put(cache_key, std::move(model));       // the compiled model stays resident for the next solve
        }
        release_all();
        if (stream) {                       // idle again (synchronised above): back to the pool
            ExecContext c;
            c.device = device;
            c.stream = stream;
            c.ev[0] = ev0; c.ev[1] = ev1; c.ev[2] = evk0; c.ev[3] = evk1;
            c.sm_count = sm_count;
            c.coop = coop_launch;
            exec_pool().give_back(c);
        }
    }

    void release_all() {
        model.reset();     // (if it was not handed to the cache) its device blocks go back to the block cache
        fb_deg.release(); fb_first.release(); fb_fill.release(); fb_outdeg.release(); fb_src.release(); fb_dst.release();
        fb_label.release(); fb_cset.release(); fb_sig.release(); fb_flags.release(); fb_failed.release(); fb_alive.release();
        fb_scan.release(); fb_fin.release(); fb_valid.release(); fb_palive.release(); fb_cover.release();
        d_jobs.release(); d_ctl.release(); frontier[0].release(); frontier[1].release();
        leaves.release(); unresolved.release(); gathered.release(); table.release(); state_key.release();
        edge_src.release(); edge_dst.release(); edge_label.release(); counters.release(); d_offsets.release();
    }

    template <class T>
    void upload(DBuf<T> &b, const std::vector<T> &v) {
        b.reserve(std::max<size_t>(v.size(), 1), 0, stream);
        if (!v.empty()) {
            CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
            h2d += (long long)(v.size() * sizeof(T));
        }
    }

    void upload_model() {
        const double t_up0 = now_s();
        upload(model->d_lb, model->sets.lb());
        upload(model->d_width, model->sets.width());
        upload(model->d_sigvars, model->sets.sig_vars());
        upload(model->d_sets, model->sets.dev_sets);
        upload(model->d_cons, model->sets.dev_cons);
        upload(model->d_props, model->sets.dev_props);
        upload(model->d_scope, model->sets.dev_scope);
        model->sets.dev_stride.resize(model->sets.dev_scope.size(), 0);
        upload(model->d_stride, model->sets.dev_stride);
        model->d_tables.reserve((size_t)std::max<long long>(model->sets.table_words, 1), (size_t)model->tables_built, stream);
        upload(model->d_code, model->sets.dev_code);
        upload(model->d_wake, model->sets.dev_wake);
        upload(model->d_aux, model->sets.dev_aux);
        upload(model->d_arr_off, model->sets.arr_off);
        upload(model->d_arr_val, model->sets.arr_val);
        CK(cudaStreamSynchronize(stream));      // the host vectors may be reallocated by the next set
        const double t_up = now_s();
        refresh_model();
        // fill the relation tables of constraints seen for the first time
        if (!model->sets.table_jobs.empty()) {
            std::vector<int32_t> jobs;
            int max_entries = 1;
            for (const TableJob &job : model->sets.table_jobs) {
                jobs.push_back(job.con);
                max_entries = std::max(max_entries, model->sets.dev_cons[job.con].table_entries);
            }
            upload(d_jobs, jobs);
            launch_build_tables(dm, d_jobs.p, (int)jobs.size(), max_entries, model->d_tables.p, stream);
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(stream));
            t_launches++;
        }
        if (opt.verbosity > 0)
            fprintf(stderr, "[stcsp r%d] upload_model: pools + copies %.2f ms, relation tables (%lld words) %.2f ms\n", rank,
                    (t_up - t_up0) * 1e3, (long long)model->sets.table_words, (now_s() - t_up) * 1e3);
        model->sets.table_jobs.clear();
        model->tables_built = model->sets.table_words;
        model->sets.clear_dirty();
        model->uploaded = true;
    }

    // dm <- the resident model (dimensions, device pointers, this solve's options)
    void refresh_model() {
        dm.V = model->sets.n_vars();
        dm.k = model->sets.k();
        dm.world = world;
        dm.rank = rank;
        dm.n_sig = (int32_t)model->sets.sig_vars().size();
        dm.sig_len = dm.n_sig + model->sets.n_until();
        dm.node_words = 4 + 2 * dm.V * dm.k;
        dm.rec_words = 4 + dm.V;
        dm.key_words = 1 + dm.sig_len;
        dm.max_scope = model->sets.max_scope();
        dm.max_stack = model->sets.max_stack();
        dm.max_words = (model->sets.max_props() + 31) / 32;
        {
            const size_t sb = model->sets.max_stage_bytes();
            dm.stage_bytes = sb <= 40 * 1024 ? (int32_t)sb : 0;     // larger model->sets stay in global memory / L1
        }
        dm.lazy_ahead = opt.lookahead == 2 ? 1 : 0;
        dm.multi_branch = opt.single_branch ? 0 : 1;
        // Measured (tools/fan_sweep.py, device ms): b6_f6_nosym 0.318 with one variable per node, 0.302 / 0.271 / 0.273 / 0.274 / 0.269
        // with room for 148 / 296 / 592 / 1184 / 3552 children per wave; the symmetric instances and digitinvader lose a few
        // per cent beyond 296 (children that fail at once still cost a node each).
        dm.fan_warps = sm_count * 2;
        {
            static const int fan_override = getenv("STCSP_FAN_WARPS") ? atoi(getenv("STCSP_FAN_WARPS")) : 0;   // tuning experiments
            if (fan_override > 0) dm.fan_warps = fan_override;
        }
        {
            static const int flags = getenv("STCSP_DBG_FLAGS") ? atoi(getenv("STCSP_DBG_FLAGS")) : 0;
            dm.dbg_flags = flags;
        }
        dm.n_sets = (int32_t)model->sets.n_sets();
        dm.scalar_walk = 192;
        dm.scalar_walk_cta = 32;
        {
            static const int cta_override = getenv("STCSP_SCALAR_WALK_CTA") ? atoi(getenv("STCSP_SCALAR_WALK_CTA")) : 0;     // tuning experiments
            if (cta_override > 0) dm.scalar_walk_cta = cta_override;
        }
        {
            static const int walk_override = getenv("STCSP_SCALAR_WALK") ? atoi(getenv("STCSP_SCALAR_WALK")) : 0;     // tuning experiments
            if (walk_override > 0) dm.scalar_walk = walk_override;
        }
        if (dm.V >= (1 << 10) - 1) dm.multi_branch = 0;        // the node header packs variable indices in ten bits
        // four node blocks per warp (quad mode for wide waves) when three CTAs still fit an SM
        dm.node_slots = 4 * kExpandWarps;
        if (expand_smem_bytes(dm) > (216 / kExpandCtasPerSm) * 1024) dm.node_slots = kExpandWarps;
        dm.force_mode = opt.expand_mode >= 1 && opt.expand_mode <= 3 ? opt.expand_mode : 0;
        dm.enum_now = opt.enum_limit_now > 0 ? opt.enum_limit_now : 8;
        dm.enum_ahead = opt.enum_limit_ahead > 0 ? opt.enum_limit_ahead : 4;
        dm.lb = model->d_lb.p;
        dm.width = model->d_width.p;
        dm.sig_vars = model->d_sigvars.p;
        dm.sets = model->d_sets.p;
        dm.cons = model->d_cons.p;
        dm.props = model->d_props.p;
        dm.scope = model->d_scope.p;
        dm.stride = model->d_stride.p;
        dm.tables = model->d_tables.p;
        dm.code = model->d_code.p;
        dm.wake = model->d_wake.p;
        dm.aux = model->d_aux.p;
        dm.arr_off = model->d_arr_off.p;
        dm.arr_val = model->d_arr_val.p;
        if (expand_smem_bytes(dm) > 200 * 1024)
            throw Failure(STCSP_ERR_UNSUPPORTED, "model needs more shared memory per CTA than an SM has");
        expand_grid_max = expand_max_grid(dm, sm_count);
    }

    void zero_wave_counters() {
        // all three sets of wave counters (search_kernel rotates through them; everything else uses set 0); the words behind
        // C_COUNT -- the look-ahead totals of set 0 -- live as long as the solve
        CK(cudaMemset2DAsync(counters.p + C_OUT, kCounterStride * sizeof(unsigned long long), 0,
                             (C_COUNT - C_OUT) * sizeof(unsigned long long), kCounterSets, stream));
    }
    void read_counters() {
        CK(cudaMemcpyAsync(h_counters, counters.p, C_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        if (opt.lookahead == 0)     // the look-ahead totals behind the counters (automatic policy, see expand())
            CK(cudaMemcpyAsync(h_counters + kHostAheadWord, counters.p + C_AHEAD_NODES, 2 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        d2h += C_COUNT * 8;
    }

    void ensure_table(long long states_needed) {
        long long want = table_size ? table_size : (1ll << 16);
        while (want < 2 * states_needed) want <<= 1;
        if (want == table_size) return;
        // a fresh table, every known state re-inserted
        DBuf<int32_t> fresh;
        fresh.reserve((size_t)want, 0, stream);
        launch_fill(fresh.p, want, -1, stream);
        // (the root's key starts with -1 unless its signature is empty, so it can sit in the table unreachable)
        if (n_states > 0 && !root_pending) {     // (root_pending: the only state is the root, which search_kernel inserts itself)
            launch_rehash(dm, fresh.p, want - 1, state_key.p, n_states, sm_count * 4, stream);
            t_launches++;
        }
        std::swap(table.p, fresh.p);
        std::swap(table.cap, fresh.cap);
        std::swap(table.bytes, fresh.bytes);
        std::swap(table.device, fresh.device);
        if (fresh.p) CK(cudaStreamSynchronize(stream));     // the old table returns to the cache when `fresh` goes out of scope
        table_size = want;
    }
    void init(const stcsp_problem_t *problem, const stcsp_options_t *options, int r, int w) {
        t_create = now_s();
        if (options) opt = *options;
        rank = r;
        world = w;
        if (w < 1 || w > kMaxWorld || r < 0 || r >= w) throw Failure(STCSP_ERR_INVALID, "bad rank / world size");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw Failure(STCSP_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                              "); this library has no CPU fallback");
        if (!opt.use_current_device) {
            device = opt.device >= 0 ? opt.device : 0;
            if (device >= ndev) throw Failure(STCSP_ERR_CUDA, "CUDA device ordinal out of range");
            CK(cudaSetDevice(device));
        } else {
            CK(cudaGetDevice(&device));
        }
        {
            const ExecContext c = exec_pool().acquire(device);
            stream = c.stream;
            ev0 = c.ev[0]; ev1 = c.ev[1]; evk0 = c.ev[2]; evk1 = c.ev[3];
            sm_count = c.sm_count;
            coop_launch = c.coop;
        }
        try {
            // Multi-rank solves number constraint sets in lock-step, so they start from a fresh state -- or from the state
            // the SAME group of ranks left behind for this model (group_tag): every rank of a group has solved the same
            // models in the same order, so their resident copies agree.
            if (w == 1 || !group_tag.empty()) {
                cache_key = model_key(*problem, device) + group_tag;
                model = model_cache().take(cache_key);
            }
            if (!model) {
                model.reset(new ModelState());
                model->sets.init(*problem);
            }
        } catch (const std::invalid_argument &ex) {
            throw Failure(STCSP_ERR_UNSUPPORTED, ex.what());
        } catch (const std::runtime_error &ex) {
            throw Failure(STCSP_ERR_INVALID, ex.what());
        }
        const double t_compiled = now_s();
        if (!model->uploaded || model->sets.dirty()) upload_model();
        else refresh_model();
        const double t_uploaded = now_s();
        counters.reserve(kCounterSets * kCounterStride, 0, stream);
        CK(cudaMemsetAsync(counters.p, 0, kCounterSets * kCounterStride * sizeof(unsigned long long), stream));
        counters_fresh = true;
        h_counters = pinned_cache().acquire();       // C_COUNT counters + room for the search control block
        h_counters[kHostAheadWord] = h_counters[kHostAheadWord + 1] = 0;     // (the block is recycled)
        {
            void *d = nullptr;
            static const bool no_push = getenv("STCSP_NO_PUSH") != nullptr;       // A/B timing
            if (!no_push) d = host_cache().device_ptr(h_counters);
            host_mirror = (unsigned long long *)d;
        }
        d_offsets.reserve(2 * kMaxWorld, 0, stream);

        const int NW = dm.node_words, KW = dm.key_words;
        // First solve of a model: pools that the shipped instances up to a few ten thousand states never outgrow (the blocks
        // come from the arena, so their size costs nothing; growing them in mid-search costs a kernel relaunch each time).
        const size_t f0 = (size_t)std::max<long long>(kFirstFrontier, model->hint_frontier),
                     s0 = (size_t)std::max<long long>(kFirstStates, model->hint_states),
                     e0 = (size_t)std::max<long long>(kFirstEdges, model->hint_edges);
        frontier[0].reserve(f0 * NW, 0, stream);
        frontier[1].reserve(f0 * NW, 0, stream);
        leaves.reserve(f0 * dm.rec_words, 0, stream);
        unresolved.reserve(f0, 0, stream);
        state_key.reserve(s0 * KW, 0, stream);
        edge_src.reserve(e0, 0, stream);
        edge_dst.reserve(e0, 0, stream);
        edge_label.reserve(e0 * dm.V, 0, stream);
        d_ctl.reserve(1, 0, stream);
        h_ctl = reinterpret_cast<SearchCtl *>(h_counters + C_COUNT);
        search_grid = search_max_grid(dm, sm_count);
        if (!coop_launch) search_grid = 0;
        if (rank == 0 && root_in_kernel && search_grid > 0 && world == 1) {
            root_pending = true;        // search_kernel writes the root state and its node itself (SearchArgs::make_root)
            n_states = 1;
            n_in = 1;
        } else if (rank == 0) {
            // root state (reference src/solveralgorithm.cpp:951-954) and its search node
            std::vector<int32_t> key(KW, 0);
            key[0] = dm.sig_len == 0 ? 0 : -1;
            CK(cudaMemcpyAsync(state_key.p, key.data(), KW * 4, cudaMemcpyHostToDevice, stream));
            std::vector<int32_t> node(NW, 0);
            node[3] = -1;
            for (int v = 0; v < dm.V; v++)
                for (int o = 0; o < dm.k; o++) {
                    const int w_ = model->sets.width()[v];
                    const unsigned long long m = w_ >= 64 ? ~0ull : ((1ull << w_) - 1ull);
                    memcpy(&node[4 + 2 * (v * dm.k + o)], &m, 8);
                }
            CK(cudaMemcpyAsync(frontier[0].p, node.data(), NW * 4, cudaMemcpyHostToDevice, stream));
            const unsigned long long one = 1;
            CK(cudaMemcpyAsync(counters.p + C_STATES, &one, 8, cudaMemcpyHostToDevice, stream));
            // (no synchronize: copies from pageable memory return once the source has been staged)
            h2d += (KW + NW) * 4 + 8;
            n_states = 1;
            n_in = 1;
        }
        ensure_table(std::max<long long>(std::max<long long>(n_states, 1), model->hint_table / 2));
        if (opt.verbosity > 0)
            fprintf(stderr, "[stcsp r%d] init: context + model lookup / compile %.2f ms, upload + relation tables %.2f ms, pools + root %.2f ms\n",
                    rank, (t_compiled - t_create) * 1e3, (t_uploaded - t_compiled) * 1e3, (now_s() - t_uploaded) * 1e3);
    }

    // Pools big enough for a wave over `nin` input nodes (see the SEARCH_GROW test in search_kernel).
    void ensure_wave_capacity(long long nin) {
        const int NW = dm.node_words, KW = dm.key_words, V = dm.V, RW = dm.rec_words;
        leaves.reserve((size_t)nin * RW, 0, stream);
        unresolved.reserve((size_t)nin, 0, stream);
        frontier[cur].reserve((size_t)2 * nin * NW, (size_t)n_in * NW, stream);
        frontier[cur ^ 1].reserve((size_t)2 * nin * NW, 0, stream);
        state_key.reserve((size_t)(n_states + nin) * KW, (size_t)n_states * KW, stream);
        edge_src.reserve((size_t)(n_edges + nin), (size_t)n_edges, stream);
        edge_dst.reserve((size_t)(n_edges + nin), (size_t)n_edges, stream);
        edge_label.reserve((size_t)(n_edges + nin) * V, (size_t)n_edges * V, stream);
        ensure_table(n_states + nin);
    }

    // The wave loop in one cooperative launch (search_kernel); the host only steps in when the kernel asks.
    void run_persistent(double deadline) {
        begin_timing();
        const int NW = dm.node_words, KW = dm.key_words, V = dm.V, RW = dm.rec_words;
        while (n_in > 0) {
            if (deadline > 0 && now_s() > deadline) throw Failure(STCSP_ERR_TIMEOUT, "time limit reached");
            if (opt.max_frontier_nodes > 0 && n_in > opt.max_frontier_nodes)
                throw Failure(STCSP_ERR_CAPACITY, "frontier wider than max_frontier_nodes");
            SearchArgs sa{};
            sa.ctl = d_ctl.p;
            sa.counters = counters.p;
            sa.frontier[0] = frontier[0].p;
            sa.frontier[1] = frontier[1].p;
            sa.out_cap = (long long)(std::min(frontier[0].cap, frontier[1].cap) / NW);
            sa.leaves = leaves.p;
            sa.leaf_cap = (long long)(leaves.cap / RW);
            sa.unresolved = unresolved.p;
            sa.unresolved_cap = (long long)unresolved.cap;
            sa.capmap = model->d_capmap.p;
            sa.capvals = model->d_capvals.p;
            sa.capmap_mask = model->capmap.empty() ? -1 : (int32_t)model->capmap.size() - 1;
            sa.table = table.p;
            sa.table_mask = table_size - 1;
            sa.state_key = state_key.p;
            sa.state_cap = (long long)(state_key.cap / KW);
            sa.edge_src = edge_src.p;
            sa.edge_dst = edge_dst.p;
            sa.edge_label = edge_label.p;
            sa.edge_cap = (long long)std::min(edge_src.cap, edge_label.cap / (size_t)V);
            sa.fuse_leaves = (dm.dbg_flags & 1) ? 0 : 1;        // STCSP_DBG_FLAGS=1: the separate leaf phase on every wave (A/B timing)
            sa.ahead_policy = opt.lookahead == 3 ? 2 : opt.lookahead == 0 ? 1 : 0;     // 0 (default): automatic, see search_kernel
            sa.leaves0 = t_leaves;
            sa.waves0 = t_waves;
            const long long wide = opt.wide_wave_nodes < 0 ? 0 : opt.wide_wave_nodes > 0 ? opt.wide_wave_nodes : kWideWaveNodes;
            sa.max_frontier = opt.max_frontier_nodes > 0 && (wide == 0 || opt.max_frontier_nodes < wide) ? opt.max_frontier_nodes : wide;
            // An instance whose waves stay narrow gets one CTA per SM: with a third of the CTAs the grid barrier is cheaper
            // and a node's CTA has the SM to itself (b6_nosym -7 %, b5_f6 -8 %); anything wider needs all resident warps
            // (digitinvader7 with 148 CTAs: +22 %).  Every solve STARTS narrow unless the last solve of this model is known
            // to have been wide; the kernel yields at the first wave wider than kNarrowInstanceWave and is relaunched on the
            // full grid (one relaunch per solve, which a wide instance does not notice).
            const bool narrow = !wide_grid && search_grid > sm_count &&
                                !(model->hint_max_wave > kNarrowInstanceWave) && n_in <= kNarrowInstanceWave;
            if (narrow && (sa.max_frontier == 0 || sa.max_frontier > kNarrowInstanceWave)) sa.max_frontier = kNarrowInstanceWave;
            // automata known (from the last solve of this model) to be small are finished inside the kernel
            if (finish_in_kernel && std::max<long long>(model->hint_edges, (long long)edge_src.cap) <= (1ll << 18)) {
                const long long cs = std::max<long long>((long long)(state_key.cap / KW), 4096), ce = std::max<long long>((long long)edge_src.cap, 8192);
                const int SLm = std::max(dm.sig_len, 1);
                fb_deg.reserve((size_t)cs + 1, 0, stream);
                fb_first.reserve((size_t)cs + 1, 0, stream);
                fb_fill.reserve((size_t)cs + 1, 0, stream);
                fb_outdeg.reserve((size_t)cs + 1, 0, stream);
                fb_failed.reserve((size_t)cs + 1, 0, stream);
                fb_alive.reserve((size_t)ce + 1, 0, stream);
                fb_src.reserve((size_t)ce + 1, 0, stream);
                fb_dst.reserve((size_t)ce + 1, 0, stream);
                fb_label.reserve((size_t)ce * V + 1, 0, stream);
                fb_cset.reserve((size_t)cs + 1, 0, stream);
                fb_sig.reserve((size_t)cs * SLm + 1, 0, stream);
                sa.fin.deg = fb_deg.p; sa.fin.first = fb_first.p; sa.fin.cursor = fb_fill.p; sa.fin.outdeg = fb_outdeg.p;
                sa.fin.failed = fb_failed.p; sa.fin.alive = fb_alive.p;
                sa.fin.s_src = fb_src.p; sa.fin.s_dst = fb_dst.p; sa.fin.s_label = fb_label.p;
                sa.fin.rows_cset = fb_cset.p; sa.fin.rows_sig = fb_sig.p;
                sa.fin.cap_states = cs; sa.fin.cap_edges = ce;
                sa.fin.do_trim = finish_trim ? 1 : 0;
                // out-degrees counted while the edges are appended: from the first launch of the solve on, as long as these two
                // blocks stay where they are and every edge is appended by the search kernel itself
                if (sa_first_launch_of_solve()) { deg_ok = true; deg_ptr = fb_deg.p; cur_ptr = fb_fill.p; deg_cs = cs; }
                else if (deg_ptr != fb_deg.p || cur_ptr != fb_fill.p || deg_cs != cs) deg_ok = false;
                sa.fin.deg_counted = deg_ok ? 1 : 0;
                // PUSH: the last solve of this model returned a small automaton -- pinned output buffers of that size (plus a
                // margin) go to the kernel, which writes the result into them itself (see FinishArgs)
                const long long hs = model->hint_out_states, he = model->hint_out_edges;
                if (host_mirror && hs > 0 && he * (V + 2) * 4 + hs * (SLm + 2) * 4 <= (4ll << 20)) {
                    if (!push_store) {
                        push_cap_states = hs + hs / 8 + 16;
                        push_cap_edges = he + he / 8 + 16;
                        auto *st = new PinnedStore();
                        try {
                            st->sig_vars.alloc(model->sets.sig_vars().size() * 4);
                            st->state_sig.alloc((size_t)push_cap_states * SLm * 4);
                            st->state_cset.alloc((size_t)push_cap_states * 4);
                            st->state_failed.alloc((size_t)push_cap_states);
                            st->edge_src.alloc((size_t)push_cap_edges * 4);
                            st->edge_dst.alloc((size_t)push_cap_edges * 4);
                            st->edge_label.alloc((size_t)push_cap_edges * V * 4);
                        } catch (...) { delete st; throw; }
                        const bool all_pinned = st->state_sig.pinned() && st->state_cset.pinned() && st->state_failed.pinned() &&
                                                st->edge_src.pinned() && st->edge_dst.pinned() && st->edge_label.pinned();
                        if (all_pinned) push_store = st; else delete st;
                    }
                    if (push_store) {
                        auto dev = [&](void *h) { return host_cache().device_ptr(h); };
                        sa.fin.h_sig = (int32_t *)dev(push_store->state_sig.p);
                        sa.fin.h_cset = (int32_t *)dev(push_store->state_cset.p);
                        sa.fin.h_failed = (uint8_t *)dev(push_store->state_failed.p);
                        sa.fin.h_src = (int32_t *)dev(push_store->edge_src.p);
                        sa.fin.h_dst = (int32_t *)dev(push_store->edge_dst.p);
                        sa.fin.h_label = (int32_t *)dev(push_store->edge_label.p);
                        sa.fin.h_cap_states = push_cap_states;
                        sa.fin.h_cap_edges = push_cap_edges;
                        if (!sa.fin.h_sig || !sa.fin.h_cset || !sa.fin.h_failed || !sa.fin.h_src || !sa.fin.h_dst || !sa.fin.h_label) {
                            cudaGetLastError();
                            sa.fin.h_src = nullptr;        // no mapping: the classic downloads
                        }
                    }
                }
            }
            DBuf<unsigned long long> trace;
            const long long trace_waves = 256;
            if (opt.verbosity > 2) {
                trace.reserve((size_t)trace_waves * 5 + 4096, 0, stream);
                CK(cudaMemsetAsync(trace.p, 0, ((size_t)trace_waves * 5 + 4096) * 8, stream));
                sa.trace = trace.p;
                sa.trace_cap = trace_waves;
            }
            memset(h_ctl, 0, sizeof *h_ctl);
            sa.n_in0 = n_in;
            sa.cur0 = cur;
            sa.waves_left0 = deadline > 0 ? 256 : (1ll << 40);
            sa.make_root = root_pending ? 1 : 0;
            root_pending = false;
            if (!counters_fresh) zero_wave_counters();
            counters_fresh = false;
            const int grid = narrow ? std::min(search_grid, sm_count) : search_grid;
            sa.h_ctl = host_mirror ? reinterpret_cast<SearchCtl *>(host_mirror + C_COUNT) : nullptr;
            sa.h_counters = host_mirror;
            CK(cudaEventRecord(evk0, stream));
            CK(launch_search(dm, sa, grid, sm_count, stream));
            CK(cudaEventRecord(evk1, stream));
            h2d += sizeof dm + sizeof sa;       // the launch parameters (model descriptor, start values) are what goes down per launch
            if (host_mirror) {
                // the kernel wrote its control block and counter set 0 into the pinned block itself: one synchronisation
                CK(cudaStreamSynchronize(stream));
                d2h += C_COUNT * 8;
            } else {
                CK(cudaMemcpyAsync(h_ctl, d_ctl.p, sizeof *h_ctl, cudaMemcpyDeviceToHost, stream));
                read_counters();
            }
            {
                float kms = 0;
                CK(cudaEventElapsedTime(&kms, evk0, evk1));
                expand_ms += kms;
                t_expand_launches++;
            }
            if (opt.verbosity > 2) {
                std::vector<unsigned long long> tr((size_t)trace_waves * 5 + 4096);
                CK(cudaMemcpy(tr.data(), trace.p, tr.size() * 8, cudaMemcpyDeviceToHost));
                {
                    // block 0's timeline in CTA mode: tags 0 wave entered, 1 node loaded, 10 scalar round, 20+n cooperative
                    // round over n propagators, 2 node propagated
                    const unsigned long long *d = tr.data() + trace_waves * 5;
                    const unsigned long long cnt = std::min<unsigned long long>(d[0], 2040);
                    std::string line;
                    for (unsigned long long i = 0; i < cnt; i++) {
                        const unsigned long long tag = d[1 + 2 * i], t = d[2 + 2 * i];
                        const double dt = i ? (double)(t - d[2 * i]) / 1e3 : 0.0;
                        char buf[64];
                        snprintf(buf, sizeof buf, tag == 0 ? "\n[stcsp] block0: wave" : " %llu:+%.1f", tag, dt);
                        line += buf;
                    }
                    fprintf(stderr, "%s\n", line.c_str());
                }
                for (long long w = 0; w < std::min<long long>(trace_waves, h_ctl->t_waves); w++)
                    fprintf(stderr, "[stcsp] kernel wave %lld: expand %.1f us, route %.1f us, ingest %.1f us, bookkeeping %.1f us, gap to next %.1f us\n",
                            w, (tr[w * 5 + 1] - tr[w * 5]) / 1e3, (tr[w * 5 + 2] - tr[w * 5 + 1]) / 1e3,
                            (tr[w * 5 + 3] - tr[w * 5 + 2]) / 1e3, (tr[w * 5 + 4] - tr[w * 5 + 3]) / 1e3,
                            w + 1 < h_ctl->t_waves ? (tr[w * 5 + 5] - tr[w * 5 + 4]) / 1e3 : 0.0);
                CK(cudaStreamSynchronize(stream));
            }
            t_launches++;
            d2h += sizeof *h_ctl;
            t_nodes += h_ctl->t_nodes;
            t_fails += h_ctl->t_fails;
            t_tuples += h_ctl->t_tuples;
            t_revisions += h_ctl->t_revisions;
            t_dominance += h_ctl->t_dominance;
            t_leaves += h_ctl->t_leaves;
            t_waves += h_ctl->t_waves;
            max_wave = std::max<long long>(max_wave, h_ctl->t_max_in);
            if (opt.verbosity > 0)
                fprintf(stderr, "[stcsp r%d] t=%.3f ms search kernel returned: status %d after %lld waves, n_in %lld, states %llu, edges %llu, "
                                "out %llu leaves %llu overflow %d finished %d (barrier-free %d) pushed %d\n",
                        rank, (now_s() - t_create) * 1e3, h_ctl->status, h_ctl->t_waves, h_ctl->n_in, h_counters[C_STATES],
                        h_counters[C_EDGES], h_counters[C_OUT], h_counters[C_LEAVES], h_ctl->overflow, h_ctl->finished, h_ctl->pad,
                        h_ctl->pushed);
            if (h_ctl->overflow & ~1) { poisoned = true; throw Failure(STCSP_ERR_CAPACITY, "internal: a pool overflowed inside the search kernel"); }
            n_in = h_ctl->n_in;
            cur = h_ctl->cur;
            n_states = (long long)h_counters[C_STATES];
            n_edges = (long long)h_counters[C_EDGES];
            switch (h_ctl->status) {
                case SEARCH_DONE:
                    n_in = 0;
                    prefinished = h_ctl->finished != 0;
                    prefinished_dead = h_ctl->dead_edges;
                    pushed = prefinished && h_ctl->pushed != 0 && sa.fin.h_src != nullptr;
                    break;
                case SEARCH_YIELD:
                    if (narrow && n_in > kNarrowInstanceWave) wide_grid = true;     // from here on: the full grid
                    // Waves this wide run faster as separate launches: the stand-alone expand kernels keep their inner
                    // loops in registers (inside search_kernel ptxas spills there) and route/ingest run at full occupancy;
                    // at this width the launches and the two host round trips per wave no longer matter.
                    while (wide > 0 && n_in > wide && !(opt.max_frontier_nodes > 0 && n_in > opt.max_frontier_nodes)) {
                        if (deadline > 0 && now_s() > deadline) throw Failure(STCSP_ERR_TIMEOUT, "time limit reached");
                        int64_t nl = 0, np = 0, next = 0;
                        time_expand = true;
                        expand(&nl, &np);
                        time_expand = false;
                        if (np > 0) {
                            std::vector<int32_t> req = pending;
                            resolve(req.data(), np);
                        }
                        ingest(nullptr, 0, &next);
                        zero_wave_counters();
                    }
                    break;
                case SEARCH_GROW:
                    ensure_wave_capacity(n_in);
                    break;
                case SEARCH_RETRY:
                    frontier[cur ^ 1].reserve((size_t)std::max<unsigned long long>(h_counters[C_OUT] + h_counters[C_OUT] / 4,
                                                                                  2 * (frontier[cur ^ 1].cap / NW)) * NW, 0, stream);
                    frontier[cur].reserve(frontier[cur ^ 1].cap, (size_t)n_in * NW, stream);
                    zero_wave_counters();
                    break;
                case SEARCH_INGEST:
                case SEARCH_RESOLVE: {
                    // the device stopped in the middle of a wave; finish it here
                    const bool routed = h_ctl->status == SEARCH_RESOLVE;
                    if (!routed) {                      // expand done, nothing routed yet
                        RouteArgs ra{};
                        ra.leaves = leaves.p;
                        ra.capmap = model->d_capmap.p;
                        ra.capvals = model->d_capvals.p;
                        ra.capmap_mask = model->capmap.empty() ? -1 : (int32_t)model->capmap.size() - 1;
                        ra.unresolved = unresolved.p;
                        ra.unresolved_cap = (long long)unresolved.cap;
                        ra.counters = counters.p;
                        launch_route(dm, ra, (int)std::min<long long>((n_in + 7) / 8, sm_count * 8), stream);
                        CK(cudaGetLastError());
                        read_counters();
                        t_launches++;
                    }
                    n_out = (long long)(h_counters[C_OUT] + h_counters[C_NEW]);
                    n_leaves = (long long)h_counters[C_LEAVES];
                    n_unres = (long long)h_counters[C_UNRESOLVED];
                    t_nodes += (long long)h_counters[C_NODES];
                    t_fails += (long long)h_counters[C_FAILS];
                    t_tuples += (long long)h_counters[C_TUPLES];
                    t_revisions += (long long)h_counters[C_REVISIONS];
                    t_leaves += n_leaves;
                    t_waves++;
                    pending.clear();
                    if (n_unres > 0) {
                        const long long listed = n_unres;
                        collect_pending();
                        std::vector<int32_t> req = pending;
                        resolve(req.data(), (int64_t)(req.size() / (size_t)(1 + V)));     // re-routes the listed leaves in place
                        if (routed) {
                            // everything else of this wave is in the automaton already: only the listed leaves remain
                            t_dominance += (long long)h_counters[C_DOMINANCE];
                            gathered.reserve((size_t)listed * RW, 0, stream);
                            launch_gather(leaves.p, unresolved.p, listed, RW, gathered.p,
                                          (int)std::min<long long>((listed + 7) / 8, sm_count * 8), stream);
                            t_launches++;
                            CK(cudaMemsetAsync(counters.p + C_DOMINANCE, 0, sizeof(unsigned long long), stream));
                            n_out = (long long)(h_counters[C_OUT] + h_counters[C_NEW]);
                            n_states = (long long)h_counters[C_STATES];
                            n_edges = (long long)h_counters[C_EDGES];
                            int64_t next = 0;
                            ingest(gathered.p, listed, &next);
                            zero_wave_counters();
                            break;
                        }
                    } else if (routed) {
                        throw Failure(STCSP_ERR_CUDA, "internal: SEARCH_RESOLVE without pending leaves");
                    }
                    int64_t next = 0;
                    ingest(nullptr, 0, &next);
                    zero_wave_counters();
                    break;
                }
                default:
                    throw Failure(STCSP_ERR_CUDA, "search kernel returned an unknown status");
            }
        }
    }

    void begin_timing() {
        if (!timing_open) {
            CK(cudaEventRecord(ev0, stream));
            timing_open = true;
        }
    }

    void expand(int64_t *out_leaves, int64_t *out_pending) {
        if (poisoned) throw Failure(STCSP_ERR_CAPACITY, "session unusable: a pool overflowed in an earlier wave");
        begin_timing();
        pending.clear();
        n_leaves = n_unres = 0;
        n_out = 0;
        zero_wave_counters();
        for (int i = C_OUT; i < C_COUNT; i++) h_counters[i] = 0;       // host mirror of the wave counters
        if (n_in > 0) {
            const int NW = dm.node_words, RW = dm.rec_words;
            const double te0 = now_s();
            max_wave = std::max(max_wave, n_in);
            leaves.reserve((size_t)n_in * RW, 0, stream);
            unresolved.reserve((size_t)n_in, 0, stream);
            DBuf<int32_t> &out = frontier[cur ^ 1];
            out.reserve((size_t)std::max<long long>(2 * n_in, 4096) * NW, 0, stream);
            if (opt.verbosity > 1) {
                CK(cudaStreamSynchronize(stream));
                fprintf(stderr, "[stcsp r%d]   expand reserve %.1f us\n", rank, (now_s() - te0) * 1e6);
            }
            for (;;) {
                ExpandArgs ea{};
                ea.in_nodes = frontier[cur].p;
                ea.n_in = n_in;
                ea.out_nodes = out.p;
                ea.out_cap = (long long)(out.cap / NW);
                ea.leaves = leaves.p;
                ea.leaf_cap = (long long)(leaves.cap / RW);
                ea.fan = branch_fan(dm, n_in, ea.out_cap);
                ea.counters = counters.p;
                // look-ahead propagators: the same automatic policy as inside search_kernel, from the totals of the last read-back
                ea.ahead_stats = opt.lookahead == 0 ? counters.p : nullptr;
                ea.skip_ahead = opt.lookahead == 3 ||
                                (opt.lookahead == 0 && !ahead_sample_wave(t_waves) &&
                                 ahead_droppable((long long)h_counters[kHostAheadWord], (long long)h_counters[kHostAheadWord + 1], t_leaves));
                // narrow wave: a whole CTA per node (intra-node parallelism); wide wave: a warp per node
                const int mode = pick_expand_mode(dm, n_in, expand_grid_max);
                const int grid = mode == EXPAND_CTA ? (int)std::min<long long>(n_in, expand_grid_max)
                               : mode == EXPAND_QUAD ? (int)std::min<long long>((n_in + 4 * kExpandWarps - 1) / (4 * kExpandWarps), expand_grid_max)
                                                     : (int)std::min<long long>((n_in + kExpandWarps - 1) / kExpandWarps, expand_grid_max);
                if (opt.profile_kernels || time_expand) CK(cudaEventRecord(evk0, stream));
                launch_expand(dm, ea, grid, mode, stream);
                if (opt.profile_kernels || time_expand) CK(cudaEventRecord(evk1, stream));
                RouteArgs ra{};
                ra.leaves = leaves.p;
                ra.list = nullptr;
                ra.capmap = model->d_capmap.p;
                ra.capvals = model->d_capvals.p;
                ra.capmap_mask = model->capmap.empty() ? -1 : (int32_t)model->capmap.size() - 1;
                ra.unresolved = unresolved.p;
                ra.unresolved_cap = (long long)unresolved.cap;
                ra.counters = counters.p;
                const double te1 = now_s();
                if (opt.verbosity > 1) CK(cudaStreamSynchronize(stream));
                const double te2 = now_s();
                launch_route(dm, ra, (int)std::min<long long>((n_in + 7) / 8, sm_count * 8), stream);
                CK(cudaGetLastError());
                read_counters();
                if (opt.verbosity > 1)
                    fprintf(stderr, "[stcsp r%d]   expand kernel wall %.1f us, route+read %.1f us\n", rank, (te2 - te1) * 1e6, (now_s() - te2) * 1e6);
                t_launches += 2;
                t_expand_launches++;
                if (opt.profile_kernels || time_expand) {
                    float ms = 0;
                    CK(cudaEventElapsedTime(&ms, evk0, evk1));
                    expand_ms += ms;
                    if (opt.verbosity > 0)
                        fprintf(stderr, "[stcsp r%d] t=%.3f ms wave %lld: in %lld out %llu leaves %llu fails %llu tuples %llu revisions %llu expand %.1f us grid %d\n",
                                rank, (now_s() - t_create) * 1e3, t_waves, n_in, h_counters[C_OUT], h_counters[C_LEAVES], h_counters[C_FAILS],
                                h_counters[C_TUPLES], h_counters[C_REVISIONS], ms * 1e3, grid);
                }
                const unsigned long long ov = h_counters[C_OVERFLOW];
                if (ov & 1ull) {        // frontier buffer too small: nothing but scratch was written, run the wave again
                    out.reserve((size_t)std::max<unsigned long long>(h_counters[C_OUT] + h_counters[C_OUT] / 4,
                                                                     2 * (out.cap / NW)) * NW, 0, stream);
                    zero_wave_counters();
                    continue;
                }
                if (ov) { poisoned = true; throw Failure(STCSP_ERR_CAPACITY, "internal: leaf buffers overflowed"); }
                break;
            }
            n_out = (long long)(h_counters[C_OUT] + h_counters[C_NEW]);
            n_leaves = (long long)h_counters[C_LEAVES];
            n_unres = (long long)h_counters[C_UNRESOLVED];
            t_nodes += (long long)h_counters[C_NODES];
            t_fails += (long long)h_counters[C_FAILS];
            t_tuples += (long long)h_counters[C_TUPLES];
            t_revisions += (long long)h_counters[C_REVISIONS];
            t_leaves += n_leaves;
            t_waves++;
            if (n_unres > 0) collect_pending();
        }
        if (out_leaves) *out_leaves = n_leaves;
        if (out_pending) *out_pending = (int64_t)(pending.size() / (size_t)(1 + dm.V));
    }

    std::vector<int32_t> cap_key(int32_t cid, const int32_t *values) const {
        const HostSet &hs = model->sets.host_set(cid);
        std::vector<int32_t> key;
        key.reserve(1 + hs.cap_vars.size());
        key.push_back(cid);
        for (int32_t v : hs.cap_vars) key.push_back(values[v]);
        return key;
    }

    void collect_pending() {
        const int RW = dm.rec_words, V = dm.V;
        gathered.reserve((size_t)n_unres * RW, 0, stream);
        launch_gather(leaves.p, unresolved.p, n_unres, RW, gathered.p, (int)std::min<long long>((n_unres + 7) / 8, sm_count * 8), stream);
        t_launches++;
        std::vector<int32_t> host((size_t)n_unres * RW);
        CK(cudaMemcpyAsync(host.data(), gathered.p, host.size() * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        d2h += (long long)host.size() * 4;
        std::map<std::vector<int32_t>, int> seen;
        for (long long i = 0; i < n_unres; i++) {
            const int32_t *rec = host.data() + i * RW;
            std::vector<int32_t> key = cap_key(rec[1], rec + 4);
            if (model->cap_lookup.count(key) || !seen.emplace(key, 1).second) continue;
            pending.push_back(rec[1]);
            pending.insert(pending.end(), rec + 4, rec + 4 + V);
        }
    }

    void capmap_insert(int32_t cid, const std::vector<int32_t> &key, int32_t next) {
        if ((model->capmap_used + 1) * 2 > (long long)model->capmap.size()) {
            std::vector<CapEntry> old;
            old.swap(model->capmap);
            CapEntry empty{};
            empty.cid = -1;
            model->capmap.assign(std::max<size_t>(64, old.size() * 2), empty);
            model->capmap_used = 0;
            for (const CapEntry &e : old)
                if (e.cid != -1) place(e);
        }
        CapEntry e{};
        e.cid = cid;
        e.next = next;
        e.off = (int32_t)model->capvals.size();
        model->capvals.insert(model->capvals.end(), key.begin() + 1, key.end());
        place(e);
    }
    void place(const CapEntry &e) {
        const int n = (int)model->sets.host_set(e.cid).cap_vars.size();
        uint32_t h = capmap_hash(e.cid, model->capvals.data() + e.off, n) & (uint32_t)(model->capmap.size() - 1);
        while (model->capmap[h].cid != -1) h = (h + 1) & (uint32_t)(model->capmap.size() - 1);
        model->capmap[h] = e;
        model->capmap_used++;
    }

    void resolve(const int32_t *requests, int64_t n_req) {
        const int V = dm.V;
        bool added = false;
        for (int64_t i = 0; i < n_req; i++) {
            const int32_t *rq = requests + i * (1 + V);
            const int32_t cid = rq[0];
            if (cid < 0 || cid >= model->sets.n_sets()) throw Failure(STCSP_ERR_INVALID, "resolve request names an unknown constraint set");
            std::vector<int32_t> key = cap_key(cid, rq + 1);
            if (model->cap_lookup.count(key)) continue;
            int32_t next;
            try {
                next = model->sets.successor(cid, rq + 1);
            } catch (const std::invalid_argument &ex) {
                throw Failure(STCSP_ERR_UNSUPPORTED, ex.what());
            }
            model->cap_lookup[key] = next;
            capmap_insert(cid, key, next);
            added = true;
        }
        if (model->sets.dirty()) upload_model();
        if (added) {
            model->d_capmap.reserve(model->capmap.size(), 0, stream);
            CK(cudaMemcpyAsync(model->d_capmap.p, model->capmap.data(), model->capmap.size() * sizeof(CapEntry), cudaMemcpyHostToDevice, stream));
            upload(model->d_capvals, model->capvals);
            CK(cudaStreamSynchronize(stream));
            h2d += (long long)(model->capmap.size() * sizeof(CapEntry));
        }
        if (n_unres > 0) {
            CK(cudaMemsetAsync(counters.p + C_UNRESOLVED, 0, sizeof(unsigned long long), stream));
            RouteArgs ra{};
            ra.leaves = leaves.p;
            ra.list = unresolved.p;
            ra.count = n_unres;
            ra.capmap = model->d_capmap.p;
            ra.capvals = model->d_capvals.p;
            ra.capmap_mask = model->capmap.empty() ? -1 : (int32_t)model->capmap.size() - 1;
            ra.unresolved = gathered.p;         // scratch: nothing may remain unresolved
            ra.unresolved_cap = (long long)gathered.cap;
            ra.counters = counters.p;
            launch_route(dm, ra, (int)std::min<long long>((n_unres + 7) / 8, sm_count * 8), stream);
            CK(cudaGetLastError());
            read_counters();
            t_launches++;
            if (h_counters[C_UNRESOLVED] != 0)
                throw Failure(STCSP_ERR_INVALID, "resolve: the request list did not cover every pending leaf of this rank");
            n_unres = 0;
        }
        pending.clear();
    }

    // wait: the records are in place when this returns (the step-wise C API promises that); the group's wave loop passes
    // false -- its exchange kernel follows on the same stream, and that is all the ordering the peers need
    void outbox(int32_t *outbox_dev, int64_t capacity, int64_t *counts, bool wait = true) {
        if (n_unres > 0) throw Failure(STCSP_ERR_INVALID, "outbox called with unresolved leaves pending");
        long long off[2 * kMaxWorld] = {0};
        long long total = 0;
        for (int q = 0; q < world; q++) {
            counts[q] = world == 1 ? (int64_t)n_leaves : (int64_t)h_counters[C_OWNER0 + q];     // (one rank: route_leaf does not count owners)
            off[q] = total;
            total += counts[q];
        }
        if (total != n_leaves) throw Failure(STCSP_ERR_INVALID, "internal: owner counts do not add up to the leaf count");
        if (total > capacity) throw Failure(STCSP_ERR_CAPACITY, "outbox too small");
        if (total == 0) return;
        CK(cudaMemcpyAsync(d_offsets.p, off, sizeof off, cudaMemcpyHostToDevice, stream));   // [world..2*world) stay 0: fill
        launch_scatter(dm, leaves.p, n_leaves, d_offsets.p, reinterpret_cast<unsigned long long *>(d_offsets.p + kMaxWorld),
                       outbox_dev, (int)std::min<long long>((n_leaves + 7) / 8, sm_count * 8), stream);
        CK(cudaGetLastError());
        if (wait) CK(cudaStreamSynchronize(stream));
        t_launches++;
    }

    struct PullSegs {               // the records of one wave where their producers left them (see IngestArgs)
        int n = 0;
        const int32_t *base[kMaxWorld];
        long long count[kMaxWorld];
    };
    void ingest(const int32_t *inbox, int64_t n, int64_t *frontier_next, const PullSegs *segs = nullptr) {
        if (poisoned) throw Failure(STCSP_ERR_CAPACITY, "session unusable: a pool overflowed in an earlier wave");
        if (n_unres > 0) throw Failure(STCSP_ERR_INVALID, "ingest called with unresolved leaves pending");
        deg_ok = false;         // edges appended by a stand-alone launch are not counted in fb_deg (FinishArgs::deg_counted)
        const int NW = dm.node_words, KW = dm.key_words, V = dm.V;
        if (segs) {
            n = 0;
            for (int q = 0; q < segs->n; q++) n += segs->count[q];
            if ((size_t)32 * dm.rec_words * 4 > 48 * 1024)
                throw Failure(STCSP_ERR_UNSUPPORTED, "models with more than 380 variables cannot be sharded (record staging)");
        } else if (!inbox) {
            n = world == 1 ? n_leaves : 0;               // single rank: this rank's own leaves, in place
        }
        const int32_t *records = segs ? nullptr : (inbox ? inbox : leaves.p);
        DBuf<int32_t> &out = frontier[cur ^ 1];
        const double tw0 = now_s();
        double tw1 = tw0, tw2 = tw0;
        if (n > 0) {
            out.reserve((size_t)(n_out + n) * NW, (size_t)n_out * NW, stream);
            state_key.reserve((size_t)(n_states + n) * KW, (size_t)n_states * KW, stream);
            edge_src.reserve((size_t)(n_edges + n), (size_t)n_edges, stream);
            edge_dst.reserve((size_t)(n_edges + n), (size_t)n_edges, stream);
            edge_label.reserve((size_t)(n_edges + n) * V, (size_t)n_edges * V, stream);
            ensure_table(n_states + n);
            tw1 = now_s();
            IngestArgs ia{};
            ia.records = records;
            ia.count = n;
            ia.table = table.p;
            ia.table_mask = table_size - 1;
            ia.state_key = state_key.p;
            ia.state_cap = (long long)(state_key.cap / KW);
            ia.edge_src = edge_src.p;
            ia.edge_dst = edge_dst.p;
            ia.edge_label = edge_label.p;
            ia.edge_cap = (long long)std::min(edge_src.cap, edge_label.cap / (size_t)V);
            ia.out_nodes = out.p;
            ia.out_base = (long long)h_counters[C_OUT];         // what expand wrote; C_NEW counts on from earlier ingests of this wave
            ia.out_cap = (long long)(out.cap / NW);
            ia.counters = counters.p;
            ia.totals = counters.p;
            if (segs) {
                ia.n_segs = segs->n;
                for (int q = 0; q < segs->n; q++) { ia.seg_base[q] = segs->base[q]; ia.seg_count[q] = segs->count[q]; }
            }
            launch_ingest(dm, ia, (int)std::min<long long>((n + 7) / 8, sm_count * 8), stream);
            CK(cudaGetLastError());
            read_counters();
            t_launches++;
            if (h_counters[C_OVERFLOW]) { poisoned = true; throw Failure(STCSP_ERR_CAPACITY, "internal: automaton pools overflowed during ingest"); }
            n_out = (long long)(h_counters[C_OUT] + h_counters[C_NEW]);
            n_states = (long long)h_counters[C_STATES];
            n_edges = (long long)h_counters[C_EDGES];
            t_dominance += (long long)h_counters[C_DOMINANCE];
            tw2 = now_s();
        }
        if (opt.verbosity > 1)
            fprintf(stderr, "[stcsp r%d]   ingest %lld records: reserve %.1f us, kernel+sync %.1f us, states %lld edges %lld table %lld\n",
                    rank, (long long)n, (tw1 - tw0) * 1e6, (tw2 - tw1) * 1e6, n_states, n_edges, table_size);
        cur ^= 1;
        n_in = n_out;
        n_out = 0;
        n_leaves = 0;
        if (frontier_next) *frontier_next = n_in;
    }

    void finish(stcsp_automaton_t *part) {
        memset(part, 0, sizeof *part);
        float ms = 0;
        if (timing_open) {
            CK(cudaEventRecord(ev1, stream));
            CK(cudaEventSynchronize(ev1));
            CK(cudaEventElapsedTime(&ms, ev0, ev1));
        }
        const int KW = dm.key_words, V = dm.V, SL = dm.sig_len;
        auto *st = new AutoStore();
        std::vector<int32_t> keys((size_t)n_states * KW);
        st->edge_src.resize((size_t)n_edges);
        st->edge_dst.resize((size_t)n_edges);
        st->edge_label.resize((size_t)n_edges * V);
        try {
            if (n_states) CK(cudaMemcpyAsync(keys.data(), state_key.p, keys.size() * 4, cudaMemcpyDeviceToHost, stream));
            if (n_edges) {
                CK(cudaMemcpyAsync(st->edge_src.data(), edge_src.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, stream));
                CK(cudaMemcpyAsync(st->edge_dst.data(), edge_dst.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, stream));
                CK(cudaMemcpyAsync(st->edge_label.data(), edge_label.p, (size_t)n_edges * V * 4, cudaMemcpyDeviceToHost, stream));
            }
            CK(cudaStreamSynchronize(stream));
        } catch (...) {
            delete st;
            throw;
        }
        d2h += (long long)keys.size() * 4 + n_edges * (2 + V) * 4;
        st->sig_vars = model->sets.sig_vars();
        st->state_sig.assign((size_t)n_states * SL, 0);
        st->state_cset.resize((size_t)n_states);
        st->state_failed.assign((size_t)n_states, 0);
        for (long long s = 0; s < n_states; s++) {
            st->state_cset[s] = std::max(0, keys[s * KW]);
            for (int j = 0; j < SL; j++) st->state_sig[s * SL + j] = keys[s * KW + 1 + j];
        }
        fill_header(part, ms);
        bind_store(part, st);
    }

    void fill_header(stcsp_automaton_t *part, float ms) {
        const int V = dm.V, SL = dm.sig_len;
        part->n_vars = V;
        part->n_sig_vars = dm.n_sig;
        part->n_until = model->sets.n_until();
        part->n_until_vars = model->sets.n_until_vars();
        part->sig_len = SL;
        part->root_final = model->sets.n_until() == 0;
        part->n_constraint_sets = model->sets.n_sets();
        if (part->n_states == 0) part->n_states = n_states;
        if (part->n_edges == 0) part->n_edges = n_edges;
        part->n_search_nodes = t_nodes;
        part->n_fails = t_fails;
        part->n_leaves = t_leaves;
        part->n_dominance = t_dominance;
        part->n_waves = t_waves;
        part->n_tuples = t_tuples;
        part->n_revisions = t_revisions;
        part->n_kernel_launches = t_launches;
        part->n_expand_launches = t_expand_launches;
        part->solve_ms = ms;
        part->expand_ms = expand_ms;
        part->wall_ms = (now_s() - t_create) * 1e3;
        part->h2d_bytes = h2d;
        part->d2h_bytes = d2h;
        // the reference's (lb, ub) int32 pairs per variable and offset
        part->algorithmic_bytes = t_nodes * 2ll * V * dm.k * 8 + t_leaves * ((SL + 1) * 4ll + 8) +
                                  n_states * ((SL + 1) * 4ll + 8) + n_edges * (V + 2) * 4ll;
    }

    // Single-rank finish: edges grouped by source and the fail rule applied ON THE DEVICE, then one copy into pinned
    // host memory.  State ids are already dense (world == 1: global id == local index, root == 0).
    // Device -> host.  Pinned destinations take one asynchronous copy; large pageable ones are filled through two pinned
    // staging chunks so that the PCIe transfer of one chunk overlaps the host memcpy of the previous one (pinning
    // gigabytes costs seconds, an unstaged pageable copy runs at ~2.5 GB/s).  Returns with the stream drained if staged.
    void download(const HostBlock &dst, const void *src, size_t bytes) {
        if (bytes == 0) return;
        if (dst.pinned()) {
            CK(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyDeviceToHost, stream));
            return;
        }
        const size_t CH = bytes > ((size_t)256 << 20) ? (size_t)32 << 20 : (size_t)4 << 20;
        HostBlock stage[2];
        stage[0].alloc(CH);
        stage[1].alloc(CH);
        if (!stage[0].pinned() || !stage[1].pinned()) {        // nothing can be pinned: let the driver stage it
            CK(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            return;
        }
        cudaEvent_t done[2] = {evk0, evk1};
        size_t off[2] = {0, 0}, len[2] = {0, 0};
        size_t pos = 0;
        for (int i = 0; pos < bytes || len[i & 1] || len[(i + 1) & 1]; i++) {
            const int slot = i & 1;
            if (len[slot]) {                            // the chunk issued two steps ago has landed: move it out
                CK(cudaEventSynchronize(done[slot]));
                parallel_copy((char *)dst.p + off[slot], stage[slot].p, len[slot]);
                len[slot] = 0;
            }
            if (pos < bytes) {
                const size_t n = std::min(CH, bytes - pos);
                CK(cudaMemcpyAsync(stage[slot].p, (const char *)src + pos, n, cudaMemcpyDeviceToHost, stream));
                CK(cudaEventRecord(done[slot], stream));
                off[slot] = pos;
                len[slot] = n;
                pos += n;
            }
        }
    }

    // Copy this rank's part (state keys, edges with global ids) into caller-provided device memory.
    void export_part(int32_t *keys, int32_t *src, int32_t *dst, int32_t *label) {
        const int KW = dm.key_words, V = dm.V;
        if (n_states) CK(cudaMemcpyAsync(keys, state_key.p, (size_t)n_states * KW * 4, cudaMemcpyDeviceToDevice, stream));
        if (n_edges) {
            CK(cudaMemcpyAsync(src, edge_src.p, (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(dst, edge_dst.p, (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, stream));
            CK(cudaMemcpyAsync(label, edge_label.p, (size_t)n_edges * V * 4, cudaMemcpyDeviceToDevice, stream));
        }
        CK(cudaStreamSynchronize(stream));
    }

    // Rank 0: all ranks' parts, concatenated in rank order in device memory, become the final automaton here.
    void finish_merged(int32_t W, const int64_t *ns_r, const int64_t *ne_r, const int32_t *keys, int32_t *src, int32_t *dst,
                       int32_t *label, const int64_t *extra, bool trim, stcsp_automaton_t *out) {
        if (W < 1 || W > kMaxWorld) throw Failure(STCSP_ERR_INVALID, "finish_merged: bad world size");
        long long ns = 0, ne = 0, counts[kMaxWorld] = {0};
        for (int r = 0; r < W; r++) { counts[r] = ns_r[r]; ns += ns_r[r]; ne += ne_r[r]; }
        if (counts[0] < 1) throw Failure(STCSP_ERR_INVALID, "finish_merged: rank 0 does not hold the root state");
        DBuf<int32_t> dense_keys;
        dense_keys.reserve((size_t)ns * dm.key_words + 1, 0, stream);
        launch_place_keys(keys, dense_keys.p, ns, dm.key_words, W, counts, sm_count, stream);
        launch_remap_ids(src, ne, W, counts, sm_count, stream);
        launch_remap_ids(dst, ne, W, counts, sm_count, stream);
        CK(cudaGetLastError());
        t_launches += 3;
        if (extra) {            // the other ranks' search statistics (same order as stcsp_session_counts)
            t_nodes += extra[0]; t_fails += extra[1]; t_leaves += extra[2]; t_dominance += extra[3]; t_tuples += extra[4];
            t_revisions += extra[5]; t_launches += extra[6]; t_expand_launches += extra[7]; h2d += extra[8]; d2h += extra[9];
        }
        const long long my_states = n_states, my_edges = n_edges;
        n_states = ns;          // fill_header reports the merged automaton
        n_edges = ne;
        try {
            finish_arrays(out, trim, dense_keys.p, ns, src, dst, label, ne);
        } catch (...) {
            n_states = my_states;
            n_edges = my_edges;
            throw;
        }
        n_states = my_states;
        n_edges = my_edges;
    }

    void finish_device(stcsp_automaton_t *out, bool trim) {
        finish_arrays(out, trim, state_key.p, n_states, edge_src.p, edge_dst.p, edge_label.p, n_edges);
    }

    // Group the edges (src, dst, label)[0 .. ne) by source, apply the fail rule, and copy states + edges to pinned host
    // memory.  State ids must be dense (0 .. ns); the edge arrays are scratch afterwards.
    void finish_arrays(stcsp_automaton_t *out, bool trim, const int32_t *keys_p, long long ns, int32_t *esrc_p, int32_t *edst_p,
                       int32_t *elabel_p, long long ne) {
        memset(out, 0, sizeof *out);
        const int KW = dm.key_words, V = dm.V, SL = dm.sig_len;
        DBuf<int32_t> &deg = fb_deg, &first = fb_first, &fill = fb_fill, &s_src = fb_src, &s_dst = fb_dst, &s_label = fb_label,
                      &outdeg = fb_outdeg, &flags = fb_flags, &rows_cset = fb_cset, &rows_sig = fb_sig;
        DBuf<uint8_t> &failed = fb_failed, &alive = fb_alive, &scan_tmp = fb_scan;
        struct Release {        // return the scratch blocks only after the stream drained
            cudaStream_t st;
            ~Release() { cudaStreamSynchronize(st); }
        };
        int32_t *f_src = esrc_p, *f_dst = edst_p, *f_label = elabel_p;
        long long n_final = ne;
        // small automata: one single-CTA launch does all of it and the host learns afterwards whether any edge died
        const bool small = ns <= 2048 && ne <= 512;    // (one CTA is slower than six launches beyond that)
        if (prefinished && keys_p == state_key.p) {
            // the search kernel did it (FinishArgs): the grouped edges, flags and state rows are in fb_*
            f_src = s_src.p; f_dst = s_dst.p; f_label = s_label.p;
        } else if (small) {
            deg.reserve((size_t)ns + 1, 0, stream);
            first.reserve((size_t)ns + 1, 0, stream);
            fill.reserve((size_t)ns + 1, 0, stream);
            outdeg.reserve((size_t)ns + 1, 0, stream);
            failed.reserve((size_t)ns + 1, 0, stream);
            alive.reserve((size_t)ne + 1, 0, stream);
            s_src.reserve((size_t)ne + 1, 0, stream);
            s_dst.reserve((size_t)ne + 1, 0, stream);
            s_label.reserve((size_t)ne * V + 1, 0, stream);
            rows_cset.reserve((size_t)ns + 1, 0, stream);
            rows_sig.reserve((size_t)ns * std::max(SL, 1) + 1, 0, stream);
            launch_finish_small(esrc_p, edst_p, elabel_p, (int)ne, V, (int)ns, deg.p, first.p, fill.p, outdeg.p,
                                failed.p, alive.p, s_src.p, s_dst.p, s_label.p, keys_p, KW, rows_cset.p, rows_sig.p,
                                trim ? 1 : 0, reinterpret_cast<int32_t *>(counters.p + C_OUT), stream);
            CK(cudaGetLastError());
            f_src = s_src.p; f_dst = s_dst.p; f_label = s_label.p;
            t_launches++;
        } else {
            deg.reserve((size_t)ns + 1, 0, stream);
            first.reserve((size_t)ns + 1, 0, stream);
            fill.reserve((size_t)ns + 1, 0, stream);
            failed.reserve((size_t)ns + 1, 0, stream);
            CK(cudaMemsetAsync(deg.p, 0, ((size_t)ns + 1) * 4, stream));
            CK(cudaMemsetAsync(fill.p, 0, ((size_t)ns + 1) * 4, stream));
            CK(cudaMemsetAsync(failed.p, 0, (size_t)ns + 1, stream));
            launch_edge_count(esrc_p, ne, deg.p, sm_count, stream);
            const size_t tmp = scan_temp_bytes(std::max<long long>(ns + 1, ne + 1));
            scan_tmp.reserve(tmp + 16, 0, stream);
            launch_exclusive_scan(scan_tmp.p, tmp, deg.p, first.p, ns + 1, stream);
            if (ne > 0) {
                s_src.reserve((size_t)ne, 0, stream);
                s_dst.reserve((size_t)ne, 0, stream);
                s_label.reserve((size_t)ne * V, 0, stream);
                launch_edge_scatter(esrc_p, edst_p, elabel_p, ne, V, first.p, fill.p, s_src.p, s_dst.p, s_label.p,
                                    sm_count, stream);
                f_src = s_src.p; f_dst = s_dst.p; f_label = s_label.p;
                t_launches += 3;
            }
            if (trim) {
                outdeg.reserve((size_t)ns + 1, 0, stream);
                alive.reserve((size_t)ne + 1, 0, stream);
                CK(cudaMemsetAsync(alive.p, 1, (size_t)ne + 1, stream));
                unsigned long long *ctl = counters.p + C_OUT;         // two scratch words: [0] changed, [1] dead edges
                CK(cudaMemsetAsync(ctl, 0, 16, stream));
                launch_trim_init(deg.p, ns, outdeg.p, failed.p, (int32_t *)ctl, sm_count, stream);
                t_launches++;
                for (;;) {
                    CK(cudaMemcpyAsync(h_counters, ctl, 16, cudaMemcpyDeviceToHost, stream));
                    CK(cudaStreamSynchronize(stream));
                    if ((int32_t)h_counters[0] == 0) break;
                    CK(cudaMemsetAsync(ctl, 0, 4, stream));
                    launch_trim_step(f_src, f_dst, ne, alive.p, outdeg.p, failed.p, (int32_t *)ctl, (int32_t *)(ctl + 1),
                                     sm_count, stream);
                    t_launches++;
                }
                const long long dead = (long long)(int32_t)h_counters[1];
                if (dead > 0) {
                    // compact the live edges back into the (now free) append-order buffers; order within a source is kept
                    flags.reserve((size_t)ne + 1, 0, stream);
                    launch_alive_to_int(alive.p, ne, flags.p, sm_count, stream);
                    const size_t tmp2 = scan_temp_bytes(ne + 1);
                    scan_tmp.reserve(tmp2 + 16, 0, stream);
                    launch_exclusive_scan(scan_tmp.p, tmp2, flags.p, flags.p, ne, stream);
                    launch_edge_compact(f_src, f_dst, f_label, alive.p, flags.p, ne, V, esrc_p, edst_p, elabel_p,
                                        sm_count, stream);
                    f_src = esrc_p; f_dst = edst_p; f_label = elabel_p;
                    n_final = ne - dead;
                    t_launches += 3;
                }
            }
            rows_cset.reserve((size_t)ns + 1, 0, stream);
            rows_sig.reserve((size_t)ns * std::max(SL, 1) + 1, 0, stream);
            launch_state_rows(keys_p, ns, KW, rows_cset.p, rows_sig.p, sm_count, stream);
            t_launches++;
            CK(cudaGetLastError());
        }
        // ---- post-processing fixpoints on the device, before anything is downloaded:
        // liveness (reference graphTraverse, src/graph.cpp:357-418) for models with `until` -- models without make every state
        // final, which the host knows without looking -- and the adversarial fixpoints -a / -z (src/graph.cpp:304-355, :247-302)
        // when the caller asked for them (stcsp_options_t::adversarial).  They need the live edges alone, so an automaton
        // finished by one CTA / inside the search kernel is compacted first if the fail rule killed edges.
        const bool pre = prefinished && keys_p == state_key.p;
        const bool liveness = model->sets.n_until() > 0 && ns > 0;
        const int adv = ns > 0 ? (opt.adversarial & 3) : 0;
        bool compacted_early = false;
        int adver1 = -1, adver2 = -1;
        if (adv && V < ((adv & 2) ? 7 : 6))
            throw Failure(STCSP_ERR_INVALID, "adversarial modes use variables #5 and #6 (reference src/graph.cpp:275,329); the model has too few");
        if (liveness || adv) {
            if (pre || small) {
                long long dead = prefinished_dead;
                if (!pre) {
                    CK(cudaMemcpyAsync(h_counters, counters.p + C_OUT, 8, cudaMemcpyDeviceToHost, stream));
                    CK(cudaStreamSynchronize(stream));
                    dead = (long long)(int32_t)h_counters[0];
                }
                if (dead > 0) {
                    flags.reserve((size_t)ne + 1, 0, stream);
                    launch_alive_to_int(alive.p, ne, flags.p, sm_count, stream);
                    const size_t tmp2 = scan_temp_bytes(ne + 1);
                    scan_tmp.reserve(tmp2 + 16, 0, stream);
                    launch_exclusive_scan(scan_tmp.p, tmp2, flags.p, flags.p, ne, stream);
                    launch_edge_compact(f_src, f_dst, f_label, alive.p, flags.p, ne, V, esrc_p, edst_p, elabel_p, sm_count, stream);
                    f_src = esrc_p; f_dst = edst_p; f_label = elabel_p;
                    n_final = ne - dead;
                    t_launches += 3;
                }
                compacted_early = true;
            }
            fb_fin.reserve((size_t)ns + 1, 0, stream);
            fb_valid.reserve((size_t)ns + 1, 0, stream);
            fb_palive.reserve((size_t)n_final + 1, 0, stream);
            unsigned long long *ctl = counters.p + C_OUT + 2;          // a scratch word the finishing kernels above do not use
            auto converge = [&](auto &&sweep) {
                for (;;) {
                    CK(cudaMemsetAsync(ctl, 0, 8, stream));
                    sweep();
                    CK(cudaMemcpyAsync(h_counters + 2, ctl, 8, cudaMemcpyDeviceToHost, stream));
                    CK(cudaStreamSynchronize(stream));
                    if ((int32_t)h_counters[2] == 0) break;
                }
            };
            auto root_valid = [&]() {
                uint8_t v = 0;
                CK(cudaMemcpyAsync(h_counters + 3, fb_valid.p, 1, cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                memcpy(&v, h_counters + 3, 1);
                return (int)v;
            };
            // final = all until flags set; without `until` every state is final and valid to begin with
            launch_liveness_init(keys_p, ns, KW, dm.n_sig, liveness ? model->sets.n_until_vars() : 0, liveness ? 0 : 1, fb_fin.p,
                                 fb_valid.p, sm_count, stream);
            t_launches++;
            if (liveness)
                converge([&] {
                    launch_liveness_step(f_src, f_dst, n_final, fb_valid.p, (int32_t *)ctl, sm_count, stream);
                    t_launches++;
                });
            const int32_t op_lb = adv ? model->sets.lb()[5] : 0, op_n = adv ? model->sets.width()[5] : 0;
            const unsigned long long full = op_n >= 64 ? ~0ull : ((1ull << op_n) - 1);
            if (adv & 1) {
                fb_cover.reserve((size_t)ns + 1, 0, stream);
                CK(cudaMemsetAsync(fb_cover.p, 0, ((size_t)ns + 1) * 8, stream));
                converge([&] {
                    launch_adv1_sweep(f_src, f_dst, f_label, n_final, V, 5, op_lb, full, ns, fb_cover.p, fb_valid.p, (int32_t *)ctl,
                                      sm_count, stream);
                    t_launches += 2;
                });
                adver1 = root_valid();
            }
            int ava_lb = 0, A = 0;
            if (adv & 2) {
                ava_lb = model->sets.lb()[6];
                A = model->sets.width()[6];
                fb_cover.reserve((size_t)ns * A + 1, 0, stream);
                converge([&] {
                    CK(cudaMemsetAsync(fb_cover.p, 0, (size_t)ns * A * 8, stream));
                    launch_adv2_sweep(f_src, f_dst, f_label, n_final, V, 5, op_lb, 6, ava_lb, A, full, ns, fb_cover.p, fb_valid.p,
                                      (int32_t *)ctl, sm_count, stream);
                    t_launches += 2;
                });
                adver2 = root_valid();
            }
            // the reference keeps the vertex's edges when -z leaves the root invalid (src/graph.cpp:288: only `if (root valid)`)
            const bool kill_by_cover = (adv & 2) && adver2 == 1;
            launch_post_alive(f_src, f_dst, f_label, n_final, V, 6, ava_lb, A, full, kill_by_cover ? fb_cover.p : nullptr, fb_valid.p,
                              fb_palive.p, sm_count, stream);
            t_launches++;
            CK(cudaGetLastError());
        }
        float ms = 0;
        if (timing_open) {
            CK(cudaEventRecord(ev1, stream));
        }
        // a small automaton that the search kernel pushed into the pinned buffers it was given is on the host already
        const bool use_push = pre && pushed && push_store != nullptr;
        PinnedStore *st = use_push ? push_store : new PinnedStore();
        if (use_push) push_store = nullptr;
        Release guard{stream};
        try {
            if (!use_push) {
                st->sig_vars.alloc(model->sets.sig_vars().size() * 4);
                st->state_sig.alloc((size_t)ns * SL * 4);
                st->state_cset.alloc((size_t)ns * 4);
                st->state_failed.alloc((size_t)ns);
                st->edge_src.alloc((size_t)n_final * 4);
                st->edge_dst.alloc((size_t)n_final * 4);
                st->edge_label.alloc((size_t)n_final * V * 4);
            }
            if (liveness || adv) {
                st->state_final.alloc((size_t)ns);
                st->state_valid.alloc((size_t)ns);
                st->edge_alive.alloc((size_t)n_final);
                download(st->state_final, fb_fin.p, (size_t)ns);
                download(st->state_valid, fb_valid.p, (size_t)ns);
                if (n_final) download(st->edge_alive, fb_palive.p, (size_t)n_final);
            }
            if (!model->sets.sig_vars().empty()) memcpy(st->sig_vars.p, model->sets.sig_vars().data(), model->sets.sig_vars().size() * 4);
            if (ns && !use_push) {
                if (SL) download(st->state_sig, rows_sig.p, (size_t)ns * SL * 4);
                download(st->state_cset, rows_cset.p, (size_t)ns * 4);
                download(st->state_failed, failed.p, (size_t)ns);
            }
            if (n_final && !use_push) {
                download(st->edge_src, f_src, (size_t)n_final * 4);
                download(st->edge_dst, f_dst, (size_t)n_final * 4);
                download(st->edge_label, f_label, (size_t)n_final * V * 4);
            }
            if (small && !pre && !compacted_early) CK(cudaMemcpyAsync(h_counters, counters.p + C_OUT, 8, cudaMemcpyDeviceToHost, stream));
            if (timing_open || !use_push || liveness || adv) CK(cudaStreamSynchronize(stream));
            if (timing_open) CK(cudaEventElapsedTime(&ms, ev0, ev1));
            const long long dead = compacted_early ? 0 : pre ? prefinished_dead : (small ? (long long)(int32_t)h_counters[0] : 0);
            if (dead > 0) {
                // rare (models with dead ends): compact the live edges into the append-order buffers and copy them again
                flags.reserve((size_t)ne + 1, 0, stream);
                launch_alive_to_int(alive.p, ne, flags.p, sm_count, stream);
                const size_t tmp2 = scan_temp_bytes(ne + 1);
                scan_tmp.reserve(tmp2 + 16, 0, stream);
                launch_exclusive_scan(scan_tmp.p, tmp2, flags.p, flags.p, ne, stream);
                launch_edge_compact(f_src, f_dst, f_label, alive.p, flags.p, ne, V, esrc_p, edst_p, elabel_p, sm_count,
                                    stream);
                CK(cudaGetLastError());
                n_final = ne - dead;
                t_launches += 3;
                if (n_final) {
                    download(st->edge_src, esrc_p, (size_t)n_final * 4);
                    download(st->edge_dst, edst_p, (size_t)n_final * 4);
                    download(st->edge_label, elabel_p, (size_t)n_final * V * 4);
                }
                CK(cudaStreamSynchronize(stream));
            }
        } catch (...) {
            delete st;
            throw;
        }
        d2h += ns * (SL + 1) * 4 + ns + n_final * (2 + V) * 4;
        out_states = ns;
        out_edges = n_final;
        out->n_states = ns;
        out->n_edges = n_final;
        fill_header(out, ms);
        out->n_edges = n_final;
        out->impl = static_cast<Store *>(st);
        out->sig_vars = (int32_t *)st->sig_vars.p;
        out->state_sig = (int32_t *)st->state_sig.p;
        out->state_cset = (int32_t *)st->state_cset.p;
        out->state_failed = (uint8_t *)st->state_failed.p;
        out->edge_src = (int32_t *)st->edge_src.p;
        out->edge_dst = (int32_t *)st->edge_dst.p;
        out->edge_label = (int32_t *)st->edge_label.p;
        if (liveness || adv) {
            out->state_final = (uint8_t *)st->state_final.p;
            out->state_valid = (uint8_t *)st->state_valid.p;
            out->edge_alive = (uint8_t *)st->edge_alive.p;
            out->post_applied = STCSP_POST_LIVENESS | (adv & 1 ? STCSP_POST_ADVERSARIAL1 : 0) | (adv & 2 ? STCSP_POST_ADVERSARIAL2 : 0);
            d2h += 2 * ns + n_final;
        }
        out->adver1 = adver1;
        out->adver2 = adver2;
    }
};

namespace stcsp {
namespace {

// ------------------------------------------------------------------------------------------------------------------
// A GROUP of ranks (one per GPU of one NVLink domain) that solve together: the persistent half of the multi-GPU path.
// It owns this rank's exchange block, the mappings of the peers' blocks and arenas, and the epoch counter; it lives
// across solves like a communicator does (mapping peer memory costs milliseconds, a solve of partialorder_18 fifty).
//
// Header row of an exchange (long long words):
enum HdrWord : int {
    H_N_IN = 0,          // search nodes this rank expanded in this wave
    H_N_LEAVES,          // leaves it routed
    H_N_PENDING,         // resolve requests it has (constraint-set transitions nobody has computed yet)
    H_STATUS,            // != 0: this rank failed (stcsp_status); everybody gives up
    H_OUTBOX_RAW, H_OUTBOX_ADDR,        // its outbox of this wave (records grouped by owner)
    H_PENDING_RAW, H_PENDING_ADDR,      // its request list
    H_COUNT0 = 8,        // + q: records for owner q
    H_OFF0 = 24,         // + q: where they start in the outbox (in records)
    H_N_STATES = 40, H_N_EDGES,
    H_KEYS_RAW, H_KEYS_ADDR, H_ESRC_RAW, H_ESRC_ADDR, H_EDST_RAW, H_EDST_ADDR, H_ELAB_RAW, H_ELAB_ADDR,
    H_STATS0 = 50,       // + 0..9: search statistics (same order as stcsp_session_counts)
    H_MODEL_SETS = 60, H_MODEL_CAPS     // constraint sets / transitions this rank's resident copy of the model knows at wave 0
};
static_assert(H_STATS0 + 10 <= H_MODEL_SETS && H_MODEL_CAPS < kHdrWords && H_OFF0 + kMaxWorld <= H_N_STATES, "header row layout");

// Meeting point of ranks that are threads of ONE process (stcsp_gpu_solve_multi): the same all-gather + barrier in host
// memory.  A waiting exchange KERNEL is no option there: with peer access enabled, a cudaMalloc or cudaHostRegister issued
// by one rank's thread maps memory into every peer context and waits for the peers' devices to drain -- which a kernel
// that is itself waiting for that rank never lets happen (measured: 30 s timeout on two B200s).
struct HostShared {
    std::atomic<long long> arrived{0}, generation{0};
    long long rows[2][kMaxWorld][kHdrWords];
};

struct ShareBlob {          // what a rank tells its peers once, when the group forms
    long long magic, pid, boot;
    int32_t rank, device;
    XBlock *block_raw;
    HostShared *shared_raw;
    cudaIpcMemHandle_t block_handle;
};

}  // namespace
}  // namespace stcsp

struct stcsp_group {
    int rank = 0, world = 1, device = 0;
    bool attached = false;
    XBlock *block = nullptr;
    XPeers peers{};
    bool same_process[kMaxWorld] = {false};
    unsigned long long epoch = 0;
    long long *d_row = nullptr;
    ArenaDir *d_dir = nullptr;
    int *d_status = nullptr;
    long long *h_rows = nullptr;            // pinned: [world][kHdrWords] + a status word + my row
    ArenaDir h_dir{};                       // what the peers know of my arenas
    ArenaDir peer_dir[kMaxWorld];
    std::map<long long, void *> peer_arena[kMaxWorld];      // arenas of rank q mapped into this process (CUDA IPC), by serial
    HostShared *my_shared = nullptr, *shared = nullptr;     // shared: rank 0's, when every rank lives in this process
    bool host_exchange = false;
    std::vector<std::string> wide_models;   // models known not to fit one GPU's wave limit: no single-GPU attempt
    cudaStream_t stream = nullptr;
    // statistics of the last sharded solve
    long long x_waves = 0, x_exchanges = 0, x_records = 0, x_bytes = 0;
    double x_exchange_ms = 0;

    ~stcsp_group() {
        if (device >= 0) cudaSetDevice(device);
        for (int q = 0; q < world; q++) {
            for (auto &kv : peer_arena[q])
                if (kv.second) cudaIpcCloseMemHandle(kv.second);
            if (attached && q != rank && !same_process[q] && peers.block[q]) cudaIpcCloseMemHandle(peers.block[q]);
        }
        if (stream) cudaStreamDestroy(stream);
        if (h_rows) cudaFreeHost(h_rows);
        if (d_status) cudaFree(d_status);
        if (d_dir) cudaFree(d_dir);
        if (d_row) cudaFree(d_row);
        if (block) cudaFree(block);
        delete my_shared;
    }

    static long long boot_id() {            // distinguishes hosts / containers that happen to share pids
        long long h = 1469598103934665603ll;
        FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r");
        if (f) {
            int c;
            while ((c = fgetc(f)) != EOF) h = (h ^ c) * 1099511628211ll;
            fclose(f);
        }
        return h;
    }

    void create(int r, int w, int dev) {
        if (w < 1 || w > kMaxWorld || r < 0 || r >= w) throw Failure(STCSP_ERR_INVALID, "bad rank / world size");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw Failure(STCSP_ERR_CUDA, "no usable CUDA device; this library has no CPU fallback");
        if (dev < 0) CK(cudaGetDevice(&dev));
        if (dev >= ndev) throw Failure(STCSP_ERR_CUDA, "CUDA device ordinal out of range");
        rank = r; world = w; device = dev;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&block, sizeof(XBlock)));
        CK(cudaMemset(block, 0, sizeof(XBlock)));
        CK(cudaMalloc(&d_row, kHdrWords * sizeof(long long)));
        CK(cudaMalloc(&d_dir, sizeof(ArenaDir)));
        CK(cudaMemset(d_dir, 0, sizeof(ArenaDir)));
        CK(cudaMalloc(&d_status, sizeof(int)));
        CK(cudaMemset(d_status, 0, sizeof(int)));
        CK(cudaMallocHost(&h_rows, ((size_t)kMaxWorld + 2) * kHdrWords * sizeof(long long)));
        my_shared = new HostShared();
        memset(my_shared->rows, 0, sizeof my_shared->rows);
        preload_search_kernels();
        preload_automaton_kernels(stream);
        preload_exchange_kernels();
        CK(cudaDeviceSynchronize());
        for (int q = 0; q < kMaxWorld; q++) peers.block[q] = nullptr;
        peers.block[rank] = block;
        same_process[rank] = true;
        memset(&h_dir, 0, sizeof h_dir);
        memset(peer_dir, 0, sizeof peer_dir);
    }

    void share(ShareBlob *b) {
        memset(b, 0, sizeof *b);
        b->magic = 0x53544353504752ll;
        b->pid = (long long)getpid();
        b->boot = boot_id();
        b->rank = rank;
        b->device = device;
        b->block_raw = block;
        b->shared_raw = my_shared;
        if (cudaIpcGetMemHandle(&b->block_handle, block) != cudaSuccess) cudaGetLastError();    // (single-process groups do not need it)
    }

    void attach(const ShareBlob *blobs) {
        CK(cudaSetDevice(device));
        ShareBlob mine;
        share(&mine);
        for (int q = 0; q < world; q++) {
            const ShareBlob &b = blobs[q];
            if (b.magic != mine.magic || b.rank != q) throw Failure(STCSP_ERR_INVALID, "group attach: blobs are not in rank order");
            if (q == rank) continue;
            same_process[q] = b.pid == mine.pid && b.boot == mine.boot;
            if (same_process[q]) {
                if (b.device != device) {
                    int can = 0;
                    CK(cudaDeviceCanAccessPeer(&can, device, b.device));
                    if (!can) throw Failure(STCSP_ERR_CUDA, "GPUs of the group cannot access each other's memory (no NVLink / P2P)");
                    cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                    cudaGetLastError();
                }
                peers.block[q] = b.block_raw;
            } else {
                void *p = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&p, b.block_handle, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    throw Failure(STCSP_ERR_CUDA, std::string("cannot map the exchange block of rank ") + std::to_string(q) +
                                                      " (CUDA IPC): " + cudaGetErrorString(e));
                }
                peers.block[q] = (XBlock *)p;
            }
        }
        // every rank a thread of this process: meet in host memory (STCSP_GROUP_EXCHANGE=device keeps the device-side
        // exchange, for tests that run all ranks on one GPU)
        bool all_local = true;
        for (int q = 0; q < world; q++) all_local = all_local && same_process[q];
        const char *force = getenv("STCSP_GROUP_EXCHANGE");
        host_exchange = all_local && !(force && !strcmp(force, "device"));
        shared = blobs[0].shared_raw;
        attached = true;
    }

    const long long *exchange_host(const long long *row) {
        const double t0 = now_s();
        epoch++;
        GTRACE("rank %d: host exchange %llu begins (n_in %lld leaves %lld pending %lld status %lld)", rank, epoch, row[H_N_IN],
               row[H_N_LEAVES], row[H_N_PENDING], row[H_STATUS]);
        memcpy(shared->rows[epoch & 1ull][rank], row, kHdrWords * sizeof(long long));
        const long long gen = shared->generation.load();
        if (shared->arrived.fetch_add(1) + 1 == world) {
            shared->arrived.store(0);
            shared->generation.store(gen + 1);
        } else {
            long spins = 0;
            while (shared->generation.load() == gen) {
                if (++spins > 2000) std::this_thread::yield();
                if ((spins & 0xfffff) == 0 && now_s() - t0 > 60.0)
                    throw Failure(STCSP_ERR_CUDA, "a rank of the group did not reach the exchange (timeout)");
            }
        }
        memcpy(h_rows, shared->rows[epoch & 1ull], (size_t)world * kHdrWords * sizeof(long long));
        x_exchanges++;
        x_exchange_ms += (now_s() - t0) * 1e3;
        for (int q = 0; q < world; q++)
            if (h_rows[(size_t)q * kHdrWords + H_STATUS] != 0)
                throw PeerFailure((int)h_rows[(size_t)q * kHdrWords + H_STATUS], "rank " + std::to_string(q) + " of the group failed");
        return h_rows;
    }

    // A device pointer of rank q as this process can use it.
    const void *peer_ptr(int q, long long raw, long long addr) {
        if (same_process[q]) return (const void *)(uintptr_t)raw;
        const long long serial = addr >> 40, off = addr & ((1ll << 40) - 1);
        if (addr < 0) throw Failure(STCSP_ERR_INVALID, "peer address outside the arenas of its rank");
        auto it = peer_arena[q].find(serial);
        if (it == peer_arena[q].end()) {
            // the directory rank q published with (or before) the row that carries this address
            CK(cudaMemcpyAsync(&peer_dir[q], &block->dir[q], sizeof(ArenaDir), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            long long idx = -1;
            for (long long i = 0; i < peer_dir[q].n && i < kMaxArenas; i++)
                if (peer_dir[q].serial[i] == serial) idx = i;
            if (idx < 0) throw Failure(STCSP_ERR_INVALID, "peer address names an arena its rank has not published");
            cudaIpcMemHandle_t h;
            memcpy(&h, peer_dir[q].handle[idx], 64);
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw Failure(STCSP_ERR_CUDA, std::string("cannot map device memory of rank ") + std::to_string(q) + " (CUDA IPC): " +
                                                  cudaGetErrorString(e));
            }
            it = peer_arena[q].emplace(serial, p).first;
        }
        return (const char *)it->second + off;
    }

    // my pointer p as (raw, arena << 40 | offset) for a header row
    void publish_ptr(const void *p, long long *row, int raw_word) {
        row[raw_word] = (long long)(uintptr_t)p;
        long long idx = 0, off = 0;
        if (p && device_cache().locate(device, p, idx, off)) row[raw_word + 1] = (idx << 40) | off;
        else row[raw_word + 1] = -1;
    }

    // All-gather of one header row per rank + barrier, on the devices.  rows: [world][kHdrWords] (pinned, valid until the
    // next exchange).  Throws if a peer reports failure or does not show up.
    const long long *exchange(const long long *row, cudaStream_t stream) {
        // (on the calling session's stream: a rank then needs ONE hardware queue to make progress; with every rank of a
        //  group on the same device -- the test layout -- more streams than CUDA_DEVICE_MAX_CONNECTIONS alias onto the same
        //  queue, and a kernel queued behind a peer's waiting exchange kernel would never start)
        if (host_exchange) return exchange_host(row);
        const double t0 = now_s();
        epoch++;
        GTRACE("rank %d: exchange %llu begins (n_in %lld leaves %lld pending %lld status %lld)", rank, epoch, row[H_N_IN], row[H_N_LEAVES],
               row[H_N_PENDING], row[H_STATUS]);
        long long *stage = h_rows + (size_t)(kMaxWorld + 1) * kHdrWords;
        memcpy(stage, row, kHdrWords * sizeof(long long));
        CK(cudaMemcpyAsync(d_row, stage, kHdrWords * sizeof(long long), cudaMemcpyHostToDevice, stream));
        ArenaDir now_dir;
        memset(&now_dir, 0, sizeof now_dir);
        device_cache().directory(device, now_dir);
        if (memcmp(&now_dir, &h_dir, sizeof now_dir) != 0) {
            h_dir = now_dir;
            CK(cudaMemcpy(d_dir, &h_dir, sizeof(ArenaDir), cudaMemcpyHostToDevice));
        }
        launch_exchange(peers, d_row, d_dir, rank, world, epoch, 30.0, d_status, stream);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_rows, &block->hdr[epoch & 1ull][0][0], (size_t)world * kHdrWords * sizeof(long long), cudaMemcpyDeviceToHost, stream));
        int *h_status = reinterpret_cast<int *>(h_rows + (size_t)kMaxWorld * kHdrWords);
        CK(cudaMemcpyAsync(h_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        GTRACE("rank %d: exchange %llu done after %.3f ms", rank, epoch, (now_s() - t0) * 1e3);
        x_exchanges++;
        x_exchange_ms += (now_s() - t0) * 1e3;
        if (*h_status != 0) throw Failure(STCSP_ERR_CUDA, "a rank of the group did not reach the exchange (timeout)");
        for (int q = 0; q < world; q++)
            if (h_rows[(size_t)q * kHdrWords + H_STATUS] != 0)
                throw PeerFailure((int)h_rows[(size_t)q * kHdrWords + H_STATUS], "rank " + std::to_string(q) + " of the group failed");
        return h_rows;
    }
};

namespace stcsp {
namespace {

// The sharded wave loop of one rank.  out: the merged automaton (rank 0 only).
void run_sharded(stcsp_session &s, stcsp_group &g, bool trim, stcsp_automaton_t *out) {
    const int W = g.world, R = g.rank, V = s.dm.V, RW = s.dm.rec_words, KW = s.dm.key_words;
    DBuf<int32_t> outbox[2], d_pending, m_keys, m_src, m_dst, m_label;
    g.x_waves = g.x_exchanges = g.x_records = g.x_bytes = 0;
    g.x_exchange_ms = 0;
    long long row[kHdrWords];
    const long long start_sets = s.model->sets.n_sets(), start_caps = s.model->capmap_used;
    bool local_failure = true;          // an exception raised here (not learnt from a peer) is announced to the peers
    try {
        for (long long wave = 0;; wave++) {
            if (s.n_states * (long long)W + R >= (1ll << 31) - (long long)W * 2)
                throw Failure(STCSP_ERR_CAPACITY, "more states than 32-bit global state ids can name at this world size");
            if (s.opt.time_limit_s > 0 && now_s() - s.t_create > s.opt.time_limit_s) throw Failure(STCSP_ERR_TIMEOUT, "time limit reached");
            memset(row, 0, sizeof row);
            row[H_N_IN] = s.n_in;
            int64_t nl = 0, np = 0;
            GTRACE("rank %d: wave %lld expand of %lld nodes", R, wave, s.n_in);
            s.expand(&nl, &np);
            GTRACE("rank %d: wave %lld expanded: %lld leaves, %lld requests", R, wave, (long long)nl, (long long)np);
            auto publish_outbox = [&]() {
                row[H_N_LEAVES] = s.n_leaves;
                row[H_N_PENDING] = (long long)(s.pending.size() / (size_t)(1 + V));
                for (int q = 0; q < W; q++) row[H_COUNT0 + q] = row[H_OFF0 + q] = 0;
                if (row[H_N_PENDING] == 0 && s.n_leaves > 0) {
                    DBuf<int32_t> &ob = outbox[wave & 1];
                    ob.reserve((size_t)s.n_leaves * RW, 0, s.stream);
                    int64_t counts[kMaxWorld] = {0};
                    s.outbox(ob.p, (int64_t)(ob.cap / RW), counts, false);
                    long long off = 0;
                    for (int q = 0; q < W; q++) {
                        row[H_COUNT0 + q] = counts[q];
                        row[H_OFF0 + q] = off;
                        off += counts[q];
                    }
                    g.publish_ptr(ob.p, row, H_OUTBOX_RAW);
                }
            };
            if (np > 0) {
                d_pending.reserve(s.pending.size(), 0, s.stream);
                CK(cudaMemcpyAsync(d_pending.p, s.pending.data(), s.pending.size() * 4, cudaMemcpyHostToDevice, s.stream));
                g.publish_ptr(d_pending.p, row, H_PENDING_RAW);
            }
            publish_outbox();
            if (wave == 0) {
                row[H_MODEL_SETS] = start_sets;
                row[H_MODEL_CAPS] = start_caps;
            }
            // the outbox is complete before anybody is told about it: the exchange kernel runs behind the scatter on this
            // stream (and fences before it raises its flags); only the meeting in host memory needs the host to wait
            if (g.host_exchange) CK(cudaStreamSynchronize(s.stream));
            const long long *rows = g.exchange(row, s.stream);
            if (wave == 0)          // set ids travel in the records: every rank must have started from the same numbering
                for (int q = 0; q < W; q++)
                    if (rows[(size_t)q * kHdrWords + H_MODEL_SETS] != start_sets || rows[(size_t)q * kHdrWords + H_MODEL_CAPS] != start_caps)
                        throw ModelMismatch();      // every rank sees the same rows: all of them start over without their copies
            long long total_in = 0, total_pending = 0;
            for (int q = 0; q < W; q++) {
                total_in += rows[(size_t)q * kHdrWords + H_N_IN];
                total_pending += rows[(size_t)q * kHdrWords + H_N_PENDING];
            }
            if (total_in == 0) break;                       // no rank had anything to expand: the search is over
            if (total_pending > 0) {
                // some rank met a constraint-set transition nobody has computed yet: every rank resolves the union of all
                // requests in one sorted order (so that set ids agree), then the outboxes are published again
                std::vector<std::vector<int32_t>> reqs;
                for (int q = 0; q < W; q++) {
                    const long long cnt = rows[(size_t)q * kHdrWords + H_N_PENDING];
                    if (cnt == 0) continue;
                    std::vector<int32_t> buf((size_t)cnt * (1 + V));
                    const void *src = g.peer_ptr(q, rows[(size_t)q * kHdrWords + H_PENDING_RAW], rows[(size_t)q * kHdrWords + H_PENDING_ADDR]);
                    CK(cudaMemcpyAsync(buf.data(), src, buf.size() * 4, cudaMemcpyDefault, s.stream));
                    CK(cudaStreamSynchronize(s.stream));
                    for (long long i = 0; i < cnt; i++) reqs.emplace_back(buf.begin() + i * (1 + V), buf.begin() + (i + 1) * (1 + V));
                }
                std::sort(reqs.begin(), reqs.end());
                reqs.erase(std::unique(reqs.begin(), reqs.end()), reqs.end());
                std::vector<int32_t> flat;
                for (const auto &r : reqs) flat.insert(flat.end(), r.begin(), r.end());
                s.resolve(flat.data(), (int64_t)reqs.size());
                memset(row, 0, sizeof row);
                row[H_N_IN] = 1;                            // (only the sum matters, and it was not zero)
                publish_outbox();
                if (g.host_exchange) CK(cudaStreamSynchronize(s.stream));
                rows = g.exchange(row, s.stream);
            }
            stcsp_session::PullSegs segs;
            segs.n = W;
            long long incoming = 0;
            for (int q = 0; q < W; q++) {
                const long long *rq = rows + (size_t)q * kHdrWords;
                segs.count[q] = rq[H_COUNT0 + R];
                segs.base[q] = nullptr;
                if (segs.count[q] > 0)
                    segs.base[q] = (const int32_t *)g.peer_ptr(q, rq[H_OUTBOX_RAW], rq[H_OUTBOX_ADDR]) + rq[H_OFF0 + R] * RW;
                incoming += segs.count[q];
                if (q != R) g.x_bytes += segs.count[q] * RW * 4;
            }
            g.x_records += incoming;
            int64_t next = 0;
            GTRACE("rank %d: wave %lld ingest of %lld records", R, wave, incoming);
            s.ingest(nullptr, 0, &next, &segs);
            GTRACE("rank %d: wave %lld ingested, next frontier %lld", R, wave, (long long)next);
            g.x_waves++;
        }
        // ---- merge: every rank publishes its part, rank 0 pulls them over NVLink and finishes on its GPU
        memset(row, 0, sizeof row);
        row[H_N_STATES] = s.n_states;
        row[H_N_EDGES] = s.n_edges;
        g.publish_ptr(s.state_key.p, row, H_KEYS_RAW);
        g.publish_ptr(s.edge_src.p, row, H_ESRC_RAW);
        g.publish_ptr(s.edge_dst.p, row, H_EDST_RAW);
        g.publish_ptr(s.edge_label.p, row, H_ELAB_RAW);
        {
            const long long v[10] = {s.t_nodes, s.t_fails, s.t_leaves, s.t_dominance, s.t_tuples, s.t_revisions, s.t_launches,
                                     s.t_expand_launches, s.h2d, s.d2h};
            for (int i = 0; i < 10; i++) row[H_STATS0 + i] = v[i];
        }
        CK(cudaStreamSynchronize(s.stream));
        const long long *rows = g.exchange(row, s.stream);
        if (R == 0) {
            int64_t ns[kMaxWorld], ne[kMaxWorld], extra[10] = {0};
            long long tot_s = 0, tot_e = 0;
            for (int q = 0; q < W; q++) {
                ns[q] = rows[(size_t)q * kHdrWords + H_N_STATES];
                ne[q] = rows[(size_t)q * kHdrWords + H_N_EDGES];
                tot_s += ns[q];
                tot_e += ne[q];
                if (q > 0)
                    for (int i = 0; i < 10; i++) extra[i] += rows[(size_t)q * kHdrWords + H_STATS0 + i];
            }
            if (tot_s >= (1ll << 31) || tot_e >= (1ll << 31)) throw Failure(STCSP_ERR_CAPACITY, "merged automaton exceeds 32-bit ids");
            m_keys.reserve((size_t)std::max<long long>(tot_s, 1) * KW, 0, s.stream);
            m_src.reserve((size_t)std::max<long long>(tot_e, 1), 0, s.stream);
            m_dst.reserve((size_t)std::max<long long>(tot_e, 1), 0, s.stream);
            m_label.reserve((size_t)std::max<long long>(tot_e, 1) * V, 0, s.stream);
            long long so = 0, eo = 0;
            // (copy the rows out first: the done-exchange below reuses the pinned table)
            std::vector<long long> keep(rows, rows + (size_t)W * kHdrWords);
            for (int q = 0; q < W; q++) {
                const long long *rq = keep.data() + (size_t)q * kHdrWords;
                if (ns[q])
                    CK(cudaMemcpyAsync(m_keys.p + so * KW, g.peer_ptr(q, rq[H_KEYS_RAW], rq[H_KEYS_ADDR]), (size_t)ns[q] * KW * 4,
                                       cudaMemcpyDefault, s.stream));
                if (ne[q]) {
                    CK(cudaMemcpyAsync(m_src.p + eo, g.peer_ptr(q, rq[H_ESRC_RAW], rq[H_ESRC_ADDR]), (size_t)ne[q] * 4, cudaMemcpyDefault, s.stream));
                    CK(cudaMemcpyAsync(m_dst.p + eo, g.peer_ptr(q, rq[H_EDST_RAW], rq[H_EDST_ADDR]), (size_t)ne[q] * 4, cudaMemcpyDefault, s.stream));
                    CK(cudaMemcpyAsync(m_label.p + eo * V, g.peer_ptr(q, rq[H_ELAB_RAW], rq[H_ELAB_ADDR]), (size_t)ne[q] * V * 4,
                                       cudaMemcpyDefault, s.stream));
                    if (q != 0) g.x_bytes += ne[q] * (2 + V) * 4 + ns[q] * KW * 4;
                }
                so += ns[q];
                eo += ne[q];
            }
            CK(cudaStreamSynchronize(s.stream));            // the parts are here: the peers may release theirs
            memset(row, 0, sizeof row);
            g.exchange(row, s.stream);
            s.search_complete = true;
            s.finish_merged(W, ns, ne, m_keys.p, m_src.p, m_dst.p, m_label.p, extra, trim, out);
        } else {
            memset(row, 0, sizeof row);
            g.exchange(row, s.stream);                      // rank 0 has copied this rank's part
            s.search_complete = true;
        }
        CK(cudaStreamSynchronize(s.stream));
    } catch (const PeerFailure &) {
        throw;                                              // everybody already knows
    } catch (const ModelMismatch &) {
        throw;                                              // everybody has seen it in the same exchange
    } catch (const Failure &f) {
        if (local_failure && f.code != STCSP_ERR_CUDA) {
            // tell the peers at the exchange they are (or will be) waiting at, so that nobody waits for the timeout
            try {
                memset(row, 0, sizeof row);
                row[H_STATUS] = f.code;
                g.exchange(row, s.stream);
            } catch (...) {
            }
        }
        throw;
    }
}

}  // namespace
}  // namespace stcsp

namespace {

int fail_with(const Failure &f) {
    set_error(f.what());
    return f.code;
}

template <class F>
int guarded(F &&body) {
    try {
        body();
        return STCSP_OK;
    } catch (const Failure &f) {
        return fail_with(f);
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return STCSP_ERR_CAPACITY;
    } catch (const std::exception &ex) {
        set_error(ex.what());
        return STCSP_ERR_INVALID;
    }
}

}  // namespace

extern "C" {

void stcsp_gpu_release_caches(void) {
    // resident models first (their blocks return to the block cache), then the blocks themselves
    {
        ModelCache &mc = model_cache();
        std::lock_guard<std::mutex> g(mc.mu);
        mc.entries.clear();
        mc.order.clear();
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (int d = 0; d < ndev; d++) device_cache().trim(d);
    host_cache().release_all();
}

int stcsp_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int stcsp_session_create(const stcsp_problem_t *problem, const stcsp_options_t *options, int32_t rank,
                         int32_t world_size, stcsp_session_t **out) {
    if (!problem || !out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    *out = nullptr;
    stcsp_session *s = nullptr;
    int rc = guarded([&] {
        s = new stcsp_session();
        s->init(problem, options, rank, world_size);
    });
    if (rc != STCSP_OK) {
        delete s;
        return rc;
    }
    *out = s;
    return STCSP_OK;
}

void stcsp_session_destroy(stcsp_session_t *s) { delete s; }

int32_t stcsp_session_record_words(const stcsp_session_t *s) { return s->dm.rec_words; }
int32_t stcsp_session_request_words(const stcsp_session_t *s) { return 1 + s->dm.V; }
int32_t stcsp_session_key_words(const stcsp_session_t *s) { return s->dm.key_words; }

int stcsp_session_expand(stcsp_session_t *s, int64_t *n_leaves, int64_t *n_pending) {
    return guarded([&] { s->expand(n_leaves, n_pending); });
}

int stcsp_session_pending(stcsp_session_t *s, int32_t *requests) {
    if (!s->pending.empty()) memcpy(requests, s->pending.data(), s->pending.size() * 4);
    return STCSP_OK;
}

int stcsp_session_resolve(stcsp_session_t *s, const int32_t *requests, int64_t n_requests) {
    return guarded([&] { s->resolve(requests, n_requests); });
}

int stcsp_session_outbox(stcsp_session_t *s, int32_t *outbox, int64_t outbox_capacity, int64_t *counts_per_rank) {
    return guarded([&] { s->outbox(outbox, outbox_capacity, counts_per_rank); });
}

int stcsp_session_ingest(stcsp_session_t *s, const int32_t *inbox, int64_t n_records, int64_t *frontier_next) {
    return guarded([&] { s->ingest(inbox, n_records, frontier_next); });
}

int stcsp_session_finish(stcsp_session_t *s, stcsp_automaton_t *part) {
    return guarded([&] { s->finish(part); });
}

int stcsp_session_counts(stcsp_session_t *s, int64_t *n_states, int64_t *n_edges, int64_t *stats) {
    *n_states = s->n_states;
    *n_edges = s->n_edges;
    if (stats) {
        const int64_t v[10] = {s->t_nodes, s->t_fails, s->t_leaves, s->t_dominance, s->t_tuples, s->t_revisions, s->t_launches,
                               s->t_expand_launches, s->h2d, s->d2h};
        memcpy(stats, v, sizeof v);
    }
    return STCSP_OK;
}

int stcsp_session_export(stcsp_session_t *s, int32_t *keys, int32_t *src, int32_t *dst, int32_t *label) {
    return guarded([&] { s->export_part(keys, src, dst, label); });
}

int stcsp_session_finish_merged(stcsp_session_t *s, int32_t world_size, const int64_t *n_states, const int64_t *n_edges,
                                const int32_t *keys, int32_t *src, int32_t *dst, int32_t *label, const int64_t *extra_stats,
                                int32_t trim, stcsp_automaton_t *out) {
    return guarded([&] { s->finish_merged(world_size, n_states, n_edges, keys, src, dst, label, extra_stats, trim != 0, out); });
}

int stcsp_automaton_assemble(const stcsp_automaton_t *parts, int32_t n_parts, stcsp_automaton_t *out) {
    if (!parts || n_parts < 1 || !out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    return guarded([&] {
        const stcsp_automaton_t &p0 = parts[0];
        const int V = p0.n_vars, SL = p0.sig_len, W = n_parts;
        int64_t total_states = 0, total_edges = 0, max_local = 0;
        for (int r = 0; r < W; r++) {
            if (parts[r].n_vars != V || parts[r].sig_len != SL) throw Failure(STCSP_ERR_INVALID, "assemble: parts disagree on the model");
            total_states += parts[r].n_states;
            total_edges += parts[r].n_edges;
            max_local = std::max<int64_t>(max_local, parts[r].n_states);
        }
        if (p0.n_states < 1) throw Failure(STCSP_ERR_INVALID, "assemble: part 0 does not hold the root state");
        // dense ids in ascending global-id order: global = local * W + rank
        std::vector<int32_t> dense((size_t)max_local * W, -1);
        int32_t next = 0;
        for (int64_t l = 0; l < max_local; l++)
            for (int r = 0; r < W; r++)
                if (l < parts[r].n_states) dense[(size_t)l * W + r] = next++;
        auto *st = new AutoStore();
        memset(out, 0, sizeof *out);
        st->sig_vars.assign(p0.sig_vars, p0.sig_vars + p0.n_sig_vars);
        st->state_sig.assign((size_t)total_states * SL, 0);
        st->state_cset.assign((size_t)total_states, 0);
        st->state_failed.assign((size_t)total_states, 0);
        for (int r = 0; r < W; r++)
            for (int64_t l = 0; l < parts[r].n_states; l++) {
                const int32_t d = dense[(size_t)l * W + r];
                st->state_cset[d] = parts[r].state_cset[l];
                if (parts[r].state_failed) st->state_failed[d] = parts[r].state_failed[l];
                for (int j = 0; j < SL; j++) st->state_sig[(size_t)d * SL + j] = parts[r].state_sig[l * SL + j];
            }
        // counting sort of the edges by dense source id
        std::vector<int64_t> first((size_t)total_states + 1, 0);
        auto map_id = [&](int32_t g) -> int32_t {
            if (g < 0 || (size_t)g >= dense.size() || dense[g] < 0) throw Failure(STCSP_ERR_INVALID, "assemble: edge names an unknown state");
            return dense[g];
        };
        for (int r = 0; r < W; r++)
            for (int64_t e = 0; e < parts[r].n_edges; e++) first[(size_t)map_id(parts[r].edge_src[e]) + 1]++;
        for (int64_t s = 0; s < total_states; s++) first[s + 1] += first[s];
        std::vector<int64_t> fillp(first.begin(), first.end() - 1);
        st->edge_src.resize((size_t)total_edges);
        st->edge_dst.resize((size_t)total_edges);
        st->edge_label.resize((size_t)total_edges * V);
        for (int r = 0; r < W; r++)
            for (int64_t e = 0; e < parts[r].n_edges; e++) {
                const int32_t s = map_id(parts[r].edge_src[e]);
                const int64_t at = fillp[s]++;
                st->edge_src[at] = s;
                st->edge_dst[at] = map_id(parts[r].edge_dst[e]);
                memcpy(&st->edge_label[(size_t)at * V], parts[r].edge_label + e * V, (size_t)V * 4);
            }
        out->n_vars = V;
        out->n_sig_vars = p0.n_sig_vars;
        out->n_until = p0.n_until;
        out->n_until_vars = p0.n_until_vars;
        out->sig_len = SL;
        out->root_final = p0.root_final;
        out->n_states = total_states;
        out->n_edges = total_edges;
        for (int r = 0; r < W; r++) {
            const stcsp_automaton_t &p = parts[r];
            out->n_constraint_sets = std::max(out->n_constraint_sets, p.n_constraint_sets);
            out->n_search_nodes += p.n_search_nodes;
            out->n_fails += p.n_fails;
            out->n_leaves += p.n_leaves;
            out->n_dominance += p.n_dominance;
            out->n_tuples += p.n_tuples;
            out->n_revisions += p.n_revisions;
            out->n_kernel_launches += p.n_kernel_launches;
            out->n_expand_launches += p.n_expand_launches;
            out->algorithmic_bytes += p.algorithmic_bytes;
            out->h2d_bytes += p.h2d_bytes;
            out->d2h_bytes += p.d2h_bytes;
            out->n_waves = std::max(out->n_waves, p.n_waves);
            out->solve_ms = std::max(out->solve_ms, p.solve_ms);
            out->wall_ms = std::max(out->wall_ms, p.wall_ms);
            out->expand_ms = std::max(out->expand_ms, p.expand_ms);
        }
        bind_store(out, st);
    });
}

int stcsp_automaton_trim(stcsp_automaton_t *a) {
    if (!a) { set_error("null argument"); return STCSP_ERR_INVALID; }
    return guarded([&] {
        const int64_t n = a->n_states, m = a->n_edges;
        const int V = a->n_vars;
        std::vector<int64_t> outdeg((size_t)n, 0), in_first((size_t)n + 1, 0);
        for (int64_t e = 0; e < m; e++) {
            outdeg[a->edge_src[e]]++;
            in_first[(size_t)a->edge_dst[e] + 1]++;
        }
        for (int64_t s = 0; s < n; s++) in_first[s + 1] += in_first[s];
        std::vector<int64_t> in_edge((size_t)m), fillp(in_first.begin(), in_first.end() - 1);
        for (int64_t e = 0; e < m; e++) in_edge[fillp[a->edge_dst[e]]++] = e;
        std::vector<uint8_t> alive((size_t)m, 1);
        std::vector<int32_t> todo;
        for (int64_t s = 0; s < n; s++)
            if (outdeg[s] == 0) { a->state_failed[s] = 1; todo.push_back((int32_t)s); }
        while (!todo.empty()) {
            const int32_t s = todo.back();
            todo.pop_back();
            for (int64_t i = in_first[s]; i < in_first[s + 1]; i++) {
                const int64_t e = in_edge[i];
                if (!alive[e]) continue;
                alive[e] = 0;
                const int32_t p = a->edge_src[e];
                if (--outdeg[p] == 0 && !a->state_failed[p]) { a->state_failed[p] = 1; todo.push_back(p); }
            }
        }
        int64_t w = 0;
        for (int64_t e = 0; e < m; e++) {       // in place: the arrays keep their capacity, n_edges shrinks
            if (!alive[e]) continue;
            if (w != e) {
                a->edge_src[w] = a->edge_src[e];
                a->edge_dst[w] = a->edge_dst[e];
                memmove(a->edge_label + (size_t)w * V, a->edge_label + (size_t)e * V, (size_t)V * 4);
            }
            w++;
        }
        a->n_edges = w;
    });
}

// ---- groups: several GPUs on one automaton --------------------------------------------------------------------------
int stcsp_group_create(int32_t rank, int32_t world_size, int32_t device, stcsp_group_t **out) {
    if (!out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    *out = nullptr;
    stcsp_group *g = nullptr;
    int rc = guarded([&] {
        g = new stcsp_group();
        g->create(rank, world_size, device);
    });
    if (rc != STCSP_OK) {
        delete g;
        return rc;
    }
    *out = g;
    return STCSP_OK;
}

void stcsp_group_destroy(stcsp_group_t *g) { delete g; }

int64_t stcsp_group_share_bytes(void) { return (int64_t)sizeof(ShareBlob); }

int stcsp_group_share(stcsp_group_t *g, void *blob) {
    if (!g || !blob) { set_error("null argument"); return STCSP_ERR_INVALID; }
    return guarded([&] { g->share((ShareBlob *)blob); });
}

int stcsp_group_attach(stcsp_group_t *g, const void *blobs) {
    if (!g || !blobs) { set_error("null argument"); return STCSP_ERR_INVALID; }
    return guarded([&] { g->attach((const ShareBlob *)blobs); });
}

int stcsp_group_solve(stcsp_group_t *g, const stcsp_problem_t *problem, const stcsp_options_t *options, stcsp_automaton_t *out,
                      stcsp_exchange_stats_t *xs) {
    if (!g || !problem || !out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    if (xs) memset(xs, 0, sizeof *xs);
    if (!g->attached && g->world > 1) { set_error("group solve before stcsp_group_attach"); return STCSP_ERR_INVALID; }
    stcsp_options_t opt;
    memset(&opt, 0, sizeof opt);
    if (options) opt = *options;
    opt.device = g->device;
    opt.use_current_device = 0;
    const double t0 = now_s();
    const std::string key = model_key(*problem, g->device);
    const bool known_wide = std::find(g->wide_models.begin(), g->wide_models.end(), key) != g->wide_models.end();
    if (g->world == 1 || (opt.shard_mode == 0 && !known_wide)) {
        // Instances whose waves fit one GPU are fastest on one GPU.  EVERY rank runs the same bounded single-GPU search on its
        // own device -- the verdict is a deterministic function of the model, so the ranks agree without exchanging a byte --
        // and rank 0 returns its result.  Only a wave wider than the bound makes the group shard the search.
        stcsp_options_t o1 = opt;
        const long long bound = 1ll << 20;
        if (g->world > 1 && (o1.max_frontier_nodes <= 0 || o1.max_frontier_nodes > bound)) o1.max_frontier_nodes = bound;
        stcsp_automaton_t tmp;
        const int rc = stcsp_gpu_solve(problem, &o1, &tmp);
        if (rc == STCSP_OK) {
            if (g->rank == 0) *out = tmp;
            else stcsp_automaton_free(&tmp);
            return STCSP_OK;
        }
        if (rc != STCSP_ERR_CAPACITY || g->world == 1) return rc;
        g->wide_models.push_back(key);
    }
    stcsp_session *s = nullptr;
    int rc = STCSP_OK;
    for (int attempt = 0; attempt < 2; attempt++) {
        bool mismatch = false;
        rc = guarded([&] {
            CK(cudaSetDevice(g->device));
            s = new stcsp_session();
            // (per rank: ranks that share a process -- stcsp_gpu_solve_multi -- must not swap their resident copies)
            s->group_tag = "|group" + std::to_string(g->world) + "r" + std::to_string(g->rank);
            GTRACE("rank %d: session init ...", g->rank);
            s->init(problem, &opt, g->rank, g->world);
            GTRACE("rank %d: ... session ready", g->rank);
            try {
                run_sharded(*s, *g, !opt.no_trim, out);
            } catch (const ModelMismatch &) {
                mismatch = true;            // the session dies without handing its copy back: the second attempt compiles afresh
                throw;
            }
        });
        delete s;
        s = nullptr;
        if (!mismatch) break;
    }
    if (xs) {
        xs->sharded = 1;
        xs->waves = g->x_waves;
        xs->exchanges = g->x_exchanges;
        xs->records = g->x_records;
        xs->bytes_pulled = g->x_bytes;
        xs->exchange_ms = g->x_exchange_ms;
    }
    if (rc == STCSP_OK && g->rank == 0) out->wall_ms = (now_s() - t0) * 1e3;
    if (rc != STCSP_OK) stcsp_automaton_free(out);
    return rc;
}

// One process, one host thread per GPU: the groups are formed once per device list and kept.
int stcsp_gpu_solve_multi(const stcsp_problem_t *problem, const stcsp_options_t *options, int32_t n_gpus, const int32_t *devices,
                          stcsp_automaton_t *out, stcsp_exchange_stats_t *xs) {
    if (!problem || !out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    if (n_gpus < 1 || n_gpus > kMaxWorld) { set_error("bad GPU count"); return STCSP_ERR_INVALID; }
    static std::mutex mu;
    static std::map<std::vector<int>, std::vector<stcsp_group *>> teams;
    std::lock_guard<std::mutex> lock(mu);                   // one multi-GPU solve at a time per process
    std::vector<int> devs;
    for (int i = 0; i < n_gpus; i++) devs.push_back(devices ? devices[i] : i);
    std::vector<int> team_key = devs;
    {
        const char *force = getenv("STCSP_GROUP_EXCHANGE");       // (part of the key: read when a team forms)
        team_key.push_back(force && !strcmp(force, "device") ? -2 : -1);
    }
    std::vector<stcsp_group *> &team = teams[team_key];
    if (team.empty()) {
        std::vector<ShareBlob> blobs((size_t)n_gpus);
        int rc = STCSP_OK;
        for (int r = 0; r < n_gpus && rc == STCSP_OK; r++) {
            stcsp_group *g = nullptr;
            rc = stcsp_group_create(r, n_gpus, devs[r], &g);
            if (rc == STCSP_OK) {
                team.push_back(g);
                rc = stcsp_group_share(g, &blobs[r]);
            }
        }
        for (int r = 0; r < n_gpus && rc == STCSP_OK; r++) rc = stcsp_group_attach(team[r], blobs.data());
        if (rc != STCSP_OK) {
            const std::string why = stcsp_last_error();
            for (stcsp_group *g : team) delete g;
            team.clear();
            set_error(why);
            return rc;
        }
    }
    std::vector<int> rcs((size_t)n_gpus, STCSP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<stcsp_automaton_t> outs((size_t)n_gpus);
    std::vector<stcsp_exchange_stats_t> xss((size_t)n_gpus);
    std::vector<std::thread> pool;
    for (int r = 1; r < n_gpus; r++)
        pool.emplace_back([&, r] {
            rcs[r] = stcsp_group_solve(team[r], problem, options, &outs[r], &xss[r]);
            if (rcs[r] != STCSP_OK) errs[r] = stcsp_last_error();
        });
    rcs[0] = stcsp_group_solve(team[0], problem, options, &outs[0], &xss[0]);
    if (rcs[0] != STCSP_OK) errs[0] = stcsp_last_error();
    for (std::thread &t : pool) t.join();
    for (int r = 1; r < n_gpus; r++) stcsp_automaton_free(&outs[r]);
    {
        // a rank that failed on its own explains more than the ranks that then waited for it in vain
        int first = -1;
        for (int r = 0; r < n_gpus; r++)
            if (rcs[r] != STCSP_OK && (first < 0 || errs[first].find("did not reach the exchange") != std::string::npos ||
                                       errs[first].find("of the group failed") != std::string::npos))
                first = first < 0 || errs[r].find("did not reach the exchange") == std::string::npos ? r : first;
        if (first >= 0) {
            std::string all = "rank " + std::to_string(first) + ": " + errs[first];
            for (int r = 0; r < n_gpus; r++)
                if (r != first && rcs[r] != STCSP_OK) all += " | rank " + std::to_string(r) + ": " + errs[r];
            stcsp_automaton_free(&outs[0]);
            set_error(all);
            return rcs[first];
        }
    }
    *out = outs[0];
    if (xs) *xs = xss[0];
    return STCSP_OK;
}

int stcsp_gpu_solve(const stcsp_problem_t *problem, const stcsp_options_t *options, stcsp_automaton_t *out) {
    if (!problem || !out) { set_error("null argument"); return STCSP_ERR_INVALID; }
    memset(out, 0, sizeof *out);
    const double t0 = now_s();
    stcsp_session *s = nullptr;
    double t_init = 0, t_loop = 0, t_finish = 0;
    const bool verbose = options && options->verbosity > 0;
    TraceRange whole("stcsp_gpu_solve");
    int rc = guarded([&] {
        s = new stcsp_session();
        s->root_in_kernel = !(options && options->profile_kernels);
        {
            TraceRange r("stcsp: init (model lookup / compile, upload, pools)");
            s->init(problem, options, 0, 1);
        }
        t_init = now_s();
        TraceRange search("stcsp: search (wave loop)");
        const double deadline = s->opt.time_limit_s > 0 ? t0 + s->opt.time_limit_s : 0;
        int64_t frontier = s->n_in;
        std::vector<int32_t> req;
        if (!s->opt.profile_kernels && s->search_grid > 0) {       // default: the wave loop runs on the device
            s->finish_in_kernel = true;
            s->finish_trim = !(options && options->no_trim);
            s->run_persistent(deadline);
            frontier = 0;
        }
        while (frontier > 0) {
            if (deadline > 0 && now_s() > deadline) throw Failure(STCSP_ERR_TIMEOUT, "time limit reached");
            if (s->opt.max_frontier_nodes > 0 && frontier > s->opt.max_frontier_nodes)
                throw Failure(STCSP_ERR_CAPACITY, "frontier wider than max_frontier_nodes");
            int64_t n_leaves = 0, n_pending = 0;
            s->expand(&n_leaves, &n_pending);
            if (n_pending > 0) {
                req = s->pending;
                s->resolve(req.data(), n_pending);
            }
            s->ingest(nullptr, 0, &frontier);
        }
        s->search_complete = true;
        t_loop = now_s();
        TraceRange fin("stcsp: finish (group, trim, post-process, download)");
        s->finish_device(out, !(options && options->no_trim));
        t_finish = now_s();
    });
    delete s;
    const double t_del = now_s();
    const double t_asm = now_s();
    if (verbose)
        fprintf(stderr, "[stcsp] wall: init %.2f ms, waves %.2f ms, download %.2f ms, release %.2f ms, assemble %.2f ms, trim %.2f ms\n",
                (t_init - t0) * 1e3, (t_loop - t_init) * 1e3, (t_finish - t_loop) * 1e3, (t_del - t_finish) * 1e3,
                (t_asm - t_del) * 1e3, (now_s() - t_asm) * 1e3);
    if (rc == STCSP_OK) out->wall_ms = (now_s() - t0) * 1e3;
    else stcsp_automaton_free(out);
    return rc;
}

int stcsp_gpu_warmup(int32_t device, int64_t pinned_bytes) {
    // everything a first call would otherwise pay for: context, kernel modules, streams, the first device and pinned
    // arenas (one solve of a two-state model walks through all of it), then the host arena for results
    stcsp_model_t *m = nullptr;
    int rc = stcsp_model_parse_text("var X : [0, 1];\nnext X == 1 - X;\n", 0, &m);
    if (rc != STCSP_OK) return rc;
    stcsp_options_t opt;
    memset(&opt, 0, sizeof opt);
    opt.device = device;
    stcsp_automaton_t a;
    rc = stcsp_gpu_solve(stcsp_model_problem(m), &opt, &a);
    stcsp_model_free(m);
    if (rc != STCSP_OK) return rc;
    stcsp_automaton_free(&a);
    return guarded([&] {
        preload_search_kernels();
        preload_automaton_kernels(nullptr);
        if (pinned_bytes > 0) host_cache().prepare_big((size_t)pinned_bytes);
    });
}

void stcsp_automaton_free(stcsp_automaton_t *a) {
    if (!a) return;
    delete (Store *)a->impl;
    memset(a, 0, sizeof *a);
}

}  // extern "C"
