// Device side of the search: data layout shared between the kernels and the host driver.
//
// SEARCH NODE (reference: the trail-protected currLB/currUB windows of all variables at one point
// of solverSolveRe, src/variable.h:19-20) -- a self-contained block of int32 words:
//     [0] source state id   [1] constraint-set id   [2] until-expired bits   [3] branched variable (-1: new state)
//     [4 ...] V*k uint64 domain bitsets, index (var*k + offset); bit b of variable v = value lb[v] + b
// LEAF RECORD (a complete consistent assignment of one time point, src/solveralgorithm.cpp:739-749):
//     [0] source state id   [1] constraint-set id   [2] until-expired bits   [3] 0   [4 ...] V values
//   after route_kernel: [1] SUCCESSOR constraint-set id   [2] successor until bits
//                       [3] hash of the successor's state key (owner rank and table slot derive from it)
// STATE IDS are global: local index * world + rank.
// STATE KEY (reference Signature, src/graph.h:14-21):
//     [0] constraint-set id (root: -1 unless the signature is empty)   [1 ...] signature values, until flags
#pragma once

#include <cstdint>

#include "compile.h"

namespace stcsp {

enum Counter : int {
    // allocation cursors, live for the whole search
    C_STATES = 0,      // states allocated on this rank
    C_EDGES,           // edges appended on this rank
    // per wave (zeroed by the host before each expand)
    C_OUT,             // nodes written to the output frontier by expand (frozen once expand is over)
    C_NEW,             // first nodes of new states appended behind them by ingest
    C_LEAVES,          // leaf records written
    C_UNRESOLVED,      // leaves whose successor constraint set the host must compute
    C_OVERFLOW,        // bit flags: 1 frontier (expand), 2 leaves, 4 states, 8 edges, 16 unresolved list, 32 frontier (ingest)
    C_NODES,           // statistics: search nodes propagated
    C_FAILS,
    C_TUPLES,
    C_REVISIONS,
    C_DOMINANCE,
    C_OWNER0,          // C_OWNER0 + r: routed leaves owned by rank r
    C_COUNT = C_OWNER0 + 16
};
// Two more running totals behind the counters proper, in set 0 (never cleared between waves): search nodes that ran their
// look-ahead propagators, and how many of those FAILED because of one (search_kernel's automatic look-ahead policy).
constexpr int C_AHEAD_NODES = 28, C_AHEAD_FAILS = 29;
static_assert(C_COUNT <= C_AHEAD_NODES, "the look-ahead totals live behind the counters");
constexpr int kCounterStride = 32;      // the session's counter block holds three sets (see search_kernel)
constexpr int kCounterSets = 3;
constexpr int kMaxWorld = 16;

struct DevModel {
    int32_t V, k;
    int32_t world, rank;        // one process per GPU; states are owned by hash (world == 1: everything local)
    int32_t node_words, rec_words, key_words;
    int32_t n_sig, sig_len;
    int32_t max_scope, max_stack, max_words;
    int32_t stage_bytes;        // shared memory reserved per CTA for one constraint set's metadata (0: never staged)
    int32_t node_slots;         // node blocks per CTA in shared memory: 8 (one per warp) or 32 (four per warp, quad mode)
    int32_t force_mode;         // 0 automatic, else ExpandMode + 1
    int32_t lazy_ahead;         // 1: pointwise propagators at look-ahead offsets run only once the current point is bound
    int32_t multi_branch;       // 1: narrow waves branch on up to three variables at once (branch_fan)
    int32_t fan_warps;          // children a narrow wave may create in total (2 x SM count), see branch_fan
    int32_t dbg_flags;          // experiments (environment STCSP_DBG_FLAGS); 0 in production
    int32_t n_sets;             // constraint sets of the model as of this launch (1: every node's set is set 0, no need to look)
    int32_t scalar_walk_cta;    // the same with a whole CTA on one node (the slowest thread is the round's duration)
    int32_t scalar_walk;        // longest relation-table walk (prefix tuples) ONE lane takes on in warp-per-node mode; longer
                                // walks go to the 32-lane revision
    long long enum_now, enum_ahead;
    const int32_t *lb, *width, *sig_vars;
    const DevSet *sets;
    const DevCon *cons;
    const DevProp *props;
    const int32_t *scope;
    const int32_t *stride;              // parallel to scope: relation-table stride of each slot
    const unsigned long long *tables;   // relation-table pool
    const Instr *code;
    const uint32_t *wake;
    const int32_t *aux;
    const int32_t *arr_off, *arr_val;
};

struct CapEntry {          // (constraint set, captured values) -> successor set; host-filled
    int32_t cid;           // -1 = empty slot
    int32_t next;
    int32_t off;           // captured values at capvals[off .. off + n_cap of the set)
    int32_t pad;
};

struct ExpandArgs {
    const int32_t *in_nodes;
    long long n_in;
    int32_t *out_nodes;
    long long out_cap;
    int32_t *leaves;
    long long leaf_cap;
    int fan;                    // most children one node may create this wave (branch_fan): > the first variable's domain
                                // means the node branches on further variables too
    unsigned long long *counters;
    unsigned long long *dbg;    // optional timeline of block 0 (CTA mode): dbg[0] = entries used, then (tag, %globaltimer) pairs
    int dbg_cap;
    // Look-ahead propagators (pointwise constraints at time offsets >= 1) only ever detect a dead end one state early: a
    // successor's first node is rebuilt from its signature, not from the offset-1 domains, and states without a way on are
    // removed by the fail rule anyway.  0: run them (or hold them until the point is bound, DevModel::lazy_ahead);
    // 1: skip them in every node of this wave.  Never changes the automaton.
    unsigned *block_stats;      // search_kernel: the block's shared-memory totals of this pass (see flush_warp_stats); else null
    int skip_ahead;
    unsigned long long *ahead_stats;    // set 0 of the counter block (C_AHEAD_NODES / C_AHEAD_FAILS), or null: do not count
    // search_kernel, narrow waves: a node that turns out to be a leaf is routed and merged into the automaton right away by
    // the warp that found it (no leaf phase, no second grid barrier); null in the stand-alone expand kernels
    const struct RouteArgs *fuse_route;
    const struct IngestArgs *fuse_ingest;
};

struct RouteArgs {
    int32_t *leaves;
    const int32_t *list;        // nullptr: leaves [0, C_LEAVES); else the leaves list[0 .. count)
    long long count;
    const CapEntry *capmap;
    const int32_t *capvals;
    int32_t capmap_mask;
    int32_t *unresolved;        // leaf indices that need the host
    long long unresolved_cap;
    unsigned long long *counters;
};

struct IngestArgs {
    const int32_t *records;     // routed leaf records
    long long count;
    int32_t *table;             // open addressing: local state index, -1 empty, -2 being written
    long long table_mask;
    int32_t *state_key;         // [local state * key_words]
    long long state_cap;
    int32_t *edge_src, *edge_dst, *edge_label;
    long long edge_cap;
    int32_t *out_nodes;
    long long out_base;             // nodes expand wrote (C_OUT after expand): new states' first nodes go behind them
    long long out_cap;
    unsigned long long *counters;   // the wave's counter set
    unsigned long long *totals;     // C_STATES, C_EDGES: never reset (set 0 of the session's counter block)
    // PULL mode (multi-GPU, n_segs > 0): `records` is null; record number i is the i-th record of the concatenation of the
    // segments, and segment q lies in the OUTBOX OF RANK q -- peer memory, read over NVLink by the ingesting warps themselves
    // (the exchange and the merge are one kernel; nothing is staged in an inbox).
    unsigned *block_stats;          // search_kernel's leaf phase: the block's shared-memory totals (flush_warp_stats), else null
    int32_t *deg;                   // optional: out-degree per (local = global, single rank) source state, counted as edges are
    long long deg_cap;              // appended (FinishArgs::deg_counted); states >= deg_cap are not counted
    int32_t n_segs;
    int32_t fused;                  // 1: called from inside expand (search_kernel, narrow waves): the first nodes of new states
                                    // are appended through expand's own cursor C_OUT (out_base = 0) instead of C_NEW
    const int32_t *seg_base[kMaxWorld];
    long long seg_count[kMaxWorld];
};

// Control block of the persistent search kernel (device memory, mirrored to the host when the kernel returns).
enum SearchStatus : int {
    SEARCH_RUN = 0,
    SEARCH_DONE = 1,       // frontier empty
    SEARCH_GROW = 2,       // a pool is too small for the next wave: the host grows it, nothing of the wave has run
    SEARCH_RETRY = 3,      // the output frontier overflowed during expand: grow it, run the wave again
    SEARCH_INGEST = 4,     // expand done, no room for new states' nodes: the host grows the frontier, routes and ingests
    SEARCH_YIELD = 5,      // wave budget used up (time-limit checks)
    SEARCH_RESOLVE = 6     // expand + route + ingest done except for leaves whose constraint-set transition is unseen
};
struct SearchCtl {
    long long n_in;
    int cur, status, overflow;
    int finished;           // 1: the kernel also grouped the edges by source and applied the fail rule (FinishArgs)
    long long waves_left;
    long long t_nodes, t_fails, t_tuples, t_revisions, t_dominance, t_leaves, t_waves;
    long long t_max_in;     // widest wave this launch ran
    int dead_edges, changed;
    int pushed, pad;        // 1: the finished automaton was also written into the host's pinned buffers (FinishArgs::h_*)
};
// Scratch for finishing a small automaton inside the search kernel (all null / 0: the host launches the finishing kernels).
struct FinishArgs {
    int32_t *deg, *first, *cursor, *outdeg;     // [cap_states + 1]
    uint8_t *failed, *alive;                    // [cap_states], [cap_edges]
    int32_t *s_src, *s_dst, *s_label;           // edges grouped by source
    int32_t *rows_cset, *rows_sig;              // state rows
    long long cap_states, cap_edges;
    int do_trim;
    // PUSH (optional, null = off): pinned host buffers, mapped into the device's address space.  A small automaton whose
    // fail rule killed no edge is written there by the kernel itself, so the host needs ONE synchronisation per solve and
    // no device-to-host copy at all (a copy of a few KB costs 5-8 us of stream time each, and there were nine of them).
    int32_t *h_cset, *h_sig, *h_src, *h_dst, *h_label;
    uint8_t *h_failed;
    long long h_cap_states, h_cap_edges;
    // 1: `deg` (zeroed, like `cursor`, by the first launch of the solve) has been counting the out-degree of every state
    // while the edges were appended (IngestArgs::deg) -- grouping then needs no counting pass and, when no state is a dead
    // end, no grid barrier at all (every block scans the degrees into its own shared memory)
    int32_t deg_counted;
};
struct SearchArgs {
    SearchCtl *ctl;
    unsigned long long *counters;
    long long n_in0, waves_left0;       // where this launch starts: frontier size, buffer index, wave budget (no control-block upload)
    int32_t cur0;
    int32_t make_root;                  // 1: first launch of a solve -- block 0 writes the root state and its search node itself
    SearchCtl *h_ctl;                   // optional: pinned host mirrors of ctl / counter set 0 (mapped); the kernel writes them
    unsigned long long *h_counters;     // when it leaves, so the host reads them after one stream synchronisation
    int32_t *frontier[2];
    long long out_cap;                  // capacity of EACH frontier buffer, in nodes
    int32_t *leaves;
    long long leaf_cap;
    int32_t *unresolved;
    long long unresolved_cap;
    const CapEntry *capmap;
    const int32_t *capvals;
    int32_t capmap_mask;
    int32_t *table;
    long long table_mask;
    int32_t *state_key;
    long long state_cap;
    int32_t *edge_src, *edge_dst, *edge_label;
    long long edge_cap;
    int32_t fuse_leaves;                // 1: narrow waves route and merge their leaves inside expand (see search_kernel)
    int32_t ahead_policy;               // look-ahead propagators: 0 always run, 1 automatic (ahead_droppable; sampled on), 2 never run
    long long leaves0, waves0;          // complete assignments found / waves run before this launch
    long long max_frontier;             // > 0: yield to the host when a wave is wider (it gives up, or runs the wave step-wise)
    FinishArgs fin;
    unsigned long long *trace;          // optional: 5 %globaltimer stamps per wave (start, expanded, routed, ingested, end)
    long long trace_cap;                // in waves
};

constexpr long long kAheadSample = 4096;   // nodes that must have run their look-ahead propagators in vain before the
                                           // automatic policy drops them (juggling_b6_f6_nosym, 3 913 nodes, never gets there:
                                           // its failures come in the last two waves, and they ARE look-ahead failures)
constexpr long long kAheadRatio = 64;
constexpr long long kAheadLeaves = 1024;   // ... and complete assignments seen before: the sample must cover search trees down to their
                                           // leaves (juggling_b7_f7_nosym fails nothing for five waves and a third of its nodes after)
__host__ __device__ inline bool ahead_droppable(long long ahead_nodes, long long ahead_fails, long long leaves) {
    return ahead_nodes >= kAheadSample && leaves >= kAheadLeaves && ahead_fails * kAheadRatio < ahead_nodes;
}
constexpr long long kAheadWavePeriod = 13; // one wave in thirteen keeps running them (odd: chains alternate branching and leaf waves)
__host__ __device__ inline bool ahead_sample_wave(long long wave) { return wave % kAheadWavePeriod == kAheadWavePeriod - 1; }
constexpr int kHostAheadWord = 256;        // where the two look-ahead totals are mirrored in the session's pinned counter block
constexpr int kExpandWarps = 8;      // warps per CTA of the expand kernel (measured: 4 warps x 6 CTAs per SM is a wash)
constexpr int kExpandCtasPerSm = 3;  // resident CTAs the launch bounds allow for (24 warps per SM, 80 registers per thread)

size_t expand_smem_bytes(const DevModel &m);
int expand_max_grid(const DevModel &m, int sm_count);      // resident CTAs of the expand kernel on this device
// mode: a warp per search node, a whole CTA per node (waves narrower than the GPU), or four nodes per warp (wide waves)
enum ExpandMode : int { EXPAND_WARP = 0, EXPAND_CTA = 1, EXPAND_QUAD = 2 };
void launch_expand(const DevModel &m, const ExpandArgs &a, int grid, int mode, cudaStream_t stream);
// the mode the automatic policy picks for a wave of n_in nodes on a grid of `ctas` resident CTAs
__host__ __device__ inline int pick_expand_mode(const DevModel &m, long long n_in, long long ctas) {
    const bool quad_ok = m.node_slots >= 4 * kExpandWarps && !m.lazy_ahead;
    if (m.force_mode == EXPAND_QUAD + 1) return quad_ok ? EXPAND_QUAD : EXPAND_WARP;
    if (m.force_mode) return m.force_mode - 1;
    if (n_in <= ctas) return EXPAND_CTA;          // every node gets a CTA of its own (measured: a second node per CTA loses to a warp per node)
    if (quad_ok && n_in >= 2ll * kExpandWarps * ctas) return EXPAND_QUAD;       // two nodes per resident warp or more
    return EXPAND_WARP;
}
// Children a node of a wave of n_in nodes may create: as long as the NEXT wave still has a resident warp per node
// (m.fan_warps, two per SM), branching on several variables at once saves whole waves and costs nothing but
// idle lanes.  1 = never more than the first unbound variable's domain.  Off in lazy look-ahead mode and when a mapping of
// nodes to threads is forced (tests compare the statistics of the paths).
__host__ __device__ inline int branch_fan(const DevModel &m, long long n_in, long long out_cap) {
    if (m.multi_branch == 0 || n_in <= 0) return 1;
    const long long warps = m.fan_warps;        // the same on every path, so that their search statistics agree
    const long long room = (warps < out_cap / 2 ? warps : out_cap / 2) / n_in;
    return room < 1 ? 1 : (room > 4096 ? 4096 : (int)room);
}
// persistent wave loop (cooperative launch); search_max_grid = co-resident CTAs, 0 if unavailable
int search_max_grid(const DevModel &m, int sm_count);
cudaError_t launch_search(const DevModel &m, const SearchArgs &a, int grid, int sm_count, cudaStream_t stream);
void launch_route(const DevModel &m, const RouteArgs &a, int grid, cudaStream_t stream);
void launch_ingest(const DevModel &m, const IngestArgs &a, int grid, cudaStream_t stream);
// group the routed local leaves by owner rank into `outbox` (dev_offsets = exclusive prefix of the owner counts)
void launch_scatter(const DevModel &m, const int32_t *leaves, long long n_leaves, const long long *dev_offsets,
                    unsigned long long *dev_fill, int32_t *outbox, int grid, cudaStream_t stream);
// dst[i] = src record list[i]
void launch_gather(const int32_t *src, const int32_t *list, long long count, int rec_words, int32_t *dst, int grid,
                   cudaStream_t stream);
// re-insert states [0, n_states) into a fresh (all -1) table
void launch_rehash(const DevModel &m, int32_t *table, long long table_mask, const int32_t *state_key,
                   long long n_states, int grid, cudaStream_t stream);
void launch_fill(int32_t *ptr, long long n, int32_t value, cudaStream_t stream);
// fill the relation tables of the constraints dev_jobs[0 .. n_jobs) (absolute indices) by evaluating their bytecode
void launch_build_tables(const DevModel &m, const int32_t *dev_jobs, int n_jobs, int max_entries, unsigned long long *tables,
                         cudaStream_t stream);

// automaton.cu: device-side finishing (group edges by source, fail-rule fixpoint, compaction)
size_t scan_temp_bytes(long long n);
void launch_exclusive_scan(void *temp, size_t temp_bytes, const int32_t *in, int32_t *out, long long n, cudaStream_t stream);
void launch_edge_count(const int32_t *src, long long n, int32_t *deg, int sm_count, cudaStream_t stream);
void launch_edge_scatter(const int32_t *src, const int32_t *dst, const int32_t *label, long long n, int V, const int32_t *first,
                         int32_t *fill, int32_t *osrc, int32_t *odst, int32_t *olabel, int sm_count, cudaStream_t stream);
void launch_trim_init(const int32_t *deg, long long n_states, int32_t *outdeg, uint8_t *failed, int32_t *changed, int sm_count,
                      cudaStream_t stream);
void launch_trim_step(const int32_t *src, const int32_t *dst, long long n, uint8_t *alive, int32_t *outdeg, uint8_t *failed,
                      int32_t *changed, int32_t *dead, int sm_count, cudaStream_t stream);
void launch_alive_to_int(const uint8_t *alive, long long n, int32_t *flag, int sm_count, cudaStream_t stream);
void launch_edge_compact(const int32_t *src, const int32_t *dst, const int32_t *label, const uint8_t *alive, const int32_t *pos,
                         long long n, int V, int32_t *osrc, int32_t *odst, int32_t *olabel, int sm_count, cudaStream_t stream);
// small automata: group by source + fail rule + state rows in one single-CTA launch; result[0] = dead edges
void launch_finish_small(const int32_t *src, const int32_t *dst, const int32_t *label, int n_edges, int V, int n_states,
                         int32_t *deg, int32_t *first, int32_t *cursor, int32_t *outdeg, uint8_t *failed, uint8_t *alive,
                         int32_t *osrc, int32_t *odst, int32_t *olabel, const int32_t *keys, int KW, int32_t *cset, int32_t *sig,
                         int do_trim, int32_t *result, cudaStream_t stream);
// multi-rank merge on one device: global ids -> dense ids; key rows (concatenated by rank) -> dense order
void launch_remap_ids(int32_t *ids, long long n, int world, const long long *n_states, int sm_count, cudaStream_t stream);
void launch_place_keys(const int32_t *keys_in, int32_t *keys_out, long long total_rows, int KW, int world, const long long *n_states,
                       int sm_count, cudaStream_t stream);
void launch_state_rows(const int32_t *keys, long long n_states, int KW, int32_t *cset, int32_t *sig, int sm_count,
                       cudaStream_t stream);

// post-processing fixpoints on the device (reference graphTraverse / adversarialTraverse / adversarialTraverse2): the host
// iterates a step / sweep until *changed stays 0
void launch_liveness_init(const int32_t *keys, long long ns, int KW, int n_sig, int n_flags, int root_final, uint8_t *fin,
                          uint8_t *valid, int sm_count, cudaStream_t stream);
void launch_liveness_step(const int32_t *src, const int32_t *dst, long long ne, uint8_t *valid, int32_t *changed, int sm_count,
                          cudaStream_t stream);
void launch_adv1_sweep(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int var, int lb,
                       unsigned long long full, long long ns, unsigned long long *have, uint8_t *valid, int32_t *changed,
                       int sm_count, cudaStream_t stream);
void launch_adv2_sweep(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int op, int op_lb, int ava,
                       int ava_lb, int A, unsigned long long full, long long ns, unsigned long long *cover, uint8_t *valid,
                       int32_t *changed, int sm_count, cudaStream_t stream);
void launch_post_alive(const int32_t *src, const int32_t *dst, const int32_t *label, long long ne, int V, int ava, int ava_lb, int A,
                       unsigned long long full, const unsigned long long *cover, const uint8_t *valid, uint8_t *alive, int sm_count,
                       cudaStream_t stream);

uint32_t capmap_hash(int cid, const int32_t *vals, int n);

// ---- exchange.cu: the per-wave meeting point of the ranks of a sharded solve, entirely on the devices ----------------
// Every rank owns one exchange block in device memory that all its peers have mapped (peer access in one process, CUDA
// IPC across processes).  At an exchange every rank writes its header row (and its arena directory) into EVERY peer's
// block, fences, raises its flag there, and waits until every peer's flag in its own block has reached the epoch: an
// all-gather of a few hundred bytes plus a barrier, with no host and no NCCL call in it.
constexpr int kHdrWords = 64;               // long long words per header row
constexpr int kMaxArenas = 48;              // device arenas a rank can publish
struct ArenaDir {                           // where a rank's device memory lives, for peers in other processes
    long long n;
    long long serial[kMaxArenas];           // process-wide unique id of the arena (indices shift when arenas are freed)
    long long size[kMaxArenas];
    unsigned char handle[kMaxArenas][64];   // cudaIpcMemHandle_t of each arena's base
};
struct XBlock {
    unsigned long long flag[kMaxWorld];             // flag[q]: the last epoch rank q has reached
    long long hdr[2][kMaxWorld][kHdrWords];         // header rows, double-buffered by epoch parity; row q is written by rank q
    ArenaDir dir[kMaxWorld];                        // dir[q] is written by rank q
};
struct XPeers {
    XBlock *block[kMaxWorld];                       // block[q]: rank q's exchange block as mapped into this process
};
// Load every kernel of the library now (see kernels.cu: lazy loading must not meet a waiting exchange kernel).
void preload_search_kernels();
void preload_automaton_kernels(cudaStream_t stream);
void preload_exchange_kernels();
// status (device int): 0 ok, 1 timed out waiting for a peer
void launch_exchange(const XPeers &peers, const long long *my_row, const ArenaDir *my_dir, int rank, int world,
                     unsigned long long epoch, double timeout_s, int *status, cudaStream_t stream);

}  // namespace stcsp
