/* stcsp_b200.h -- C ABI of the B200-native search-and-propagate path of the stcsp solver.
 *
 * The reference (AllenZzw/stcsp-solver) has no plugin or FFI layer; its hot path is entered
 * through exactly one call, `double solverSolve(Solver *solver, bool testing)`
 * (reference src/solveralgorithm.h:11, body src/solveralgorithm.cpp:945-1005, callers
 * src/solver.cpp:289 and :323).  This header is the drop-in boundary for that call: plain
 * pointers and sizes, no C++ or torch types.  A reference maintainer flattens `Solver` into
 * `stcsp_problem_t`, calls `stcsp_gpu_solve`, and rebuilds `Graph` from `stcsp_automaton_t`
 * (the binding is shown in INTEGRATION.md).
 *
 * Everything here runs on the GPU; there is no CPU fallback.  If no CUDA device is usable the
 * calls return STCSP_ERR_CUDA and `stcsp_last_error()` says why.
 */
#ifndef STCSP_B200_H
#define STCSP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STCSP_ABI_VERSION 2

/* ---------------------------------------------------------------------------------------------
 * Constraint expressions: one postfix (post-order) token list per constraint.
 *
 * Replaces the heap-allocated `ConstraintNode` trees (reference src/constraint.h:24-31).  A tree
 * is dumped left-to-right, children before parent; the arity of every operator is fixed, so the
 * list is the tree.  Token meaning follows the reference's evaluator `solverValidateRe`
 * (src/solveralgorithm.cpp:336-424).
 * ------------------------------------------------------------------------------------------- */
typedef struct stcsp_tok {
    int32_t op;   /* enum stcsp_op */
    int32_t arg;  /* CONST: value; VAR: variable index; ARR: array index; AT: time offset n */
} stcsp_tok_t;

enum stcsp_op {
    /* leaves */
    STCSP_OP_CONST = 1,   /* reference token CONSTANT   */
    STCSP_OP_VAR = 2,     /* reference token IDENTIFIER */
    /* unary: operand precedes */
    STCSP_OP_ARR = 3,     /* ARR_IDENTIFIER: T[operand]; out-of-range index poisons the evaluation */
    STCSP_OP_ABS = 4,
    STCSP_OP_NOT = 5,
    STCSP_OP_FIRST = 6,
    STCSP_OP_NEXT = 7,    /* only as the right side of `x == next y` */
    STCSP_OP_AT = 8,      /* only as the right side of `x == y@n`; arg = n */
    /* ternary: cond, then, else precede (reference IF/THEN node pair) */
    STCSP_OP_IF = 9,
    /* binary, expression level: left, right precede */
    STCSP_OP_LT = 10, STCSP_OP_GT = 11, STCSP_OP_LE = 12, STCSP_OP_GE = 13,
    STCSP_OP_EQ = 14, STCSP_OP_NE = 15,
    STCSP_OP_AND = 16, STCSP_OP_OR = 17,
    STCSP_OP_ADD = 18, STCSP_OP_SUB = 19, STCSP_OP_MUL = 20, STCSP_OP_DIV = 21, STCSP_OP_MOD = 22,
    /* binary, constraint level: always the last token of a constraint */
    STCSP_CON_LT = 32, STCSP_CON_GT = 33, STCSP_CON_LE = 34, STCSP_CON_GE = 35,
    STCSP_CON_EQ = 36, STCSP_CON_NE = 37, STCSP_CON_IMPLY = 38, STCSP_CON_UNTIL = 39
};

/* ---------------------------------------------------------------------------------------------
 * Problem: what `solverSolve` reads from `Solver` (reference src/solver.h:22-49).
 * Host-owned, read-only during the call.
 *   variables   <- solver->varQueue   (src/variable.h:10-26: name, lb, ub), declaration order with
 *                  the normaliser's auxiliaries `_V<n>` in creation order
 *   arrays      <- solver->arrayQueue (src/variable.h:54-59)
 *   constraints <- solver->constrQueue (src/constraint.h:38-49), normalised, in queue order.  Kind
 *                  (NEXT / POINT / UNTIL / AT), `hasFirst` and the signature variables are derived
 *                  from the token lists exactly as `solverConstraintQueuePush` does
 *                  (src/constraint.cpp:254-318); the caller does not pass them.
 * ------------------------------------------------------------------------------------------- */
typedef struct stcsp_problem {
    int32_t abi_version;             /* STCSP_ABI_VERSION */
    int32_t prefix_k;                /* solver->prefixK (flag -k, default 2) */
    int32_t n_vars;
    const int32_t *var_lb;           /* [n_vars] */
    const int32_t *var_ub;           /* [n_vars] */
    const char *const *var_names;    /* [n_vars] or NULL (only used for DOT headers) */
    int32_t n_arrays;
    const int32_t *arr_offsets;      /* [n_arrays + 1] into arr_values */
    const int32_t *arr_values;
    int32_t n_constraints;
    const int32_t *con_offsets;      /* [n_constraints + 1] into con_tokens */
    const stcsp_tok_t *con_tokens;
} stcsp_problem_t;

/* Options.  Zero-initialise, then set what you need; 0 always means "default". */
typedef struct stcsp_options {
    int32_t device;                  /* CUDA device ordinal; -1 or 0-with-use_current = current device */
    int32_t use_current_device;      /* 1: do not call cudaSetDevice (the host framework already did) */
    int32_t time_limit_s;            /* flag -m; 0 = none.  On expiry: STCSP_ERR_TIMEOUT */
    int32_t verbosity;               /* flag -l */
    int64_t enum_limit_now;          /* max tuples enumerated per constraint revision at the current time point */
    int64_t enum_limit_ahead;        /* same, at look-ahead time points 1..k-1 */
    int64_t max_frontier_nodes;      /* > 0: give up with STCSP_ERR_CAPACITY as soon as a wave is wider than this (a multi-GPU
                                        driver uses it to keep small instances on one GPU); 0: no limit */
    int64_t max_states;              /* reserved */
    int64_t max_edges;               /* reserved */
    int32_t expand_mode;             /* mapping of search nodes to threads: 0 automatic (a CTA per node on narrow waves, four nodes
                                        per warp on wide ones, a warp per node in between), 1 warp per node always, 2 CTA per
                                        node always, 3 four nodes per warp always (testing / tuning; never changes results) */
    int32_t profile_kernels;         /* 1: step-wise path (one expand / route / ingest launch per wave) with every expand launch
                                        timed by CUDA events, instead of the persistent search kernel */
    int32_t no_trim;                 /* 1: stcsp_gpu_solve returns the untrimmed automaton */
    int32_t lookahead;               /* pointwise constraints at time offsets >= 1: 0 automatic (eager; dropped while 4096 or
                                        more search nodes have run them, fewer than one in 64 of those failed because of one
                                        and 1024 complete assignments have been seen -- but for one wave in 13, which keeps
                                        measuring), 1 eager (propagated like the current
                                        point, the reference's prefix-k consistency), 2 lazy (checked once, when every variable
                                        of the current point is bound), 3 never (a dead end is then found one state later and
                                        removed by the fail rule).  Never changes the automaton. */
    int32_t wide_wave_nodes;         /* waves wider than this run as separate full-occupancy launches instead of inside the
                                        persistent search kernel: 0 = default (32768), < 0 = never */
    int32_t single_branch;           /* 1: a search node branches on ONE variable (the first unbound one), like the reference's
                                        solverGetFirstUnboundVar; 0 (default): narrow waves branch on up to three variables at
                                        once, which shortens the search tree.  Never changes the automaton. */
    int32_t shard_mode;              /* group solves: 0 adaptive (instances whose waves fit one GPU stay on one GPU), 1 always shard
                                        the search over the GPUs of the group */
    int32_t adversarial;             /* post-processing fixpoints to run ON THE DEVICE before the download: bit 0 = flag -a
                                        (reference adversarialTraverse, src/graph.cpp:304-355), bit 1 = flag -z
                                        (adversarialTraverse2, :247-302).  The automaton then carries state_valid / edge_alive /
                                        adver1 / adver2 and stcsp_postprocess, called with the same flags, only numbers it. */
} stcsp_options_t;

/* ---------------------------------------------------------------------------------------------
 * Automaton: what `solverSolve` leaves in `solver->graph` (reference src/graph.h:47-72) before
 * post-processing, i.e. after the fail rule (src/solveralgorithm.cpp:904-910) and before
 * `graphTraverse`.  Library-owned until `stcsp_automaton_free`.
 *   state 0 is the root (signature "S", constraint set 0).
 *   sig[s]  = values of the signature variables in variable order, then one 0/1 flag per UNTIL
 *             constraint (src/solveralgorithm.cpp:810-837); undefined for the root.
 *   failed  = state has no surviving out-edge (src/solveralgorithm.cpp:865-868, 907-909).
 *   edges   = (src, dst, the full assignment of the source time point), dst never failed
 *             (src/graph.cpp:78-89).  Grouped by src (ascending); order within one source is unspecified.
 *   constraint-set ids are dense but their numbering is this library's (the reference numbers
 *   them in DFS discovery order); compare automata after canonical relabelling.
 * ------------------------------------------------------------------------------------------- */
typedef struct stcsp_automaton {
    int32_t n_vars;
    int32_t n_sig_vars;              /* solver->numSignVar */
    int32_t n_until;                 /* number of UNTIL constraints (flags per state) */
    int32_t n_until_vars;            /* solver->numUntil (distinct right-hand variables) */
    int32_t sig_len;                 /* n_sig_vars + n_until */
    int32_t *sig_vars;               /* [n_sig_vars] variable index of each signature column */
    int32_t root_final;              /* 1 iff the initial constraint set has no UNTIL constraint */
    int32_t n_constraint_sets;
    int64_t n_states;
    int32_t *state_sig;              /* [n_states * sig_len] */
    int32_t *state_cset;             /* [n_states] constraint-set id */
    uint8_t *state_failed;           /* [n_states] */
    int64_t n_edges;
    int32_t *edge_src;               /* [n_edges] */
    int32_t *edge_dst;               /* [n_edges] */
    int32_t *edge_label;             /* [n_edges * n_vars] */
    /* statistics (search-order dependent, not parity targets) */
    int64_t n_search_nodes;          /* propagate-to-fixpoint calls (reference: generalisedArcConsistent calls) */
    int64_t n_fails;                 /* search nodes that wiped out a domain */
    int64_t n_leaves;                /* complete consistent assignments found */
    int64_t n_dominance;             /* leaves that hit an existing state */
    int64_t n_waves;                 /* frontier iterations */
    int64_t n_tuples;                /* constraint evaluations */
    int64_t n_revisions;             /* propagator executions */
    int64_t n_kernel_launches;
    double solve_ms;                 /* device time, CUDA events around the whole search */
    double wall_ms;                  /* host wall clock of the call incl. uploads/downloads */
    double expand_ms;                /* device time inside the dominant kernel, CUDA events around every launch: the persistent
                                        search kernel (default path) or the expand kernel (profile_kernels = 1, step-wise path) */
    int64_t n_expand_launches;       /* launches of that kernel */
    int64_t algorithmic_bytes;       /* SURVEY.md section 8(d) formula with this run's counts */
    int64_t h2d_bytes, d2h_bytes;
    /* Post-processing done on the device.  Liveness of reference graphTraverse
     * (src/graph.cpp:357-418) is computed for every model with `until`: final = all until flags set, valid = a final state can
     * be reached; the adversarial fixpoints when stcsp_options_t::adversarial asks for them.  post_applied says which ran
     * (0: none -- models without `until`, automata assembled on the host -- the three pointers are NULL and
     * stcsp_postprocess derives everything itself). */
    uint8_t *state_final;            /* [n_states] or NULL */
    uint8_t *state_valid;            /* [n_states] or NULL: valid after every fixpoint that ran */
    uint8_t *edge_alive;             /* [n_edges] or NULL: 0 = removed by the fixpoints that ran (an end is invalid, or -z dropped it) */
    int32_t post_applied;            /* STCSP_POST_* bits */
    int32_t adver1, adver2;          /* root valid after -a / after -z (what the reference prints, src/solver.cpp:300-321); -1 = not run */
    int32_t pad0;
    void *impl;                      /* private */
} stcsp_automaton_t;

enum stcsp_post { STCSP_POST_LIVENESS = 1, STCSP_POST_ADVERSARIAL1 = 2, STCSP_POST_ADVERSARIAL2 = 4 };

enum stcsp_status {
    STCSP_OK = 0,
    STCSP_ERR_INVALID = 1,           /* malformed problem (bad token list, bad index, abi mismatch) */
    STCSP_ERR_UNSUPPORTED = 2,       /* domain wider than 64 values, too many variables ... */
    STCSP_ERR_CUDA = 3,              /* no device / CUDA failure: there is NO CPU fallback */
    STCSP_ERR_CAPACITY = 4,          /* frontier / state table / edge store full: raise the option */
    STCSP_ERR_TIMEOUT = 5,
    STCSP_ERR_PARSE = 6              /* front end: syntax error or undefined name */
};

/* Replaces solverSolve() (reference src/solveralgorithm.cpp:945-1005) up to, not including,
 * graphTraverse.  `out` is filled on STCSP_OK and must be released with stcsp_automaton_free. */
int stcsp_gpu_solve(const stcsp_problem_t *problem, const stcsp_options_t *options, stcsp_automaton_t *out);

/* Replaces graphFree() (reference src/graph.cpp:156-163). */
void stcsp_automaton_free(stcsp_automaton_t *a);

/* Message of the last failing call on this thread (the reference logs to error.txt and exits,
 * src/util.cpp:161-178; a library must not exit). */
const char *stcsp_last_error(void);

/* The library keeps device blocks, pinned host blocks, streams and the compiled form of recently solved models alive
 * between calls (a solve in steady state allocates nothing).  This gives the memory back (idle sessions only). */
void stcsp_gpu_release_caches(void);

/* Number of usable CUDA devices (0 if none); never fails. */
int stcsp_gpu_device_count(void);

/* Optional one-time set-up for a process that will solve more than once (a service, a benchmark loop; the command-line
 * tool, like the reference, solves once and does not call it): creates the context on `device` (as
 * stcsp_options_t::device with use_current_device = 0: a negative ordinal means device 0), loads
 * the kernel modules, reserves the first device arena and pins `pinned_bytes` of host memory for results (0 = none), so
 * that the first solve of a model costs what a new MODEL costs (compile, upload, relation tables), not what a new process
 * costs.  No reference counterpart (the reference is a one-shot program, src/solver.cpp:181-345). */
int stcsp_gpu_warmup(int32_t device, int64_t pinned_bytes);

/* ---------------------------------------------------------------------------------------------
 * Step-wise session API: the same search, one frontier wave at a time, so that a multi-GPU
 * driver (one process per GPU) can exchange leaf records by hash owner between `expand` and
 * `ingest`.  stcsp_gpu_solve() is exactly create / loop(expand, [resolve], ingest) / finish /
 * assemble / trim on one rank.
 *
 * One wave on every rank:
 *   expand   propagate + branch every search node of the local frontier; route every leaf
 *            (successor constraint set, until flags, hash of the successor's state key)
 *   pending / resolve   only when a leaf's successor constraint set is not known yet (models whose
 *            `first` captures values, reference src/constraint.cpp:466-548): the driver collects the
 *            requests of ALL ranks, sorts them, and hands the same list to every rank, so that
 *            constraint-set ids are identical everywhere
 *   outbox   (world_size > 1) leaves grouped by owner rank into caller-provided DEVICE memory; the
 *            caller's collective (NCCL all-to-all through torch.distributed) moves them
 *   ingest   dedup the received leaves against the local state table, append edges, create the
 *            first search node of every new state; ends the wave
 *
 * A leaf record is `rec_words` int32: [src_state_global, successor_cset, successor_until_bits,
 * state_key_hash, label[n_vars]].
 * ------------------------------------------------------------------------------------------- */
typedef struct stcsp_session stcsp_session_t;

int stcsp_session_create(const stcsp_problem_t *problem, const stcsp_options_t *options,
                         int32_t rank, int32_t world_size, stcsp_session_t **out);
void stcsp_session_destroy(stcsp_session_t *s);
/* int32 words per leaf record / per resolve request (1 + n_vars: constraint set, assignment) */
int32_t stcsp_session_record_words(const stcsp_session_t *s);
int32_t stcsp_session_request_words(const stcsp_session_t *s);
/* int32 words per state key (1 + sig_len: constraint set, signature values, until flags) */
int32_t stcsp_session_key_words(const stcsp_session_t *s);
/* Expand the local frontier by one wave.  *n_leaves: leaves found; *n_pending: distinct resolve
 * requests this rank has (0 almost always). */
int stcsp_session_expand(stcsp_session_t *s, int64_t *n_leaves, int64_t *n_pending);
/* Copy the pending requests to host memory [n_pending * request_words]. */
int stcsp_session_pending(stcsp_session_t *s, int32_t *requests);
/* Resolve requests (the union over all ranks, same order on every rank; duplicates allowed). */
int stcsp_session_resolve(stcsp_session_t *s, const int32_t *requests, int64_t n_requests);
/* Group this wave's leaves by owner rank into `outbox` (device, capacity in records >= n_leaves);
 * counts_per_rank[world_size] (host) receives the record count for each rank; rank q's records
 * start at record index sum(counts_per_rank[0..q)). */
int stcsp_session_outbox(stcsp_session_t *s, int32_t *outbox, int64_t outbox_capacity, int64_t *counts_per_rank);
/* Insert `n_records` routed leaf records (device memory) owned by this rank and end the wave.
 * inbox == NULL: world_size == 1 ingests this rank's own leaves in place; otherwise nothing was received.  *frontier_next = local search
 * nodes waiting for the next wave. */
int stcsp_session_ingest(stcsp_session_t *s, const int32_t *inbox, int64_t n_records, int64_t *frontier_next);
/* Local part of the automaton: the states this rank owns (rows in local-index order, global id =
 * local_index * world_size + rank) and the edges INTO them with global src/dst ids, untrimmed and
 * unsorted.  Release with stcsp_automaton_free. */
int stcsp_session_finish(stcsp_session_t *s, stcsp_automaton_t *part);

/* Device-side merge (preferred over finish + assemble for large automata):
 *   counts        this rank's state / edge counts and its 10 search statistics
 *   export        copy this rank's part into caller-provided DEVICE memory: keys [n_states * (1 + sig_len)], src / dst
 *                 [n_edges] (global ids), label [n_edges * n_vars]; the caller moves them to rank 0 (NCCL send / recv)
 *   finish_merged rank 0 only: the parts of all ranks, concatenated in rank order in device memory (keys, src, dst, label;
 *                 src / dst / label are overwritten), are renumbered to dense ids, grouped by source, trimmed and copied to
 *                 the host.  extra_stats = element-wise sum of the other ranks' statistics (or NULL). */
int stcsp_session_counts(stcsp_session_t *s, int64_t *n_states, int64_t *n_edges, int64_t *stats10);
int stcsp_session_export(stcsp_session_t *s, int32_t *keys, int32_t *src, int32_t *dst, int32_t *label);
int stcsp_session_finish_merged(stcsp_session_t *s, int32_t world_size, const int64_t *n_states, const int64_t *n_edges,
                                const int32_t *keys, int32_t *src, int32_t *dst, int32_t *label, const int64_t *extra_stats,
                                int32_t trim, stcsp_automaton_t *out);

/* ---------------------------------------------------------------------------------------------
 * Groups: several GPUs of one NVLink domain on ONE automaton.
 *
 * States are owned by hash(state key) mod world_size.  Every rank expands the search nodes of its own
 * states, groups the leaf records it found by owner in an OUTBOX in its own HBM, and meets the other
 * ranks once per wave ON THE DEVICES: an all-gather of one small header row per rank plus a barrier,
 * written straight into the peers' memory (peer access inside one process, CUDA IPC across processes).
 * The owners' ingest kernels then read their records directly out of the producers' outboxes over
 * NVLink -- the exchange and the merge are one kernel, nothing is staged, NCCL is not involved.  At the
 * end rank 0 pulls the parts and groups / trims / downloads like a single GPU does.
 *
 * A group is a communicator: form it once (create on every rank, exchange the share blobs by any means
 * -- torch.distributed, MPI, a file --, attach), then solve as often as needed.  One group per rank
 * and process for the multi-process layout (torchrun: one process per GPU); stcsp_gpu_solve_multi runs
 * one host thread per GPU inside a single process (the command-line tool's --gpus N).
 * ------------------------------------------------------------------------------------------- */
typedef struct stcsp_group stcsp_group_t;

typedef struct stcsp_exchange_stats {
    int32_t sharded;                 /* 0: every wave fitted one GPU, rank 0 solved it alone (no exchange) */
    int32_t pad;
    int64_t waves, exchanges;        /* device-side exchanges (one per wave, two when constraint sets had to be resolved) */
    int64_t records;                 /* leaf records this rank ingested */
    int64_t bytes_pulled;            /* bytes this rank read out of its peers' memory over NVLink */
    double exchange_ms;              /* host wall time spent in the exchanges (launch, wait for the slowest rank, read-back) */
} stcsp_exchange_stats_t;

/* device < 0: the current device */
int stcsp_group_create(int32_t rank, int32_t world_size, int32_t device, stcsp_group_t **out);
void stcsp_group_destroy(stcsp_group_t *g);
int64_t stcsp_group_share_bytes(void);
/* blob [share_bytes]: what this rank's peers need to map its exchange block */
int stcsp_group_share(stcsp_group_t *g, void *blob);
/* blobs [world_size * share_bytes]: every rank's blob, in rank order */
int stcsp_group_attach(stcsp_group_t *g, const void *blobs);
/* Collective: every rank of the group calls it with the same problem and options.  Rank 0 receives the automaton
 * (as from stcsp_gpu_solve); on the other ranks *out stays empty. */
int stcsp_group_solve(stcsp_group_t *g, const stcsp_problem_t *problem, const stcsp_options_t *options, stcsp_automaton_t *out,
                      stcsp_exchange_stats_t *stats);
/* The same from one process: n_gpus host threads, one per device (devices == NULL: 0 .. n_gpus - 1). */
int stcsp_gpu_solve_multi(const stcsp_problem_t *problem, const stcsp_options_t *options, int32_t n_gpus, const int32_t *devices,
                          stcsp_automaton_t *out, stcsp_exchange_stats_t *stats);

/* Merge the per-rank parts (parts[r] = part of rank r; arrays may live in caller memory) into one
 * automaton with dense state ids (ascending global id, root = 0) and edges grouped by source;
 * statistics are summed, times are the maximum over parts.  Does not trim. */
int stcsp_automaton_assemble(const stcsp_automaton_t *parts, int32_t n_parts, stcsp_automaton_t *out);

/* Fail rule as a greatest fixpoint (reference src/solveralgorithm.cpp:904-910): marks states without
 * surviving out-edges, removes edges into them, repeats.  In place; edges stay grouped by source. */
int stcsp_automaton_trim(stcsp_automaton_t *a);

#ifdef __cplusplus
}
#endif
#endif /* STCSP_B200_H */
