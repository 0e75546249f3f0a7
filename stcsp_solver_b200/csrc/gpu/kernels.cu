// sm_100a kernels of the search-and-propagate path.
//
//   expand_kernel : one warp per search node.  Stages the node's domain block in shared memory,
//                   propagates to a fixpoint (reference generalisedArcConsistent,
//                   src/solveralgorithm.cpp:617-706), then either fails, emits a leaf record, or
//                   branches on the first unbound variable (solverSolveRe, :911-939; here all values
//                   of the variable at once instead of a binary split -- search order is not a
//                   parity target).
//   ingest_kernel : one warp per leaf record.  Builds the successor signature (:810-837), finds or
//                   inserts it in the open-addressing state table (vertexTableGetVertex/AddVertex,
//                   src/graph.cpp:108-123), appends the edge (edgeNew, src/graph.cpp:78-89) and, for a
//                   new state, writes its first search node (variableAdvanceOneTimeStep,
//                   src/variable.cpp:94-108).
//
// Propagation of a pointwise constraint replaces the reference's per-value nested-loop support
// search (findSupportRe, :435-464) by ONE cooperative enumeration of the tuples over the current
// domains of the still-unbound variables: 32 lanes walk the mixed-radix tuple space, evaluate the
// constraint's bytecode (solverValidateRe, :336-424) and OR the bits of every satisfying tuple into
// per-variable support sets; the new domains are the support sets (domain consistency, at least as
// strong as the reference's bounds revision :476-523).  Enumerations larger than the budget are
// skipped (sound: a skipped propagator only prunes less) and re-armed when a domain shrinks; at a
// leaf every constraint has exactly one tuple, so every leaf is checked exactly.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>

#include "kernels.cuh"

namespace stcsp {

namespace {

typedef unsigned long long u64;

struct WarpMem {
    // node-group state: private to a warp (warp-per-node mode) or shared by the CTA (CTA-per-node mode)
    int32_t *nodew;     // node words (header + domains)
    uint32_t *dirty;    // propagator bitmask: to be revised
    uint32_t *dcur;     // CTA mode: the propagators of the current cooperative round
    uint32_t *hvy;      // cheap-class propagators that turned out to need 32 lanes (many unbound variables)
    int32_t *flag;      // [0] wipe-out, [1] propagators in the current round, [2..3] broadcast of the child base index
    // per-warp scratch of the revision in flight
    u64 *supp;          // [max_scope] support sets of the constraint being revised
    int32_t *flb;       // [max_scope] lb (bytecode) / table stride of free variable r
    int32_t *fdi;       // [max_scope] domain index (var*k+off) of free variable r
    int32_t *cur;       // [max_scope][32] current tuple values per lane
    int32_t *stk;       // [max_stack+1][32] evaluator stack per lane
    uint8_t *vals;      // [max_scope][64] bit positions of the values of each scope variable
    uint8_t *dig;       // [max_scope][32] mixed-radix digits per lane
    uint8_t *fl, *fd, *sd, *rk;   // [max_scope] free slot list, radix, +32 step digits, slot -> rank
};

__host__ __device__ inline size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }

__host__ __device__ inline size_t node_bytes(const DevModel &m) {
    return align8((size_t)m.node_words * 4) + 3 * align8((size_t)m.max_words * 4) + 16;
}

__host__ __device__ inline size_t scratch_bytes(const DevModel &m) {
    size_t b = 0;
    b += (size_t)m.max_scope * 8;
    b += align8((size_t)m.max_scope * 4) * 2;
    b += (size_t)m.max_scope * 32 * 4;
    b += (size_t)(m.max_stack + 1) * 32 * 4;
    b += (size_t)m.max_scope * 64;
    b += (size_t)m.max_scope * 32;
    b += align8((size_t)m.max_scope) * 4;
    return align8(b);
}

// layout of a CTA's dynamic shared memory: m.node_slots node blocks, then kExpandWarps scratch blocks, then the staged set
__device__ inline WarpMem carve(unsigned char *smem, const DevModel &m, int node_slot, int warp) {
    WarpMem w;
    unsigned char *p = smem + (size_t)node_slot * node_bytes(m);
    w.nodew = (int32_t *)p; p += align8((size_t)m.node_words * 4);
    w.dirty = (uint32_t *)p; p += align8((size_t)m.max_words * 4);
    w.dcur = (uint32_t *)p; p += align8((size_t)m.max_words * 4);
    w.hvy = (uint32_t *)p; p += align8((size_t)m.max_words * 4);
    w.flag = (int32_t *)p;
    p = smem + (size_t)m.node_slots * node_bytes(m) + (size_t)warp * scratch_bytes(m);
    w.supp = (u64 *)p; p += (size_t)m.max_scope * 8;
    w.flb = (int32_t *)p; p += align8((size_t)m.max_scope * 4);
    w.fdi = (int32_t *)p; p += align8((size_t)m.max_scope * 4);
    w.cur = (int32_t *)p; p += (size_t)m.max_scope * 32 * 4;
    w.stk = (int32_t *)p; p += (size_t)(m.max_stack + 1) * 32 * 4;
    w.vals = p; p += (size_t)m.max_scope * 64;
    w.dig = p; p += (size_t)m.max_scope * 32;
    w.fl = p; p += align8((size_t)m.max_scope);
    w.fd = p; p += align8((size_t)m.max_scope);
    w.sd = p; p += align8((size_t)m.max_scope);
    w.rk = p;
    return w;
}

__device__ __forceinline__ u64 width_mask(int w) { return w >= 64 ? ~0ull : ((1ull << w) - 1ull); }
__device__ __forceinline__ u64 shift_bits(u64 v, int s) {
    if (s >= 64 || s <= -64) return 0ull;
    return s >= 0 ? (v << s) : (v >> (-s));
}

// ---- branching header ------------------------------------------------------------------------------------------
// Word [3] of a search node names the variables its parent branched on: -1 = first node of a new state (every propagator
// runs), else up to three variable indices, ten bits each, the upper two stored + 1 (0 = none).  Narrow waves branch on
// SEVERAL variables at once (the cartesian product of their domains): the depth of the search tree, not its width, is
// what a narrow instance waits for, and the machine has idle warps for the extra children.
constexpr int kBranchVarBits = 10;
__device__ __forceinline__ int pack_branch(int v0, int v1, int v2) {
    return v0 | ((v1 + 1) << kBranchVarBits) | ((v2 + 1) << (2 * kBranchVarBits));
}
// propagators to run first in a node whose header word is `hdr` (word w of the mask)
__device__ __forceinline__ uint32_t initial_dirty(const DevModel &M, const DevSet &S, int hdr, int w) {
    if (hdr < 0) {
        const int left = S.n_prop - w * 32;
        return left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    const uint32_t *wk = M.wake + S.wake_off + w;
    const int mask = (1 << kBranchVarBits) - 1;
    uint32_t m = wk[((size_t)(hdr & mask) * M.k) * S.n_words];
    const int v1 = ((hdr >> kBranchVarBits) & mask) - 1, v2 = ((hdr >> (2 * kBranchVarBits)) & mask) - 1;
    if (v1 >= 0) m |= wk[((size_t)v1 * M.k) * S.n_words];
    if (v2 >= 0) m |= wk[((size_t)v2 * M.k) * S.n_words];
    return m;
}
// the n-th (0-based) set bit of d, as a one-bit mask
__device__ __forceinline__ u64 nth_bit(u64 d, int n) {
    for (int i = 0; i < n; i++) d &= d - 1;
    return d & (~d + 1ull);
}

// ---- bytecode evaluator: one tuple per lane (reference solverValidateRe) --------------------------
template <int LS>   // LS = distance between consecutive stack/scope slots of one lane (32: warp-interleaved, 1: private)
__device__ __forceinline__ int eval_tuple(const Instr *__restrict__ code, const int32_t *cur, int32_t *stk, int lane,
                                          const DevModel &M) {
    int pc = 0, sp = 0, tos = 0;
    bool valid = true;
    for (;;) {
        const int2 in = reinterpret_cast<const int2 *>(code)[pc];       // generic load: the bytecode may be staged in shared memory
        pc++;
        const int arg = in.y;
        int l;
        switch (in.x) {
            case BC_END: return tos != 0;
            case BC_PUSHC: stk[sp * LS + lane] = tos; sp++; tos = arg; break;
            case BC_PUSHV: stk[sp * LS + lane] = tos; sp++; tos = cur[arg * LS + lane]; break;
            case BC_ARR: {
                const int lo = M.arr_off[arg], n = M.arr_off[arg + 1] - lo;
                if (tos < 0 || tos >= n) { valid = false; tos = 0; }
                else tos = M.arr_val[lo + tos];
                break;
            }
            case BC_ABS: tos = tos < 0 ? (int)(0u - (unsigned)tos) : tos; break;
            case BC_NOT: tos = (tos == 0); break;
            case BC_LT: sp--; l = stk[sp * LS + lane]; tos = valid ? (l < tos) : 0; break;
            case BC_GT: sp--; l = stk[sp * LS + lane]; tos = valid ? (l > tos) : 0; break;
            case BC_LE: sp--; l = stk[sp * LS + lane]; tos = valid ? (l <= tos) : 0; break;
            case BC_GE: sp--; l = stk[sp * LS + lane]; tos = valid ? (l >= tos) : 0; break;
            case BC_EQ: sp--; l = stk[sp * LS + lane]; tos = valid ? (l == tos) : 0; break;
            case BC_NE: sp--; l = stk[sp * LS + lane]; tos = valid ? (l != tos) : 0; break;
            case BC_ADD: sp--; l = stk[sp * LS + lane]; tos = valid ? (int)((unsigned)l + (unsigned)tos) : 0; break;
            case BC_SUB: sp--; l = stk[sp * LS + lane]; tos = valid ? (int)((unsigned)l - (unsigned)tos) : 0; break;
            case BC_MUL: sp--; l = stk[sp * LS + lane]; tos = valid ? (int)((unsigned)l * (unsigned)tos) : 0; break;
            case BC_DIV:
            case BC_MOD:
                sp--; l = stk[sp * LS + lane];
                if (!valid) tos = 0;
                else if (tos == 0 || (l == (int)0x80000000 && tos == -1)) { valid = false; tos = 0; }
                else tos = in.x == BC_DIV ? l / tos : l % tos;
                break;
            case BC_JZ: l = tos; sp--; tos = stk[sp * LS + lane]; if (l == 0) pc = arg; break;
            case BC_JMP: pc = arg; break;
            case BC_AND_SC: if (tos == 0) pc = arg; else { sp--; tos = stk[sp * LS + lane]; } break;
            case BC_OR_SC: if (tos != 0) { tos = 1; pc = arg; } else { sp--; tos = stk[sp * LS + lane]; } break;
            case BC_IMPLY_SC: if (tos == 0) { tos = 1; pc = arg; } break;
            case BC_IMPLY_FIN: sp--; l = stk[sp * LS + lane]; tos = (l <= tos); break;
            default: return 0;
        }
    }
}

__device__ __forceinline__ u64 sat_mul(u64 a, u64 b) {
    const u64 cap = 1ull << 40;
    if (a >= cap || b >= cap) return cap;
    u64 p = a * b;          // < 2^80 would overflow only if both >= 2^24; clamp conservatively
    if (a >= (1ull << 24) && b >= (1ull << 24)) return cap;
    return p > cap ? cap : p;
}

__device__ __forceinline__ u64 reduce_or64(u64 v) {
    return ((u64)__reduce_or_sync(0xffffffffu, (unsigned)(v >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (unsigned)v);
}

struct NodeCtx {
    const DevModel &M;
    const DevSet &S;
    WarpMem &wm;
    u64 *dom;
    int lane;
    int expire;
    unsigned tuples;            // per launch and thread: 32 bits are plenty, and registers are scarce here
    unsigned long long *dbg;    // timeline of this node's propagation (one thread writes), or nullptr
    int dbg_cap;
};

// Timeline entry (tag, time) -- only thread 0 of block 0 ever has a non-null dbg.
__device__ __forceinline__ void dbg_stamp(unsigned long long *dbg, int cap, unsigned long long tag) {
    if (!dbg) return;
    const unsigned long long k = dbg[0];
    if ((long long)(2 * k + 2) >= (long long)cap) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[1 + 2 * k] = tag;
    dbg[2 + 2 * k] = t;
    dbg[0] = k + 1;
}

// A domain of (var, off) shrank: every propagator watching it must run (again).
template <bool CTA>
__device__ __forceinline__ void mark_changed(NodeCtx &c, int var, int off) {
    const uint32_t *wk = c.M.wake + c.S.wake_off + ((size_t)var * c.M.k + off) * c.S.n_words;
    for (int w = c.lane; w < c.S.n_words; w += 32) {
        const uint32_t m = wk[w];
        if (CTA) { if (m) atomicOr(&c.wm.dirty[w], m); }
        else c.wm.dirty[w] |= m;
    }
    __syncwarp();
}

// Intersect domain `idx` with `nd` (warp-uniform arguments).  Domains only ever shrink, and they are only
// written through atomicAnd, so warps of one CTA may revise different propagators of the same node at once:
// a revision that read a stale (larger) domain merely prunes less, and the shrink that it missed re-arms it.
// Returns false on wipe-out.
template <bool CTA>
__device__ __forceinline__ bool shrink_dom(NodeCtx &c, int idx, u64 nd) {
    u64 old;
    if (CTA) {
        old = 0ull;
        if (c.lane == 0) old = atomicAnd(&c.dom[idx], nd);
        old = __shfl_sync(0xffffffffu, old, 0);
    } else {                                            // the warp owns the node: plain read-modify-write
        old = c.dom[idx];
        __syncwarp();
        if (c.lane == 0) c.dom[idx] = old & nd;
    }
    const u64 now = old & nd;
    if (now == 0ull) return false;
    if (now != old) mark_changed<CTA>(c, idx / c.M.k, idx % c.M.k);
    else __syncwarp();
    return true;
}

// Per-lane variant: lanes with `mine` set hold (idx, nd) of different domains.  Returns false on wipe-out.
template <bool CTA>
__device__ __forceinline__ bool shrink_doms(NodeCtx &c, bool mine, int idx, u64 nd) {
    bool changed = false, wiped = false;
    if (mine) {
        u64 old;
        if (CTA) old = atomicAnd(&c.dom[idx], nd);
        else { old = c.dom[idx]; c.dom[idx] = old & nd; }       // distinct domains per lane
        const u64 now = old & nd;
        wiped = now == 0ull;
        changed = now != old;
    }
    if (__any_sync(0xffffffffu, wiped)) return false;
    unsigned cm = __ballot_sync(0xffffffffu, changed);
    while (cm) {
        const int r = __ffs(cm) - 1;
        cm &= cm - 1;
        const int i = __shfl_sync(0xffffffffu, idx, r);
        mark_changed<CTA>(c, i / c.M.k, i % c.M.k);
    }
    __syncwarp();
    return true;
}

// Pointwise constraint at one time offset, bytecode enumeration.  Returns false on wipe-out.
template <bool CTA>
__device__ bool revise_point(NodeCtx &c, const DevCon &con, int off) {
    const DevModel &M = c.M;
    WarpMem &wm = c.wm;
    const int lane = c.lane, n = con.n_scope, k = M.k;
    const int myvar = lane < n ? M.scope[con.scope_off + lane] : 0;
    const u64 myd = lane < n ? c.dom[myvar * k + off] : 1ull;
    const int dsz = __popcll(myd);
    if (__any_sync(0xffffffffu, myd == 0ull)) return false;
    const unsigned fmask = __ballot_sync(0xffffffffu, lane < n && dsz > 1);
    const int nf = __popc(fmask);
    if (nf == 0) {
        // every variable bound: one tuple, evaluated by lane 0 (the exact check every leaf gets)
        if (lane < n) wm.cur[lane * 32] = M.lb[myvar] + __ffsll((long long)myd) - 1;
        __syncwarp();
        int ok = 0;
        if (lane == 0) ok = eval_tuple<32>(M.code + con.code_off, wm.cur, wm.stk, 0, M);
        c.tuples += 1;
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
    }
    float pf = (float)dsz;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) pf *= __shfl_xor_sync(0xffffffffu, pf, s);
    const long long limit = off == 0 ? M.enum_now : M.enum_ahead;
    if (pf > (float)limit) return true;                 // skipped: re-armed when a domain of the scope shrinks
    u64 prod;
    if (pf < 16777216.0f) {
        prod = (u64)pf;                                 // exact below 2^24
    } else {
        prod = (u64)dsz;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) prod = sat_mul(prod, __shfl_xor_sync(0xffffffffu, prod, s));
    }
    if (lane < n) {
        int j = 0;
        for (u64 w = myd; w; w &= w - 1) wm.vals[lane * 64 + j++] = (uint8_t)(__ffsll((long long)w) - 1);
        wm.supp[lane] = 0ull;
        if (dsz > 1) {
            const int r = __popc(fmask & ((1u << lane) - 1u));
            wm.fl[r] = (uint8_t)lane;
            wm.fd[r] = (uint8_t)dsz;
            wm.flb[r] = M.lb[myvar];
            wm.fdi[r] = myvar * k + off;
            wm.rk[lane] = (uint8_t)r;
        } else {
            wm.rk[lane] = 255;
        }
    }
    __syncwarp();
    {
        int t = lane, t32 = 32;
        for (int r = 0; r < nf; r++) {
            const int f = wm.fd[r];
            wm.dig[r * 32 + lane] = (uint8_t)(t % f);
            t /= f;
            if (lane == 0) wm.sd[r] = (uint8_t)(t32 % f);
            t32 /= f;
        }
    }
    __syncwarp();
    for (int i = 0; i < n; i++) {
        const int v = M.scope[con.scope_off + i];
        const int r = wm.rk[i];
        const int d = r == 255 ? 0 : wm.dig[r * 32 + lane];
        wm.cur[i * 32 + lane] = M.lb[v] + wm.vals[i * 64 + d];
    }
    const Instr *code = M.code + con.code_off;
    bool any = false;
    u64 done = 0;
    // the domains as this revision saw them (another warp may shrink them meanwhile)
    const u64 seen_sh = __shfl_sync(0xffffffffu, myd, wm.fl[lane < nf ? lane : 0]);
    const u64 seen = lane < nf ? seen_sh : 0ull;
    for (u64 tbase = 0; tbase < prod; tbase += 32) {
        const bool active = tbase + lane < prod;
        int ok = 0;
        if (active) ok = eval_tuple<32>(code, wm.cur, wm.stk, lane, M);
        if (ok) {
            for (int r = 0; r < nf; r++) {
                const int i = wm.fl[r];
                const u64 bit = 1ull << wm.vals[i * 64 + wm.dig[r * 32 + lane]];
                if (!(wm.supp[i] & bit)) atomicOr(&wm.supp[i], bit);
            }
        }
        any |= __ballot_sync(0xffffffffu, ok) != 0u;
        done = tbase + 32;
        __syncwarp();
        const bool full = lane < nf ? (wm.supp[wm.fl[lane]] == seen) : true;
        if (any && __all_sync(0xffffffffu, full)) break;    // every value already supported: nothing to prune
        int carry = 0;
        for (int r = 0; r < nf; r++) {
            const int f = wm.fd[r];
            int d = wm.dig[r * 32 + lane] + wm.sd[r] + carry;
            carry = d >= f;
            if (carry) d -= f;
            wm.dig[r * 32 + lane] = (uint8_t)d;
            const int i = wm.fl[r];
            wm.cur[i * 32 + lane] = wm.flb[r] + wm.vals[i * 64 + d];
        }
    }
    c.tuples += done < prod ? done : prod;
    if (!any) return false;
    __syncwarp();
    return shrink_doms<CTA>(c, lane < nf, lane < nf ? wm.fdi[lane] : 0, lane < nf ? wm.supp[wm.fl[lane]] : 0ull);
}

// Pointwise constraint with a relation table: the pivot variable's values are handled 64 at a time by one
// table word per prefix tuple; the lanes walk the prefix tuples over the current domains of the other variables.
template <bool CTA>
__device__ bool revise_table(NodeCtx &c, const DevCon &con, int off) {
    const DevModel &M = c.M;
    WarpMem &wm = c.wm;
    const int lane = c.lane, n = con.n_scope, k = M.k, pv = con.pivot;
    const int myvar = lane < n ? M.scope[con.scope_off + lane] : 0;
    const int mystride = lane < n ? M.stride[con.scope_off + lane] : 0;
    const u64 myd = lane < n ? c.dom[myvar * k + off] : 1ull;
    const int dsz = __popcll(myd);
    const u64 Dp = __shfl_sync(0xffffffffu, myd, pv);
    if (__any_sync(0xffffffffu, myd == 0ull)) return false;
    const bool is_pref = lane < n && lane != pv;
    const bool is_free = is_pref && dsz > 1;
    const unsigned fmask = __ballot_sync(0xffffffffu, is_free);
    const int nf = __popc(fmask);
    // bound prefix variables contribute a constant to the table index
    const int base = __reduce_add_sync(0xffffffffu, (is_pref && dsz == 1) ? (__ffsll((long long)myd) - 1) * mystride : 0);
    const u64 *T = M.tables + con.table_off;
    const int pidx = __shfl_sync(0xffffffffu, myvar, pv) * k + off;

    if (nf <= 2) {
        // Fast paths: no enumeration state.  The lanes stand for the bit positions of one free variable (two
        // passes when it is wider than 32); a second free variable is walked value by value, uniformly.
        int ly = -1, lz = -1;                   // scope slots: y = outer (uniform) variable, z = lane variable
        if (nf >= 1) lz = __ffs(fmask) - 1;
        if (nf == 2) {
            ly = 31 - __clz(fmask);
            if (__shfl_sync(0xffffffffu, dsz, ly) > __shfl_sync(0xffffffffu, dsz, lz)) { const int t = ly; ly = lz; lz = t; }
        }
        const u64 Dz = nf >= 1 ? __shfl_sync(0xffffffffu, myd, lz) : 1ull;
        const int sz = nf >= 1 ? __shfl_sync(0xffffffffu, mystride, lz) : 0;
        const u64 Dy = nf == 2 ? __shfl_sync(0xffffffffu, myd, ly) : 1ull;
        const int sy = nf == 2 ? __shfl_sync(0xffffffffu, mystride, ly) : 0;
        const int passes = (Dz >> 32) ? 2 : 1;
        u64 pm = 0ull, supp_z = 0ull, supp_y = 0ull;
        unsigned long long looked = 0;
        for (u64 wy = Dy; wy; wy &= wy - 1) {
            const int py = __ffsll((long long)wy) - 1;
            bool any_y = false;
            for (int h = 0; h < passes; h++) {
                const int pos = lane + 32 * h;
                u64 m = 0ull;
                if ((Dz >> pos) & 1ull) m = __ldg(T + base + py * sy + pos * sz) & Dp;
                const unsigned b = __ballot_sync(0xffffffffu, m != 0ull);
                supp_z |= (u64)b << (32 * h);
                any_y |= b != 0u;
                pm |= m;
            }
            if (any_y) supp_y |= 1ull << py;
            looked += __popcll(Dz);
        }
        c.tuples += looked;
        pm = reduce_or64(pm);
        if (pm == 0ull) return false;
        if (pm != Dp && !shrink_dom<CTA>(c, pidx, pm)) return false;
        if (nf >= 1 && supp_z != Dz && !shrink_dom<CTA>(c, __shfl_sync(0xffffffffu, myvar, lz) * k + off, supp_z)) return false;
        if (nf == 2 && supp_y != Dy && !shrink_dom<CTA>(c, __shfl_sync(0xffffffffu, myvar, ly) * k + off, supp_y)) return false;
        return true;
    }

    // General path: mixed-radix walk over the prefix tuples, 32 per step.
    float pf = is_free ? (float)dsz : 1.0f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) pf *= __shfl_xor_sync(0xffffffffu, pf, s);
    const long long limit = off == 0 ? M.enum_now : M.enum_ahead;
    if (pf > (float)limit) return true;             // skipped: re-armed when a domain of the scope shrinks
    u64 prod;
    if (pf < 16777216.0f) {
        prod = (u64)pf;                             // exact below 2^24
    } else {
        prod = is_free ? (u64)dsz : 1ull;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) prod = sat_mul(prod, __shfl_xor_sync(0xffffffffu, prod, s));
    }
    if (is_free) {
        int j = 0;
        for (u64 w = myd; w; w &= w - 1) wm.vals[lane * 64 + j++] = (uint8_t)(__ffsll((long long)w) - 1);
        const int r = __popc(fmask & ((1u << lane) - 1u));
        wm.fl[r] = (uint8_t)lane;
        wm.fd[r] = (uint8_t)dsz;
        wm.flb[r] = mystride;
        wm.fdi[r] = myvar * k + off;
        wm.supp[r] = 0ull;
    }
    __syncwarp();
    const u64 seen_sh = __shfl_sync(0xffffffffu, myd, wm.fl[lane < nf ? lane : 0]);
    const u64 seen = lane < nf ? seen_sh : 0ull;
    {
        int t = lane, t32 = 32;
        for (int r = 0; r < nf; r++) {
            const int f = wm.fd[r];
            wm.dig[r * 32 + lane] = (uint8_t)(t % f);
            t /= f;
            if (lane == 0) wm.sd[r] = (uint8_t)(t32 % f);
            t32 /= f;
        }
    }
    __syncwarp();
    u64 pm = 0ull;
    bool any = false;
    u64 done = 0;
    for (u64 tbase = 0; tbase < prod; tbase += 32) {
        u64 m = 0ull;
        if (tbase + lane < prod) {
            int idx = base;
            for (int r = 0; r < nf; r++) idx += (int)wm.vals[wm.fl[r] * 64 + wm.dig[r * 32 + lane]] * wm.flb[r];
            m = __ldg(T + idx) & Dp;
        }
        if (m) {
            pm |= m;
            for (int r = 0; r < nf; r++) {
                const u64 bit = 1ull << wm.vals[wm.fl[r] * 64 + wm.dig[r * 32 + lane]];
                if (!(wm.supp[r] & bit)) atomicOr(&wm.supp[r], bit);
            }
        }
        any |= __ballot_sync(0xffffffffu, m != 0ull) != 0u;
        done = tbase + 32;
        if (done >= prod) break;
        __syncwarp();
        const bool full = lane < nf ? (wm.supp[lane] == seen) : true;
        if (reduce_or64(pm) == Dp && __all_sync(0xffffffffu, full)) break;
        int carry = 0;
        for (int r = 0; r < nf; r++) {
            const int f = wm.fd[r];
            int d = wm.dig[r * 32 + lane] + wm.sd[r] + carry;
            carry = d >= f;
            if (carry) d -= f;
            wm.dig[r * 32 + lane] = (uint8_t)d;
        }
    }
    c.tuples += done < prod ? done : prod;
    if (!any) return false;
    __syncwarp();
    pm = reduce_or64(pm);
    const bool mine = lane <= nf;
    const int idx = lane < nf ? wm.fdi[lane] : pidx;
    const u64 nd = lane < nf ? wm.supp[lane] : pm;
    return shrink_doms<CTA>(c, mine, idx, nd);
}

// x == next y: values of y at offset p+1 are the values of x at offset p (reference
// enforceNextConsistency, src/solveralgorithm.cpp:544-593).
template <bool CTA>
__device__ bool revise_next(NodeCtx &c, const DevCon &con) {
    const DevModel &M = c.M;
    const int k = M.k, x = con.x, y = con.y;
    const int s = M.lb[x] - M.lb[y];                    // bit index in y = bit index in x + s
    for (int p = 0; p + 1 < k; p++) {
        const u64 X = c.dom[x * k + p], Y = c.dom[y * k + p + 1];
        const u64 nX = X & shift_bits(Y, -s) & width_mask(M.width[x]);
        const u64 nY = Y & shift_bits(X, s) & width_mask(M.width[y]);
        if (nX == 0ull || nY == 0ull) return false;
        if (nX != X && !shrink_dom<CTA>(c, x * k + p, nX)) return false;
        if (nY != Y && !shrink_dom<CTA>(c, y * k + p + 1, nY)) return false;
    }
    return true;
}

// x until y (reference enforceUntilConsistency, src/solveralgorithm.cpp:598-614)
__device__ bool revise_until(NodeCtx &c, const DevCon &con) {
    if ((c.expire >> con.until_idx) & 1) return true;
    const u64 L = c.dom[con.x * c.M.k], R = c.dom[con.y * c.M.k];
    if (__popcll(L) == 1 && __popcll(R) == 1) {
        const int lv = c.M.lb[con.x] + __ffsll((long long)L) - 1;
        const int rv = c.M.lb[con.y] + __ffsll((long long)R) - 1;
        if (lv != 1 && rv != 1) return false;
    }
    return true;
}

template <bool CTA>
__device__ __forceinline__ bool revise(NodeCtx &c, int q) {
    const DevProp pr = c.M.props[c.S.prop_off + q];
    const DevCon con = c.M.cons[pr.con];
    if (con.kind == DK_POINT)
        return con.pivot >= 0 ? revise_table<CTA>(c, con, pr.offset) : revise_point<CTA>(c, con, pr.offset);
    if (con.kind == DK_NEXT) return revise_next<CTA>(c, con);
    return revise_until(c, con);
}

// ---- scalar revisions: ONE THREAD per propagator ------------------------------------------------------------
// NEXT, UNTIL and relation-table propagators with at most two unbound prefix variables are cheap enough for a
// single thread, so a warp revises up to 32 of them at once (a CTA: 256) instead of spending 32 lanes on one.
// Domains are shared by the lanes: every write is an atomicAnd, every wake an atomicOr.
enum ScalarResult : int { SR_OK = 0, SR_FAIL = 1, SR_HEAVY = 2 };
constexpr int kScalarWalk = 192;        // longest enumeration (prefix tuples) one thread takes on when nodes are plenty
// (DevModel::scalar_walk_cta = 32 when a whole CTA works on one node: the slowest thread is the round's duration, so longer
//  walks go to 32 lanes; measured again at the end of round 2: 16 / 32 / 64 / 128 -> b5_f6 1.09 / 1.02 / 1.15 / 1.15 ms)

__device__ __forceinline__ bool scalar_shrink(const DevModel &M, const DevSet &S, u64 *dom, uint32_t *dirty, int q, int idx,
                                              u64 nd) {
    const u64 old = atomicAnd(&dom[idx], nd);
    const u64 now = old & nd;
    if (now == 0ull) return false;
    if (now != old) {
        __threadfence_block();                          // the shrink is visible before anyone sees the wake
        const uint32_t *wk = M.wake + S.wake_off + (size_t)idx * S.n_words;
        for (int w = 0; w < S.n_words; w++) {
            uint32_t m = wk[w];
            if (w == (q >> 5)) m &= ~(1u << (q & 31));  // a revision is idempotent: it need not wake itself
            if (m) atomicOr(&dirty[w], m);
        }
    }
    return true;
}

__device__ int scalar_revise(const DevModel &M, const DevSet &S, int q, u64 *dom, uint32_t *dirty, int expire,
                             unsigned &tuples, int walk_limit) {
    const DevProp pr = M.props[S.prop_off + q];
    const DevCon con = M.cons[pr.con];
    const int k = M.k, off = pr.offset;
    if (con.kind == DK_NEXT) {
        const int x = con.x, y = con.y;
        const int s = M.lb[x] - M.lb[y];
        for (int p = 0; p + 1 < k; p++) {
            const u64 X = dom[x * k + p], Y = dom[y * k + p + 1];
            const u64 nX = X & shift_bits(Y, -s) & width_mask(M.width[x]);
            const u64 nY = Y & shift_bits(X, s) & width_mask(M.width[y]);
            if (nX == 0ull || nY == 0ull) return SR_FAIL;
            if (nX != X && !scalar_shrink(M, S, dom, dirty, q, x * k + p, nX)) return SR_FAIL;
            if (nY != Y && !scalar_shrink(M, S, dom, dirty, q, y * k + p + 1, nY)) return SR_FAIL;
        }
        return SR_OK;
    }
    if (con.kind == DK_UNTIL) {
        if ((expire >> con.until_idx) & 1) return SR_OK;
        const u64 L = dom[con.x * k], R = dom[con.y * k];
        if (__popcll(L) == 1 && __popcll(R) == 1) {
            const int lv = M.lb[con.x] + __ffsll((long long)L) - 1;
            const int rv = M.lb[con.y] + __ffsll((long long)R) - 1;
            if (lv != 1 && rv != 1) return SR_FAIL;
        }
        return SR_OK;
    }
    // relation table
    const int n = con.n_scope, pv = con.pivot;
    int base = 0, nfree = 0, iy = 0, iz = 0, sy = 0, sz = 0, pidx = 0;
    u64 Dy = 1ull, Dz = 1ull, Dp = 0ull;
    for (int i = 0; i < n; i++) {
        const int idx = M.scope[con.scope_off + i] * k + off;
        const u64 d = dom[idx];
        if (d == 0ull) return SR_FAIL;
        if (i == pv) { Dp = d; pidx = idx; continue; }
        const int st = M.stride[con.scope_off + i];
        if ((d & (d - 1ull)) == 0ull) {
            base += (__ffsll((long long)d) - 1) * st;
        } else {
            if (nfree == 0) { Dz = d; iz = idx; sz = st; }
            else if (nfree == 1) { Dy = d; iy = idx; sy = st; }
            nfree++;
        }
    }
    // Three or more unbound prefix variables: the 32-lane walk (revise_table) would only take it on below its enumeration
    // budget.  Deciding that HERE costs one multiplication per variable; handing it over just to be turned away cost a
    // microsecond of a warp per propagator and round (digitinvader9: twenty such calls per search node).  A skipped
    // propagator is re-armed by the next shrink in its scope, as before.
    if (nfree > 2) {
        float walk = 1.0f;                              // prefix tuples (exact below 2^24; the comparison is all that is needed)
        for (int i = 0; i < n; i++)
            if (i != pv) walk *= (float)__popcll(dom[M.scope[con.scope_off + i] * k + off]);
        return walk > (float)(off == 0 ? M.enum_now : M.enum_ahead) ? SR_OK : SR_HEAVY;
    }
    if (nfree == 2 && __popcll(Dy) > __popcll(Dz)) {            // y: the smaller domain, walked in the outer loop
        const u64 td = Dy; Dy = Dz; Dz = td;
        int t = iy; iy = iz; iz = t;
        t = sy; sy = sz; sz = t;
    }
    if (__popcll(Dy) * __popcll(Dz) > walk_limit) return SR_HEAVY;     // long walks belong to 32 lanes
    const u64 *T = M.tables + con.table_off + base;
    u64 pm = 0ull, supp_y = 0ull, supp_z = 0ull;
    for (u64 wy = Dy; wy; wy &= wy - 1ull) {
        const int py = __ffsll((long long)wy) - 1;
        const u64 *Ty = T + py * sy;
        u64 row = 0ull;
        for (u64 wz = Dz; wz; wz &= wz - 1ull) {
            const int pz = __ffsll((long long)wz) - 1;
            const u64 m = __ldg(Ty + pz * sz) & Dp;
            if (m) { row |= m; supp_z |= 1ull << pz; }
        }
        if (row) { pm |= row; supp_y |= 1ull << py; }
    }
    tuples += (unsigned long long)(__popcll(Dy) * __popcll(Dz));
    if (pm == 0ull) return SR_FAIL;
    if (pm != Dp && !scalar_shrink(M, S, dom, dirty, q, pidx, pm)) return SR_FAIL;
    if (nfree >= 1 && supp_z != Dz && !scalar_shrink(M, S, dom, dirty, q, iz, supp_z)) return SR_FAIL;
    if (nfree == 2 && supp_y != Dy && !scalar_shrink(M, S, dom, dirty, q, iy, supp_y)) return SR_FAIL;
    return SR_OK;
}

__device__ __forceinline__ uint32_t cheap_mask(const DevSet &S, int w) {
    const int left = S.n_cheap - w * 32;
    return left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
}

// Propagate the node to a fixpoint.  Returns true on wipe-out.
//   phase A  rounds of scalar revisions over the dirty cheap propagators, one per thread of the group
//   phase B  what needs 32 lanes (bytecode enumerations, tables with many unbound variables): warp-cooperative,
//            one at a time (warp per node) or dealt to the warps of the CTA (CTA per node); then back to A
//   never       the look-ahead propagators are not run at all (ExpandArgs::skip_ahead)
//   ahead_fail  out: the wipe-out came from a look-ahead propagator (uniform over the group; only set when true is returned)
template <bool CTA>
__device__ bool propagate(NodeCtx &ctx, int gw, int gtid, int gthreads, unsigned &st_rev, unsigned &my_tuples, const bool never,
                          bool &ahead_fail) {
    const DevModel &M = ctx.M;
    const DevSet &S = ctx.S;
    WarpMem &wm = ctx.wm;
    const int lane = ctx.lane;
    // look-ahead row of the wake table: pointwise propagators at offsets >= 1
    const uint32_t *ahead = M.wake + S.wake_off + (size_t)M.V * M.k * S.n_words;
    bool held = M.lazy_ahead != 0 || never;             // look-ahead propagators are held back (uniform over the group)
    ahead_fail = false;
    for (;;) {
        if (held && !never) {
            // released once every variable of the current time point is bound; then all of them run, once
            bool unbound = false;
            for (int v = gtid; v < M.V; v += gthreads) unbound |= __popcll(ctx.dom[v * M.k]) > 1;
            const bool any_unbound = CTA ? __syncthreads_or(unbound) != 0 : __any_sync(0xffffffffu, unbound);
            if (!any_unbound) {
                held = false;
                for (int w = gtid; w < S.n_words; w += gthreads) {
                    const uint32_t m = ahead[w];
                    if (m) atomicOr(&wm.dirty[w], m);
                }
                if (CTA) __syncthreads(); else __syncwarp();
            }
        }
        // ---- phase A
        for (;;) {
            bool myfail = false, my_af = false;
            for (int q = gtid; q < S.n_cheap; q += gthreads) {
                const uint32_t bit = 1u << (q & 31);
                if (!(wm.dirty[q >> 5] & bit)) continue;
                if (held && (ahead[q >> 5] & bit)) continue;
                atomicAnd(&wm.dirty[q >> 5], ~bit);
                __threadfence_block();      // the domains are read AFTER the bit is cleared: a wake that lands in between re-arms
                                            // this propagator instead of being erased while it still sees the old domain
                const int r = scalar_revise(M, S, q, ctx.dom, wm.dirty, ctx.expire, my_tuples, CTA ? M.scalar_walk_cta : M.scalar_walk);
                st_rev++;
                if (r == SR_FAIL) { myfail = true; my_af = (ahead[q >> 5] & bit) != 0u; }
                else if (r == SR_HEAVY) atomicOr(&wm.hvy[q >> 5], bit);
            }
            bool more, bad;
            if (CTA) {
                bad = __syncthreads_or(myfail) != 0;    // barrier: every revision of the round is done
                more = __syncthreads_or(gtid < S.n_words &&
                                        (wm.dirty[gtid] & cheap_mask(S, gtid) & (held ? ~ahead[gtid] : ~0u)) != 0u) != 0;
            } else {
                bad = __any_sync(0xffffffffu, myfail);
                more = __any_sync(0xffffffffu, lane < S.n_words &&
                                                   (wm.dirty[lane] & cheap_mask(S, lane) & (held ? ~ahead[lane] : ~0u)) != 0u);
            }
            dbg_stamp(ctx.dbg, ctx.dbg_cap, 10);           // scalar round done
            if (bad) {
                ahead_fail = CTA ? __syncthreads_or(my_af) != 0 : __any_sync(0xffffffffu, my_af) != 0;
                return true;
            }
            if (!more) break;
        }
        // ---- phase B
        if (!CTA) {
            int q = -1;
            for (int base = 0; base < S.n_words; base += 32) {
                uint32_t w = base + lane < S.n_words ? (wm.hvy[base + lane] | (wm.dirty[base + lane] & ~cheap_mask(S, base + lane))) : 0u;
                if (held && base + lane < S.n_words) w &= ~ahead[base + lane];
                const unsigned b = __ballot_sync(0xffffffffu, w != 0u);
                if (b) {
                    const int l = __ffs(b) - 1;
                    const uint32_t ww = __shfl_sync(0xffffffffu, w, l);
                    q = (base + l) * 32 + __ffs(ww) - 1;
                    break;
                }
            }
            if (q < 0) {                                // fixpoint ...
                if (!held || never) return false;
                bool unbound = false;                   // ... unless the look-ahead propagators were held and are due now
                for (int v = lane; v < M.V; v += 32) unbound |= __popcll(ctx.dom[v * M.k]) > 1;
                if (__any_sync(0xffffffffu, unbound)) return false;
                continue;
            }
            __syncwarp();
            if (lane == 0) {
                wm.hvy[q >> 5] &= ~(1u << (q & 31));
                wm.dirty[q >> 5] &= ~(1u << (q & 31));
            }
            __syncwarp();
            const bool ok = revise<false>(ctx, q);
            __syncwarp();
            if (lane == 0) wm.dirty[q >> 5] &= ~(1u << (q & 31));       // idempotent: its own wake is void
            __syncwarp();
            dbg_stamp(ctx.dbg, ctx.dbg_cap, 30);           // one 32-lane revision done (warp per node)
            st_rev += lane == 0;
            if (!ok) {
                ahead_fail = ((ahead[q >> 5] >> (q & 31)) & 1u) != 0u;
                return true;
            }
        } else {
            if (threadIdx.x == 0) {
                int total = 0;
                for (int w = 0; w < S.n_words; w++) {
                    uint32_t take = wm.hvy[w] | (wm.dirty[w] & ~cheap_mask(S, w));
                    if (held) take &= ~ahead[w];
                    wm.dcur[w] = take;
                    wm.hvy[w] = 0u;
                    wm.dirty[w] &= ~take;
                    total += __popc(take);
                }
                wm.flag[1] = total;
            }
            __syncthreads();
            if (wm.flag[1] == 0) {                      // fixpoint ...
                if (!held || never) return false;
                bool unbound = false;                   // ... unless the held look-ahead propagators are due now
                for (int v = gtid; v < M.V; v += gthreads) unbound |= __popcll(ctx.dom[v * M.k]) > 1;
                if (__syncthreads_or(unbound)) return false;
                continue;
            }
            int seen_bits = 0;
            bool ok = true, my_af = false;
            for (int w = 0; ok && w < S.n_words; w++) {
                uint32_t bits = wm.dcur[w];
                while (ok && bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if ((seen_bits++ % kExpandWarps) != gw) continue;
                    ok = revise<true>(ctx, w * 32 + b);
                    if (!ok) my_af = ((ahead[w] >> b) & 1u) != 0u;
                    __syncwarp();
                    st_rev += lane == 0;
                }
            }
            dbg_stamp(ctx.dbg, ctx.dbg_cap, 20 + (unsigned long long)wm.flag[1]);   // cooperative round done (20 + props)
            if (__syncthreads_or(!ok)) {
                ahead_fail = __syncthreads_or(my_af) != 0;
                return true;
            }
        }
    }
}

// A warp's statistics of one expand pass.  Stand-alone kernels add them to the wave's counters in global memory; inside
// search_kernel they go to the BLOCK's shared-memory totals, which the block flushes once before the grid barrier: eight
// times fewer reductions on the few lines every block reads right after that barrier -- the L2 works such a burst off one
// operation at a time, and the reads queue behind it (3.5 us per wave on juggling_b6_f6_nosym).
enum BlockStat : int { BS_NODES, BS_FAILS, BS_TUPLES, BS_REV, BS_DOM, BS_AHEAD_NODES, BS_AHEAD_FAILS, BS_COUNT };
__device__ __forceinline__ void flush_warp_stats(const ExpandArgs &P, unsigned nodes, unsigned fails, unsigned tuples, unsigned rev,
                                                 unsigned long long dom, unsigned an, unsigned af) {
    if (P.block_stats != nullptr) {
        if (nodes) atomicAdd(&P.block_stats[BS_NODES], nodes);
        if (fails) atomicAdd(&P.block_stats[BS_FAILS], fails);
        if (tuples) atomicAdd(&P.block_stats[BS_TUPLES], tuples);
        if (rev) atomicAdd(&P.block_stats[BS_REV], rev);
        if (dom) atomicAdd(&P.block_stats[BS_DOM], (unsigned)dom);
        if (an) atomicAdd(&P.block_stats[BS_AHEAD_NODES], an);
        if (af) atomicAdd(&P.block_stats[BS_AHEAD_FAILS], af);
        return;
    }
    if (nodes) atomicAdd(&P.counters[C_NODES], (unsigned long long)nodes);
    if (fails) atomicAdd(&P.counters[C_FAILS], (unsigned long long)fails);
    if (tuples) atomicAdd(&P.counters[C_TUPLES], (unsigned long long)tuples);
    if (rev) atomicAdd(&P.counters[C_REVISIONS], (unsigned long long)rev);
    if (dom) atomicAdd(&P.counters[C_DOMINANCE], dom);
    if (an && P.ahead_stats != nullptr) {
        atomicAdd(&P.ahead_stats[C_AHEAD_NODES], (unsigned long long)an);
        if (af) atomicAdd(&P.ahead_stats[C_AHEAD_FAILS], (unsigned long long)af);
    }
}

// ---- constraint-set metadata staged in shared memory ------------------------------------------------------
// The propagators of a set (descriptors, scopes, strides, wake masks, bytecode, variable bounds) are read over and
// over by every revision; the CTA copies the set most of its nodes belong to into shared memory once per wave and
// hands the revisions a DevModel whose pointers are biased so that the ABSOLUTE pool indices still work.
__device__ __forceinline__ unsigned char *stage_base(unsigned char *smem, const DevModel &m) {
    return smem + (size_t)m.node_slots * node_bytes(m) + (size_t)kExpandWarps * scratch_bytes(m);
}

// Called by every thread of the CTA (contains barriers).  Returns the constraint set staged, -1 if none.
// resident (search_kernel only): the set this CTA staged in an earlier wave of the same launch is still there.
__device__ int stage_set(const DevModel &M, unsigned char *smem, int cid, int *resident) {
    DevModel *sm = reinterpret_cast<DevModel *>(stage_base(smem, M));
    if (M.stage_bytes == 0 || cid < 0) return -1;
    if (resident && *resident == cid) return cid;       // (read before the barrier below, written after it)
    unsigned char *p = reinterpret_cast<unsigned char *>(sm) + align8(sizeof(DevModel));
    const DevSet S = M.sets[cid];
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t wake_words = ((size_t)M.V * M.k + 1) * S.n_words;      // + the look-ahead row
    int2 *s_props = reinterpret_cast<int2 *>(p); p += align8((size_t)S.n_prop * sizeof(DevProp));
    int2 *s_cons = reinterpret_cast<int2 *>(p); p += align8((size_t)S.n_con * sizeof(DevCon));
    int32_t *s_scope = reinterpret_cast<int32_t *>(p); p += align8((size_t)S.n_scope * 4);
    int32_t *s_stride = reinterpret_cast<int32_t *>(p); p += align8((size_t)S.n_scope * 4);
    uint32_t *s_wake = reinterpret_cast<uint32_t *>(p); p += align8(wake_words * 4);
    int32_t *s_lb = reinterpret_cast<int32_t *>(p); p += align8((size_t)M.V * 4);
    int32_t *s_width = reinterpret_cast<int32_t *>(p); p += align8((size_t)M.V * 4);
    int2 *s_code = reinterpret_cast<int2 *>(p);
    __syncthreads();                                    // nobody still reads the previous staging
    const int2 *g_props = reinterpret_cast<const int2 *>(M.props + S.prop_off);
    for (int i = tid; i < S.n_prop; i += nt) s_props[i] = g_props[i];
    const int2 *g_cons = reinterpret_cast<const int2 *>(M.cons + S.con_off);
    for (int i = tid; i < S.n_con * (int)(sizeof(DevCon) / 8); i += nt) s_cons[i] = g_cons[i];
    for (int i = tid; i < S.n_scope; i += nt) {
        s_scope[i] = M.scope[S.scope_off + i];
        s_stride[i] = M.stride[S.scope_off + i];
    }
    for (int i = tid; i < (int)wake_words; i += nt) s_wake[i] = M.wake[S.wake_off + i];
    for (int i = tid; i < M.V; i += nt) {
        s_lb[i] = M.lb[i];
        s_width[i] = M.width[i];
    }
    const int2 *g_code = reinterpret_cast<const int2 *>(M.code + S.code_off);
    for (int i = tid; i < S.n_code; i += nt) s_code[i] = g_code[i];
    if (tid == 0) {
        DevModel m = M;
        m.props = reinterpret_cast<const DevProp *>(s_props) - S.prop_off;
        m.cons = reinterpret_cast<const DevCon *>(s_cons) - S.con_off;
        m.scope = s_scope - S.scope_off;
        m.stride = s_stride - S.scope_off;
        m.wake = s_wake - S.wake_off;
        m.lb = s_lb;
        m.width = s_width;
        m.code = reinterpret_cast<const Instr *>(s_code) - S.code_off;
        *sm = m;
        if (resident) *resident = cid;
    }
    __syncthreads();
    return cid;
}

// CTA = true : one CTA per search node, its warps revise different dirty propagators of the node concurrently
//              (narrow waves: fewer nodes than resident CTAs, the latency of one node is the wave's duration).
// CTA = false: one warp per search node (wide waves: throughput).
// Leaf found inside expand (search_kernel on narrow waves): route it and merge it into the automaton right here.  Out of line:
// it is off the propagation path and must not cost the expand bodies registers.  Defined behind route_leaf / ingest_leaf.
__device__ __noinline__ void fused_leaf(const DevModel &M, const RouteArgs &R, const IngestArgs &I, int32_t *rec, long long li,
                                        int lane, unsigned long long *st_dom);

template <bool CTA>
__device__ __forceinline__ void expand_body(const DevModel &Mg, const ExpandArgs &P, unsigned char *smem, int *resident = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpMem wm = carve(smem, Mg, CTA ? 0 : warp, warp);
    const long long n_in = P.n_in;
    const long long first = CTA ? (long long)blockIdx.x : (long long)blockIdx.x * kExpandWarps + warp;
    const long long step = CTA ? (long long)gridDim.x : (long long)gridDim.x * kExpandWarps;
    const int V = Mg.V, k = Mg.k, NW = Mg.node_words;
    const int gwarps = CTA ? kExpandWarps : 1;          // warps working on one node
    const int gw = CTA ? warp : 0;                      // this warp's index among them
    const int gtid = CTA ? threadIdx.x : lane, gthreads = gwarps * 32;
    unsigned st_nodes = 0, st_fails = 0, st_tuples = 0, st_rev = 0, my_tuples = 0;    // per launch and thread
    unsigned st_an = 0, st_af = 0;                      // nodes that ran their look-ahead propagators / failed because of one
    unsigned long long st_dom = 0;                      // fused leaves that hit an existing state
    // stage the constraint set of this CTA's first node (waves are almost always homogeneous)
    const long long probe = CTA ? (long long)blockIdx.x : (long long)blockIdx.x * kExpandWarps;
    const int staged = stage_set(Mg, smem, probe < n_in ? (Mg.n_sets == 1 ? 0 : P.in_nodes[probe * NW + 1]) : -1, resident);
    const DevModel &Ms = *reinterpret_cast<const DevModel *>(stage_base(smem, Mg));
    if (blockIdx.x == 0 && threadIdx.x == 0) dbg_stamp(P.dbg, P.dbg_cap, 0);     // wave entered, set staged

    for (long long ni = first; ni < n_in; ni += step) {
        const int32_t *src = P.in_nodes + ni * NW;
        if (CTA) __syncthreads();                       // the previous node's shared state is no longer in use
        for (int w = gtid; w < NW; w += gthreads) wm.nodew[w] = src[w];
        if (gtid == 0) wm.flag[0] = 0;
        for (int w = gtid; w < Mg.max_words; w += gthreads) wm.hvy[w] = 0u;
        if (CTA) __syncthreads(); else __syncwarp();
        const int cid = wm.nodew[1], bvar = wm.nodew[3];
        const DevModel &M = cid == staged ? Ms : Mg;    // metadata from shared memory when this node's set is the staged one
        const DevSet S = Mg.sets[cid];
        unsigned long long *dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? P.dbg : nullptr;    // (warp mode: warp 0's nodes)
        NodeCtx ctx{M, S, wm, reinterpret_cast<u64 *>(wm.nodew + 4), lane, wm.nodew[2], 0ull, dbg, P.dbg_cap};
        dbg_stamp(dbg, P.dbg_cap, 1);                   // node loaded
        u64 *dom = ctx.dom;

        bool empty = false;
        for (int i = lane; i < V * k; i += 32) empty |= dom[i] == 0ull;
        bool fail = __any_sync(0xffffffffu, empty);

        for (int w = gtid; w < S.n_words; w += gthreads) wm.dirty[w] = initial_dirty(M, S, bvar, w);
        if (CTA) __syncthreads(); else __syncwarp();

        const bool never = P.skip_ahead != 0;
        bool ahead_fail = false;
        if (!fail) {
            fail = propagate<CTA>(ctx, gw, gtid, gthreads, st_rev, my_tuples, never, ahead_fail);   // `fail` is uniform over the group
            if (gw == 0 && !never) { st_an++; st_af += fail && ahead_fail; }
        }
        dbg_stamp(dbg, P.dbg_cap, 2);                   // propagated
        st_tuples += ctx.tuples;
        if (gw == 0) st_nodes++;
        if (fail) { if (gw == 0) st_fails++; continue; }

        int bv = -1;
        for (int base = 0; base < V; base += 32) {
            const int v = base + lane;
            const bool unbound = v < V && __popcll(dom[v * k]) > 1;
            const unsigned b = __ballot_sync(0xffffffffu, unbound);
            if (b) { bv = base + __ffs(b) - 1; break; }
        }
        if (bv < 0) {
            // leaf: every variable bound at the current time point
            if (gw != 0) continue;
            unsigned long long li = 0;
            if (lane == 0) li = atomicAdd(&P.counters[C_LEAVES], 1ull);
            li = __shfl_sync(0xffffffffu, li, 0);
            if ((long long)li >= P.leaf_cap) {
                if (lane == 0) atomicOr(&P.counters[C_OVERFLOW], 2ull);
                continue;
            }
            int32_t *rec = P.leaves + li * M.rec_words;
            if (lane < 4) rec[lane] = lane == 3 ? 0 : wm.nodew[lane];
            for (int v = lane; v < V; v += 32) rec[4 + v] = M.lb[v] + __ffsll((long long)dom[v * k]) - 1;
            if (P.fuse_route) {
                __syncwarp();                           // the record is complete for every lane of this warp
                fused_leaf(Mg, *P.fuse_route, *P.fuse_ingest, rec, (long long)li, lane, &st_dom);
                dbg_stamp(dbg, P.dbg_cap, 5);           // leaf routed and merged
            }
        } else {
            // branch: on the first unbound variable, and -- while the wave is narrow enough that every child still gets a
            // warp of its own (P.fan, set per wave) -- on the next one or two as well
            int bv1 = -1, bv2 = -1;
            u64 D1 = 1ull, D2 = 1ull;
            const u64 D = dom[bv * k];
            int d = __popcll(D);
            if (d < P.fan) {
                for (int base = 0; base < V && bv2 < 0; base += 32) {
                    const int v = base + lane;
                    const bool unbound = v < V && v > bv && __popcll(dom[v * k]) > 1;
                    unsigned b = __ballot_sync(0xffffffffu, unbound);
                    while (b && bv2 < 0) {
                        const int cand = base + __ffs(b) - 1;
                        b &= b - 1;
                        const u64 Dc = dom[cand * k];
                        if ((long long)d * __popcll(Dc) > (long long)P.fan) { bv2 = -2; break; }       // would be too many: stop here
                        if (bv1 < 0) { bv1 = cand; D1 = Dc; } else { bv2 = cand; D2 = Dc; }
                        d *= __popcll(Dc);
                    }
                }
                if (bv2 == -2) bv2 = -1;
            }
            const int d0 = __popcll(D), d1 = __popcll(D1);
            unsigned long long base = 0;
            if (CTA) {
                if (threadIdx.x == 0) {
                    const unsigned long long b0 = atomicAdd(&P.counters[C_OUT], (unsigned long long)d);
                    wm.flag[2] = (int)(b0 & 0xffffffffull);
                    wm.flag[3] = (int)(b0 >> 32);
                }
                __syncthreads();
                base = ((unsigned long long)(unsigned)wm.flag[3] << 32) | (unsigned)wm.flag[2];
            } else {
                if (lane == 0) base = atomicAdd(&P.counters[C_OUT], (unsigned long long)d);
                base = __shfl_sync(0xffffffffu, base, 0);
            }
            if ((long long)(base + d) > P.out_cap) {
                if (gtid == 0) atomicOr(&P.counters[C_OVERFLOW], 1ull);
                continue;
            }
            dbg_stamp(dbg, P.dbg_cap, 3);               // branching variables chosen, room reserved
            const int hdr = pack_branch(bv, bv1, bv2);
            const int dw = 4 + 2 * (bv * k), dw1 = bv1 >= 0 ? 4 + 2 * (bv1 * k) : -1, dw2 = bv2 >= 0 ? 4 + 2 * (bv2 * k) : -1;
            // Writing a child takes ~130 dependent instructions if a warp does it word by word with selects (0.6 us per
            // child: 25 us for the 343 children of the root of juggling_b6_f6_nosym).  Two passes instead:
            //   copy   every lane OWNS word indices lane, lane + 32, ... : it reads them from the parent once and stores them
            //          into one child after the other (address arithmetic and stores only);
            //   patch  every lane owns a CHILD: it works out that child's one-bit domains and overwrites the seven words
            //          that differ from the parent.
            const int my_children = (d - gw + gwarps - 1) / gwarps;        // children gw, gw + gwarps, ...
            for (int c0 = 0; c0 < NW; c0 += 128) {
                const int i0 = c0 + lane, i1 = i0 + 32, i2 = i0 + 64, i3 = i0 + 96;
                const int32_t w0 = i0 < NW ? wm.nodew[i0] : 0, w1 = i1 < NW ? wm.nodew[i1] : 0, w2 = i2 < NW ? wm.nodew[i2] : 0,
                              w3 = i3 < NW ? wm.nodew[i3] : 0;
                int32_t *dst = P.out_nodes + (base + gw) * NW;
                const long long stride = (long long)gwarps * NW;
                for (int t = 0; t < my_children; t++, dst += stride) {
                    if (i0 < NW) dst[i0] = w0;
                    if (i1 < NW) dst[i1] = w1;
                    if (i2 < NW) dst[i2] = w2;
                    if (i3 < NW) dst[i3] = w3;
                }
            }
            __syncwarp();                               // the patches below overwrite words other lanes have just stored
            for (int r0 = 0; r0 < my_children; r0 += 32) {
                if (r0 + lane >= my_children) continue;
                const int j = gw + (r0 + lane) * gwarps;
                int32_t *dst = P.out_nodes + (base + j) * NW;
                const u64 m0 = nth_bit(D, j % d0);
                dst[3] = hdr;
                dst[dw] = (int32_t)(uint32_t)(m0 & 0xffffffffull);
                dst[dw + 1] = (int32_t)(uint32_t)(m0 >> 32);
                if (bv1 >= 0) {
                    const u64 m1 = nth_bit(D1, (j / d0) % d1);
                    dst[dw1] = (int32_t)(uint32_t)(m1 & 0xffffffffull);
                    dst[dw1 + 1] = (int32_t)(uint32_t)(m1 >> 32);
                }
                if (bv2 >= 0) {
                    const u64 m2 = nth_bit(D2, j / (d0 * d1));
                    dst[dw2] = (int32_t)(uint32_t)(m2 & 0xffffffffull);
                    dst[dw2 + 1] = (int32_t)(uint32_t)(m2 >> 32);
                }
            }
            dbg_stamp(dbg, P.dbg_cap, 4);               // children written (this warp's share)
        }
    }
    // st_rev / my_tuples are per thread (scalar revisions), st_tuples is warp-uniform (cooperative revisions)
    st_rev = __reduce_add_sync(0xffffffffu, st_rev);
    st_tuples += __reduce_add_sync(0xffffffffu, my_tuples);
    if (lane == 0) flush_warp_stats(P, st_nodes, st_fails, st_tuples, st_rev, st_dom, P.ahead_stats != nullptr ? st_an : 0u, st_af);
}

// ---- wide waves: FOUR search nodes per warp, eight lanes each ----------------------------------------------------
// A scalar round rarely has more than a handful of dirty propagators per node, so with one node per warp most lanes
// idle.  Here the four groups of a warp run their scalar rounds side by side (lanes that execute the same revision
// code converge regardless of their group); the rare revisions that need 32 lanes are served one group at a time.
__device__ __forceinline__ void expand_body_quad(const DevModel &Mg, const ExpandArgs &P, unsigned char *smem, int *resident = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
    const unsigned gmask = 0xffu << (8 * g);
    WarpMem wm = carve(smem, Mg, warp * 4 + g, warp);           // my group's node slot, the warp's scratch
    const long long n_in = P.n_in;
    const int V = Mg.V, k = Mg.k, NW = Mg.node_words;
    unsigned st_nodes = 0, st_fails = 0, st_tuples = 0, st_rev = 0, my_tuples = 0;    // per launch and thread
    unsigned st_an = 0, st_af = 0;                      // nodes that ran their look-ahead propagators / failed because of one
    const long long probe = (long long)blockIdx.x * kExpandWarps * 4;
    const int staged = stage_set(Mg, smem, probe < n_in ? (Mg.n_sets == 1 ? 0 : P.in_nodes[probe * NW + 1]) : -1, resident);
    const DevModel &Ms = *reinterpret_cast<const DevModel *>(stage_base(smem, Mg));
    const long long first = ((long long)blockIdx.x * kExpandWarps + warp) * 4, step = (long long)gridDim.x * kExpandWarps * 4;

    for (long long q4 = first; q4 < n_in; q4 += step) {
        const long long ni = q4 + g;
        const bool have = ni < n_in;
        __syncwarp();
        if (have) {
            const int32_t *src = P.in_nodes + ni * NW;
            for (int w = gl; w < NW; w += 8) wm.nodew[w] = src[w];
        } else if (gl < 4) {
            wm.nodew[gl] = gl == 3 ? -1 : 0;                    // a harmless header for the idle group
        }
        for (int w = gl; w < Mg.max_words; w += 8) { wm.hvy[w] = 0u; wm.dirty[w] = 0u; }
        __syncwarp();
        const int cid = wm.nodew[1], bvar = wm.nodew[3], expire = wm.nodew[2];
        const DevModel &M = cid == staged ? Ms : Mg;
        const DevSet S = Mg.sets[cid];
        u64 *dom = reinterpret_cast<u64 *>(wm.nodew + 4);
        bool empty = false;
        if (have)
            for (int i = gl; i < V * k; i += 8) empty |= dom[i] == 0ull;
        bool fail = (__ballot_sync(0xffffffffu, empty) & gmask) != 0u;
        if (have)
            for (int w = gl; w < S.n_words; w += 8) wm.dirty[w] = initial_dirty(M, S, bvar, w);
        __syncwarp();

        // ---- propagate the four nodes in lock step (the host never picks this mode with lazy look-ahead)
        // never: this group's node does not run its look-ahead propagators (ExpandArgs::skip_ahead); their dirty bits stay set
        // and are masked out wherever work is looked for
        const bool never = P.skip_ahead != 0;
        const uint32_t *ahead = M.wake + S.wake_off + (size_t)V * k * S.n_words;
        bool done = !have || fail, ahead_fail = false;
        for (;;) {
            // phase A: one scalar round for every group that still has cheap work
            bool myfail = false, my_af = false;
            if (!done) {
                for (int q = gl; q < S.n_cheap; q += 8) {
                    const uint32_t bit = 1u << (q & 31);
                    if (!(wm.dirty[q >> 5] & bit)) continue;
                    if (never && (ahead[q >> 5] & bit)) continue;
                    atomicAnd(&wm.dirty[q >> 5], ~bit);
                    __threadfence_block();  // as in propagate(): clear the bit, THEN read the domains
                    const int r = scalar_revise(M, S, q, dom, wm.dirty, expire, my_tuples, kScalarWalk);
                    st_rev++;
                    if (r == SR_FAIL) { myfail = true; my_af = (ahead[q >> 5] & bit) != 0u; }
                    else if (r == SR_HEAVY) atomicOr(&wm.hvy[q >> 5], bit);
                }
            }
            __syncwarp();
            bool p_more = false, p_heavy = false;
            if (!done) {
                for (int w = gl; w < S.n_words; w += 8) {
                    const uint32_t cm = cheap_mask(S, w), keep = never ? ~ahead[w] : ~0u;
                    p_more |= (wm.dirty[w] & cm & keep) != 0u;
                    p_heavy |= ((wm.hvy[w] | (wm.dirty[w] & ~cm)) & keep) != 0u;
                }
            }
            const unsigned bf = __ballot_sync(0xffffffffu, myfail);
            const unsigned bm = __ballot_sync(0xffffffffu, p_more);
            const unsigned bh = __ballot_sync(0xffffffffu, p_heavy);
            const unsigned ba = __ballot_sync(0xffffffffu, my_af);
            if (!done && (bf & gmask)) { fail = true; done = true; ahead_fail = (ba & gmask) != 0u; }
            const bool more = !done && (bm & gmask) != 0u;
            const bool heavy = !done && !more && (bh & gmask) != 0u;
            if (!done && !more && !heavy) done = true;          // this node is at its fixpoint
            const unsigned any_more = __ballot_sync(0xffffffffu, more);
            const unsigned any_heavy = __ballot_sync(0xffffffffu, heavy);
            if (any_more) continue;                             // some group has another scalar round to run
            if (!any_heavy) break;                              // every node is at its fixpoint (or failed)
            // phase B: revisions that need 32 lanes, one group at a time
            for (int gg = 0; gg < 4; gg++) {
                if (!(any_heavy & (0xffu << (8 * gg)))) continue;       // uniform
                WarpMem wq = carve(smem, Mg, warp * 4 + gg, warp);
                const int cq = wq.nodew[1];
                const DevModel &Mq = cq == staged ? Ms : Mg;
                const DevSet Sq = Mg.sets[cq];
                const uint32_t *aheadq = Mq.wake + Sq.wake_off + (size_t)V * k * Sq.n_words;
                const bool neverq = __shfl_sync(0xffffffffu, (int)never, gg * 8) != 0;
                int q = -1;
                for (int base = 0; base < Sq.n_words; base += 32) {
                    uint32_t w = base + lane < Sq.n_words ? (wq.hvy[base + lane] | (wq.dirty[base + lane] & ~cheap_mask(Sq, base + lane))) : 0u;
                    if (neverq && base + lane < Sq.n_words) w &= ~aheadq[base + lane];
                    const unsigned b = __ballot_sync(0xffffffffu, w != 0u);
                    if (b) {
                        const int l = __ffs(b) - 1;
                        const uint32_t ww = __shfl_sync(0xffffffffu, w, l);
                        q = (base + l) * 32 + __ffs(ww) - 1;
                        break;
                    }
                }
                if (q < 0) continue;
                __syncwarp();
                if (lane == 0) {
                    wq.hvy[q >> 5] &= ~(1u << (q & 31));
                    wq.dirty[q >> 5] &= ~(1u << (q & 31));
                }
                __syncwarp();
                NodeCtx cx{Mq, Sq, wq, reinterpret_cast<u64 *>(wq.nodew + 4), lane, wq.nodew[2], 0ull, nullptr, 0};
                const bool ok = revise<false>(cx, q);
                __syncwarp();
                if (lane == 0) wq.dirty[q >> 5] &= ~(1u << (q & 31));
                __syncwarp();
                st_rev += lane == 0;
                st_tuples += cx.tuples;
                if (!ok && g == gg) { fail = true; done = true; ahead_fail = ((aheadq[q >> 5] >> (q & 31)) & 1u) != 0u; }
            }
        }

        // ---- emit: eight lanes per node
        if (have && gl == 0) {
            st_nodes++;
            st_fails += fail;
            if (!never) { st_an++; st_af += fail && ahead_fail; }
        }
        const bool live = have && !fail;
        int bv = -1;
        for (int base = 0; base < V; base += 8) {               // uniform trip count
            const int v = base + gl;
            const bool unbound = live && v < V && __popcll(dom[v * k]) > 1;
            const unsigned b = (__ballot_sync(0xffffffffu, unbound) & gmask) >> (8 * g);
            if (bv < 0 && b) bv = base + __ffs(b) - 1;
        }
        const bool leaf = live && bv < 0, branch = live && bv >= 0;
        const u64 D = branch ? dom[bv * k] : 0ull;
        const int d = __popcll(D);
        // one cursor update per warp and kind (single-address atomics are the scarce resource on wide waves)
        unsigned long long slot = 0;
        {
            const unsigned lb = __ballot_sync(0xffffffffu, leaf && gl == 0);
            int before = 0, total = 0;                          // children of the groups below mine / of the warp
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int dq = __shfl_sync(0xffffffffu, d, q * 8);
                if (q < g) before += dq;
                total += dq;
            }
            unsigned long long lbase = 0, obase = 0;
            if (lane == 0) {
                if (lb) lbase = atomicAdd(&P.counters[C_LEAVES], (unsigned long long)__popc(lb));
                if (total) obase = atomicAdd(&P.counters[C_OUT], (unsigned long long)total);
            }
            lbase = __shfl_sync(0xffffffffu, lbase, 0);
            obase = __shfl_sync(0xffffffffu, obase, 0);
            slot = leaf ? lbase + (unsigned long long)__popc(lb & ((1u << (g * 8)) - 1u)) : obase + (unsigned long long)before;
        }
        if (leaf) {
            if ((long long)slot >= P.leaf_cap) {
                if (gl == 0) atomicOr(&P.counters[C_OVERFLOW], 2ull);
            } else {
                int32_t *rec = P.leaves + slot * M.rec_words;
                if (gl < 4) rec[gl] = gl == 3 ? 0 : wm.nodew[gl];
                for (int v = gl; v < V; v += 8) rec[4 + v] = M.lb[v] + __ffsll((long long)dom[v * k]) - 1;
            }
        } else if (branch) {
            if ((long long)(slot + d) > P.out_cap) {
                if (gl == 0) atomicOr(&P.counters[C_OVERFLOW], 1ull);
            } else {
                const int dw = 4 + 2 * (bv * k);
                int j = 0;
                for (u64 w = D; w; w &= w - 1, j++) {
                    const u64 one = w & (~w + 1ull);
                    int32_t *dst = P.out_nodes + (slot + j) * NW;
                    for (int i = gl; i < NW; i += 8) {
                        int32_t val = wm.nodew[i];
                        if (i == 3) val = pack_branch(bv, -1, -1);
                        else if (i == dw) val = (int32_t)(uint32_t)(one & 0xffffffffull);
                        else if (i == dw + 1) val = (int32_t)(uint32_t)(one >> 32);
                        dst[i] = val;
                    }
                }
            }
        }
    }
    st_rev = __reduce_add_sync(0xffffffffu, st_rev);
    st_nodes = __reduce_add_sync(0xffffffffu, st_nodes);
    st_fails = __reduce_add_sync(0xffffffffu, st_fails);
    st_tuples += __reduce_add_sync(0xffffffffu, my_tuples);
    st_an = __reduce_add_sync(0xffffffffu, st_an);
    st_af = __reduce_add_sync(0xffffffffu, st_af);
    if (lane == 0) flush_warp_stats(P, st_nodes, st_fails, st_tuples, st_rev, 0ull, P.ahead_stats != nullptr ? st_an : 0u, st_af);
}

__global__ void __launch_bounds__(kExpandWarps * 32, kExpandCtasPerSm) expand_quad_kernel(const DevModel M, const ExpandArgs P) {
    extern __shared__ __align__(16) unsigned char smem[];
    expand_body_quad(M, P, smem);
}

template <bool CTA>
__global__ void __launch_bounds__(kExpandWarps * 32, kExpandCtasPerSm) expand_kernel(const DevModel M, const ExpandArgs P) {
    extern __shared__ __align__(16) unsigned char smem[];
    expand_body<CTA>(M, P, smem);
}

// ---- hashing ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

// Order-independent over the words, so a warp can hash a key cooperatively with an xor-reduction.
__host__ __device__ __forceinline__ uint32_t key_word_hash(int32_t word, int j) {
    return mix32((uint32_t)word * 0x9e3779b1u + (uint32_t)j * 0x7f4a7c15u + 0x165667b1u);
}

__host__ __device__ inline uint32_t owner_hash(uint32_t h) { return mix32(h ^ 0x5bd1e995u); }

__host__ __device__ inline uint32_t cap_hash_begin(int cid) { return 0x9e3779b9u * (uint32_t)(cid + 1); }
__host__ __device__ inline uint32_t cap_hash_step(uint32_t h, int32_t v) {
    return h ^ ((uint32_t)v + 0x9e3779b9u + (h << 6) + (h >> 2));
}
__host__ __device__ inline uint32_t cap_hash_end(uint32_t h) {
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    return h;
}
__host__ __device__ inline uint32_t cap_hash(int cid, const int32_t *vals, int n) {
    uint32_t h = cap_hash_begin(cid);
    for (int i = 0; i < n; i++) h = cap_hash_step(h, vals[i]);
    return cap_hash_end(h);
}

// Rank that owns the successor state of a ROUTED record.  The root (constraint set 0, empty signature) was created on
// rank 0 before the search started, so a stateless model's only key must map there too.
__device__ __forceinline__ int leaf_owner(const DevModel &M, int ncid, uint32_t h) {
    if (M.world <= 1 || (M.sig_len == 0 && ncid == 0)) return 0;
    return (int)(owner_hash(h) % (uint32_t)M.world);
}

// word j of the state key of the successor described by a routed record
__device__ __forceinline__ int key_word(const DevModel &M, const int32_t *rec, int ncid, int nexp, int j) {
    if (j == 0) return ncid;
    if (j <= M.n_sig) return rec[4 + M.sig_vars[j - 1]];
    return (nexp >> (j - 1 - M.n_sig)) & 1;
}

// ---- route: successor constraint set, until flags and state-key hash of every leaf --------------------
// (reference src/solveralgorithm.cpp:755-837: constraint rewriting -> constraintID, signature)
// Route one leaf (all lanes of the warp).  Returns false when the successor constraint set is not known yet:
// the leaf is then listed for the host.
__device__ __forceinline__ bool route_leaf(const DevModel &M, const RouteArgs &P, int32_t *rec, long long li, int lane) {
    const int KW = M.key_words;
    const int cid = rec[1], exp = rec[2];
    const DevSet S = M.sets[cid];
    int ncid = S.static_next;
    if (ncid < 0) {
        // the successor set depends on the values captured by `first`: host-filled map
        int found = -1;
        if (lane == 0 && P.capmap_mask >= 0) {
            const int32_t *cap = M.aux + S.cap_off;
            uint32_t h = cap_hash_begin(cid);
            for (int i = 0; i < S.n_cap; i++) h = cap_hash_step(h, rec[4 + cap[i]]);
            h = cap_hash_end(h) & (uint32_t)P.capmap_mask;
            for (;;) {
                const CapEntry e = P.capmap[h];
                if (e.cid == -1) break;
                bool eq = e.cid == cid;
                for (int i = 0; eq && i < S.n_cap; i++) eq = P.capvals[e.off + i] == rec[4 + cap[i]];
                if (eq) { found = e.next; break; }
                h = (h + 1) & (uint32_t)P.capmap_mask;
            }
        }
        found = __shfl_sync(0xffffffffu, found, 0);
        if (found < 0) {
            if (lane == 0) {
                const unsigned long long u = atomicAdd(&P.counters[C_UNRESOLVED], 1ull);
                if ((long long)u < P.unresolved_cap) P.unresolved[u] = (int32_t)li;
                else atomicOr(&P.counters[C_OVERFLOW], 16ull);
            }
            return false;
        }
        ncid = found;
    }
    const DevSet NS = M.sets[ncid];
    int nexp = exp;                                   // until flags (src/solveralgorithm.cpp:821-834)
    for (int u = 0; u < NS.n_until; u++)
        if (rec[4 + M.aux[NS.until_off + u]] == 1) nexp |= 1 << u;

    uint32_t h = 0;
    for (int j = lane; j < KW; j += 32) h ^= key_word_hash(key_word(M, rec, ncid, nexp, j), j);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, s);
    h = mix32(h);
    __syncwarp();
    if (lane == 0) {
        rec[1] = ncid;
        rec[2] = nexp;
        rec[3] = (int32_t)h;
        if (M.world > 1) atomicAdd(&P.counters[C_OWNER0 + leaf_owner(M, ncid, h)], 1ull);      // (who gets how many: sharded searches only)
    }
    __syncwarp();
    return true;
}

__device__ __forceinline__ void route_body(const DevModel &M, const RouteArgs &P) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n = P.list ? P.count : (long long)P.counters[C_LEAVES];
    for (long long it = warp_id; it < n; it += total_warps) {
        const long long li = P.list ? (long long)P.list[it] : it;
        route_leaf(M, P, P.leaves + li * M.rec_words, li, lane);
    }
}


// ---- scatter: group routed leaves by owner rank (multi-GPU only) ---------------------------------------
__global__ void __launch_bounds__(256) scatter_kernel(const DevModel M, const int32_t *leaves, long long n,
                                                      const long long *offsets, unsigned long long *fill,
                                                      int32_t *outbox) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int RW = M.rec_words;
    // 32 leaves per warp and turn: lanes with the same owner share one cursor update, then the warp copies row by row
    for (long long base = warp_id * 32; base < n; base += total_warps * 32) {
        const long long li = base + lane;
        const bool have = li < n;
        const int owner = have ? leaf_owner(M, leaves[li * RW + 1], (uint32_t)leaves[li * RW + 3]) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        const int leader = __ffs(peers) - 1;
        unsigned long long pos = 0;
        if (have && lane == leader) pos = atomicAdd(&fill[owner], (unsigned long long)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader) + (unsigned long long)__popc(peers & ((1u << lane) - 1u));
        const long long row = have ? offsets[owner] + (long long)pos : -1;
        const int cnt = (int)min(32ll, n - base);
        for (int t = 0; t < cnt; t++) {
            const long long r = __shfl_sync(0xffffffffu, row, t);
            const int32_t *rec = leaves + (base + t) * RW;
            int32_t *dst = outbox + r * RW;
            for (int w = lane; w < RW; w += 32) dst[w] = rec[w];
        }
    }
}

__global__ void __launch_bounds__(256) gather_kernel(const int32_t *src, const int32_t *list, long long count,
                                                     int rec_words, int32_t *dst) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp_id; i < count; i += total_warps) {
        const int32_t *rec = src + (long long)list[i] * rec_words;
        for (int w = lane; w < rec_words; w += 32) dst[i * rec_words + w] = rec[w];
    }
}

// ---- ingest: state dedup, edge append, first search node of every new state ------------------------------
// One warp per routed leaf record (reference vertexTableGetVertex/AddVertex src/graph.cpp:108-123,
// edgeNew src/graph.cpp:78-89, variableAdvanceOneTimeStep src/variable.cpp:94-108).
// Merge one routed leaf record into the automaton (all lanes of the warp).
__device__ __forceinline__ void ingest_leaf(const DevModel &M, const IngestArgs &P, const int32_t *rec, int lane,
                                            unsigned long long &st_dom) {
    const int V = M.V, k = M.k, KW = M.key_words, NW = M.node_words;
    const int src = rec[0], ncid = rec[1], nexp = rec[2];
    const uint32_t h = (uint32_t)rec[3];
    const DevSet NS = M.sets[ncid];

    long long slot = (long long)h & P.table_mask;
    int dst = -1;                                     // local index of the destination state
    bool is_new = false, abort_leaf = false;
    for (;;) {
        int v = 0;
        if (lane == 0) v = atomicCAS(&P.table[slot], -1, -2);
        v = __shfl_sync(0xffffffffu, v, 0);
        if (v == -1) {                                // slot claimed: this leaf creates the state
            unsigned long long id = 0;
            if (lane == 0) id = atomicAdd(&P.totals[C_STATES], 1ull);
            id = __shfl_sync(0xffffffffu, id, 0);
            if ((long long)id >= P.state_cap) {
                if (lane == 0) { atomicOr(&P.counters[C_OVERFLOW], 4ull); atomicExch(&P.table[slot], -3); }
                abort_leaf = true;
                break;
            }
            for (int j = lane; j < KW; j += 32) P.state_key[id * KW + j] = key_word(M, rec, ncid, nexp, j);
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(&P.table[slot], (int)id);
            dst = (int)id;
            is_new = true;
            break;
        }
        if (v == -2) {                                // another warp is writing this slot's key
            if (lane == 0) {
                volatile int32_t *vs = P.table + slot;
                do { v = *vs; } while (v == -2);
            }
            v = __shfl_sync(0xffffffffu, v, 0);
        }
        if (v == -3) { abort_leaf = true; break; }
        __threadfence();
        bool eq = true;
        for (int j = lane; j < KW; j += 32)
            eq &= __ldcg(&P.state_key[(long long)v * KW + j]) == key_word(M, rec, ncid, nexp, j);
        if (__all_sync(0xffffffffu, eq)) { dst = v; break; }
        slot = (slot + 1) & P.table_mask;
    }
    if (abort_leaf) return;
    const int dst_global = dst * M.world + M.rank;

    // both cursors at once: two atomics in flight instead of two round trips one after the other
    unsigned long long o = 0, e = 0;
    if (lane == 0) {
        if (is_new) o = (unsigned long long)P.out_base + atomicAdd(&P.counters[P.fused ? C_OUT : C_NEW], 1ull);
        e = atomicAdd(&P.totals[C_EDGES], 1ull);
    }
    o = __shfl_sync(0xffffffffu, o, 0);
    e = __shfl_sync(0xffffffffu, e, 0);
    if (is_new) {
        // first search node of the new state: next-linked variables take the values just chosen,
        // everything else restarts from its declared range
        if ((long long)o >= P.out_cap) {
            if (lane == 0) atomicOr(&P.counters[C_OVERFLOW], 32ull);
        } else {
            int32_t *node = P.out_nodes + o * NW;
            if (lane == 0) { node[0] = dst_global; node[1] = ncid; node[2] = nexp; node[3] = -1; }
            u64 *nd = reinterpret_cast<u64 *>(node + 4);
            for (int i = lane; i < V * k; i += 32) {
                const int v = i / k, p = i % k;
                u64 m = width_mask(M.width[v]);
                if (p == 0 && k > 1) {                // with k == 1 the reference never links time points
                    for (int t = 0; t < NS.n_next; t++) {
                        if (M.aux[NS.next_off + 2 * t + 1] != v) continue;
                        const long long b = (long long)rec[4 + M.aux[NS.next_off + 2 * t]] - (long long)M.lb[v];
                        m &= (b >= 0 && b < 64) ? (1ull << b) : 0ull;
                    }
                }
                nd[i] = m;
            }
        }
    } else {
        st_dom++;
    }
    if ((long long)e >= P.edge_cap) {
        if (lane == 0) atomicOr(&P.counters[C_OVERFLOW], 8ull);
        return;
    }
    if (lane == 0) {
        P.edge_src[e] = src;
        P.edge_dst[e] = dst_global;
        if (P.deg != nullptr && src < P.deg_cap) atomicAdd(&P.deg[src], 1);
    }
    for (int v = lane; v < V; v += 32) P.edge_label[e * V + v] = rec[4 + v];
}

// PULL mode: where record `it` of the concatenated segments lives (peer memory of the rank that produced it)
__device__ __forceinline__ const int32_t *pull_source(const IngestArgs &P, long long it, int RW) {
    int q = 0;
    while (q + 1 < P.n_segs && it >= P.seg_count[q]) { it -= P.seg_count[q]; q++; }
    return P.seg_base[q] + it * RW;
}

// pull_smem (PULL mode): one record slot per warp, the record is copied out of the producer's outbox once
__device__ __forceinline__ void ingest_body(const DevModel &M, const IngestArgs &P, int32_t *pull_smem) {
    const int lane = threadIdx.x & 31, RW = M.rec_words;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long st_dom = 0;
    for (long long it = warp_id; it < P.count; it += total_warps) {
        const int32_t *rec = P.records + it * RW;
        if (P.n_segs > 0) {
            const int32_t *src = pull_source(P, it, RW);
            int32_t *slot = pull_smem + (threadIdx.x >> 5) * 4 * RW;
            __syncwarp();
            for (int w = lane; w < RW; w += 32) slot[w] = __ldcv(src + w);     // peer memory: never from a stale cache line
            __syncwarp();
            rec = slot;
        }
        ingest_leaf(M, P, rec, lane, st_dom);
    }
    if (lane == 0 && st_dom) {
        if (P.block_stats != nullptr) atomicAdd(&P.block_stats[BS_DOM], (unsigned)st_dom);      // (flushed by the block, see search_kernel)
        else atomicAdd(&P.counters[C_DOMINANCE], st_dom);
    }
}

// Route and merge in one pass (single rank): what cannot be routed yet is left for the host.
__device__ __forceinline__ void leaf_body(const DevModel &M, const RouteArgs &R, const IngestArgs &P, long long n_leaves) {
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long st_dom = 0;
    for (long long li = warp_id; li < n_leaves; li += total_warps) {
        int32_t *rec = R.leaves + li * M.rec_words;
        if (route_leaf(M, R, rec, li, lane)) ingest_leaf(M, P, rec, lane, st_dom);
    }
    if (lane == 0 && st_dom) {
        if (P.block_stats != nullptr) atomicAdd(&P.block_stats[BS_DOM], (unsigned)st_dom);      // (flushed by the block, see search_kernel)
        else atomicAdd(&P.counters[C_DOMINANCE], st_dom);
    }
}


__device__ __noinline__ void fused_leaf(const DevModel &M, const RouteArgs &R, const IngestArgs &I, int32_t *rec, long long li,
                                        int lane, unsigned long long *st_dom) {
    if (route_leaf(M, R, rec, li, lane)) ingest_leaf(M, I, rec, lane, *st_dom);
}

// ---- wide waves: route + merge FOUR leaves per warp, eight lanes each ---------------------------------------------
// One leaf is a chain of dependent global round trips (table slot, state key, edge cursor); with one leaf per warp the
// chain's latency is all there is.  Four independent chains per warp overlap it.  Same results as leaf_body.
// ROUTE without INGEST is the producing rank's half (multi-GPU), INGEST without ROUTE the owner's half over its inbox.
template <bool ROUTE, bool INGEST>
__device__ __forceinline__ void leaf_body_quad(const DevModel &M, const RouteArgs &R, const IngestArgs &P, long long n_leaves,
                                               int32_t *pull_smem = nullptr) {
    const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7, lead = g * 8;
    const unsigned gmask = 0xffu << lead;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int V = M.V, k = M.k, KW = M.key_words, NW = M.node_words;
    unsigned long long st_dom = 0;
    for (long long l4 = warp_id * 4; l4 < n_leaves; l4 += total_warps * 4) {
        const long long it = l4 + g;
        const bool have = it < n_leaves;
        long long li = have ? it : 0;
        if (ROUTE && R.list) li = R.list[li];
        int32_t *rec = ROUTE ? R.leaves + li * M.rec_words : const_cast<int32_t *>(P.records) + li * M.rec_words;
        if (!ROUTE && P.n_segs > 0) {
            // PULL mode: the eight lanes of the group copy their record out of the producer's outbox (peer memory, NVLink)
            int32_t *slot = pull_smem + ((threadIdx.x >> 5) * 4 + g) * M.rec_words;
            __syncwarp();
            if (have) {
                const int32_t *src = pull_source(P, li, M.rec_words);
                for (int w = gl; w < M.rec_words; w += 8) slot[w] = __ldcv(src + w);
            }
            __syncwarp();
            rec = slot;
        }
        // ---- route (reference src/solveralgorithm.cpp:755-837)
        int ncid = -1, nexp = 0;
        bool routed = false;
        uint32_t h = 0;
        if (!ROUTE) {
            if (have) { ncid = rec[1]; nexp = rec[2]; h = (uint32_t)rec[3]; }
            routed = have;
        } else {
        if (have) {
            const int cid = rec[1];
            nexp = rec[2];
            const DevSet S = M.sets[cid];
            ncid = S.static_next;
            if (ncid < 0) {
                if (gl == 0) {
                    if (R.capmap_mask >= 0) {
                        const int32_t *cap = M.aux + S.cap_off;
                        uint32_t ch = cap_hash_begin(cid);
                        for (int i = 0; i < S.n_cap; i++) ch = cap_hash_step(ch, rec[4 + cap[i]]);
                        ch = cap_hash_end(ch) & (uint32_t)R.capmap_mask;
                        for (;;) {
                            const CapEntry e = R.capmap[ch];
                            if (e.cid == -1) break;
                            bool eq = e.cid == cid;
                            for (int i = 0; eq && i < S.n_cap; i++) eq = R.capvals[e.off + i] == rec[4 + cap[i]];
                            if (eq) { ncid = e.next; break; }
                            ch = (ch + 1) & (uint32_t)R.capmap_mask;
                        }
                    }
                    if (ncid < 0) {
                        const unsigned long long u = atomicAdd(&R.counters[C_UNRESOLVED], 1ull);
                        if ((long long)u < R.unresolved_cap) R.unresolved[u] = (int32_t)li;
                        else atomicOr(&R.counters[C_OVERFLOW], 16ull);
                    }
                }
            }
        }
        ncid = __shfl_sync(0xffffffffu, ncid, lead);
        routed = have && ncid >= 0;
        if (routed) {
            const DevSet NS = M.sets[ncid];
            for (int u = 0; u < NS.n_until; u++)
                if (rec[4 + M.aux[NS.until_off + u]] == 1) nexp |= 1 << u;
            for (int j = gl; j < KW; j += 8) h ^= key_word_hash(key_word(M, rec, ncid, nexp, j), j);
        }
        h ^= __shfl_xor_sync(0xffffffffu, h, 4);
        h ^= __shfl_xor_sync(0xffffffffu, h, 2);
        h ^= __shfl_xor_sync(0xffffffffu, h, 1);
        h = mix32(h);
        if (routed && gl == 0) {
            rec[1] = ncid;
            rec[2] = nexp;
            rec[3] = (int32_t)h;
        }
        if (M.world <= 1) {   // one counter update per warp, not per leaf: these are single-address atomics
            const unsigned rb = __ballot_sync(0xffffffffu, routed && gl == 0);
            if (lane == 0 && rb) atomicAdd(&R.counters[C_OWNER0], (unsigned long long)__popc(rb));
        } else if (routed && gl == 0) {
            atomicAdd(&R.counters[C_OWNER0 + leaf_owner(M, ncid, h)], 1ull);
        }
        }
        if (!INGEST) continue;
        __syncwarp();
        // ---- find or insert the successor state (reference vertexTableGetVertex / AddVertex)
        long long slot = (long long)h & P.table_mask;
        int dst = -1;
        bool is_new = false, searching = routed;
        while (__any_sync(0xffffffffu, searching)) {
            int v = 0;
            if (searching && gl == 0) v = atomicCAS(&P.table[slot], -1, -2);
            v = __shfl_sync(0xffffffffu, v, lead);
            if (searching && v == -1) {                     // slot claimed: this leaf creates the state
                unsigned long long id = 0;
                if (gl == 0) id = atomicAdd(&P.totals[C_STATES], 1ull);
                id = __shfl_sync(gmask, id, lead);
                if ((long long)id >= P.state_cap) {
                    if (gl == 0) { atomicOr(&P.counters[C_OVERFLOW], 4ull); atomicExch(&P.table[slot], -3); }
                    searching = false;
                    routed = false;
                } else {
                    for (int j = gl; j < KW; j += 8) P.state_key[id * KW + j] = key_word(M, rec, ncid, nexp, j);
                    __threadfence();
                    __syncwarp(gmask);
                    if (gl == 0) atomicExch(&P.table[slot], (int)id);
                    dst = (int)id;
                    is_new = true;
                    searching = false;
                }
            } else if (searching && v == -3) {
                searching = false;
                routed = false;
            } else if (searching && v >= 0) {
                __threadfence();
                bool eq = true;
                for (int j = gl; j < KW; j += 8)
                    eq &= __ldcg(&P.state_key[(long long)v * KW + j]) == key_word(M, rec, ncid, nexp, j);
                const unsigned be = __ballot_sync(gmask, eq);
                if ((be & gmask) == gmask) { dst = v; searching = false; }
                else slot = (slot + 1) & P.table_mask;
            }
            // v == -2: another group or warp is writing this slot's key -- look again in the next turn
        }
        // ---- first search node of a new state, and the edge
        unsigned long long o = 0, e = 0;
        {
            const unsigned rb = __ballot_sync(0xffffffffu, routed && gl == 0);
            const unsigned nb = __ballot_sync(0xffffffffu, routed && is_new && gl == 0);
            if (lane == 0) {
                if (rb) e = atomicAdd(&P.totals[C_EDGES], (unsigned long long)__popc(rb));
                if (nb) o = (unsigned long long)P.out_base + atomicAdd(&P.counters[C_NEW], (unsigned long long)__popc(nb));
            }
            const unsigned below = (1u << lead) - 1u;
            e = __shfl_sync(0xffffffffu, e, 0) + (unsigned long long)__popc(rb & below);
            o = __shfl_sync(0xffffffffu, o, 0) + (unsigned long long)__popc(nb & below);
        }
        if (routed) {
            const int dst_global = dst * M.world + M.rank;
            if (is_new) {
                if ((long long)o >= P.out_cap) {
                    if (gl == 0) atomicOr(&P.counters[C_OVERFLOW], 32ull);
                } else {
                    const DevSet NS = M.sets[ncid];
                    int32_t *node = P.out_nodes + o * NW;
                    if (gl == 0) { node[0] = dst_global; node[1] = ncid; node[2] = nexp; node[3] = -1; }
                    u64 *nd = reinterpret_cast<u64 *>(node + 4);
                    for (int i = gl; i < V * k; i += 8) {
                        const int v = i / k, p = i % k;
                        u64 m = width_mask(M.width[v]);
                        if (p == 0 && k > 1) {
                            for (int t = 0; t < NS.n_next; t++) {
                                if (M.aux[NS.next_off + 2 * t + 1] != v) continue;
                                const long long b = (long long)rec[4 + M.aux[NS.next_off + 2 * t]] - (long long)M.lb[v];
                                m &= (b >= 0 && b < 64) ? (1ull << b) : 0ull;
                            }
                        }
                        nd[i] = m;
                    }
                }
            } else if (gl == 0) {
                st_dom++;
            }
            if ((long long)e >= P.edge_cap) {
                if (gl == 0) atomicOr(&P.counters[C_OVERFLOW], 8ull);
            } else {
                if (gl == 0) {
                    P.edge_src[e] = rec[0];
                    P.edge_dst[e] = dst_global;
                    if (P.deg != nullptr && rec[0] < P.deg_cap) atomicAdd(&P.deg[rec[0]], 1);
                }
                for (int v = gl; v < V; v += 8) P.edge_label[e * V + v] = rec[4 + v];
            }
        }
    }
    st_dom = (unsigned long long)__reduce_add_sync(0xffffffffu, (unsigned)st_dom);
    if (lane == 0 && st_dom) {
        if (P.block_stats != nullptr) atomicAdd(&P.block_stats[BS_DOM], (unsigned)st_dom);      // (flushed by the block, see search_kernel)
        else atomicAdd(&P.counters[C_DOMINANCE], st_dom);
    }
}

// Step-wise launches (multi-GPU sessions, profile_kernels): wide lists take the four-per-warp path too.
__device__ __forceinline__ bool wide_leaf_list(const DevModel &M, long long n) {
    return n >= 4ll * ((long long)gridDim.x * blockDim.x >> 5) || M.force_mode == EXPAND_QUAD + 1;
}

__global__ void __launch_bounds__(256) route_kernel(const DevModel M, const RouteArgs P) {
    const long long n = P.list ? P.count : (long long)P.counters[C_LEAVES];
    if (wide_leaf_list(M, n)) leaf_body_quad<true, false>(M, P, IngestArgs{}, n);
    else route_body(M, P);
}

__global__ void __launch_bounds__(256) ingest_kernel(const DevModel M, const IngestArgs P) {
    extern __shared__ __align__(16) unsigned char smem[];      // PULL mode: 32 record slots (four per warp)
    int32_t *pull_smem = reinterpret_cast<int32_t *>(smem);
    if (wide_leaf_list(M, P.count)) leaf_body_quad<false, true>(M, RouteArgs{}, P, P.count, pull_smem);
    else ingest_body(M, P, pull_smem);
}

// ---- the whole wave loop in one cooperative launch -------------------------------------------------------
// Waves of a small or deep search cost more in launches and host synchronisation than in work.  search_kernel
// runs expand -> route + ingest -> bookkeeping for wave after wave and returns to the host only when it is done or
// needs it (buffers to grow, an unseen constraint-set transition, a wave wide enough for stand-alone launches).
//
// Grid barriers are what a narrow wave pays for, so there are as few as possible: ONE after expand, and one more after
// the leaf phase only if the wave produced leaves.  There is no controller block: every block keeps the wave state in
// its own shared memory and takes every decision itself, from counters that are frozen while they are read, so all
// blocks decide alike without another barrier.  That needs the wave counters THREE times over (wave w uses set w % 3):
// block 0 clears the set of wave w-1 after the barrier of wave w -- every block has finished reading it by then --
// and no block touches that set again before it has passed the barrier of wave w+1, which block 0 reaches only after
// the clearing.  C_STATES and C_EDGES (never reset) live in set 0.
// Two instantiations: CTAS = kExpandCtasPerSm for the full grid (80 registers per thread, the inlined expand bodies spill
// a little), CTAS = 1 for the narrow grid of one CTA per SM, where ptxas may use 255 registers and nothing spills.
template <int CTAS>
__global__ void __launch_bounds__(kExpandWarps * 32, CTAS) search_kernel(const DevModel M, SearchArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    volatile SearchCtl *ctl = A.ctl;
    volatile unsigned long long *tot = A.counters;
    // Wave state and running totals of THIS block (thread 0), in shared memory: nothing of it may stay live in
    // registers across the expand bodies, whose inner loops need every register the launch bounds allow.
    enum { S_N_IN, S_STATES, S_EDGES, S_WAVES_LEFT, S_NODES, S_FAILS, S_TUPLES, S_REV, S_DOM, S_LEAVES, S_WAVES, S_CUR,
           S_OVERFLOW, S_WAVE, S_MAX_IN, S_COUNT };
    __shared__ long long cs[S_COUNT];
    __shared__ ExpandArgs ea;           // the wave's arguments (launch parameters in the stand-alone kernels)
    __shared__ int s_status, s_set;
    __shared__ int s_fuse;              // this wave's leaves are routed and merged inside expand (no leaf phase)
    __shared__ RouteArgs s_ra;          // ... with these arguments
    __shared__ IngestArgs s_ia;
    __shared__ int s_resident;          // constraint set whose metadata this CTA holds in shared memory (kept across waves)
    __shared__ long long s_an, s_af;    // look-ahead totals as of the last wave's end (C_AHEAD_NODES / C_AHEAD_FAILS)
    __shared__ unsigned s_bstats[BS_COUNT];     // this block's statistics of the running expand pass (flush_warp_stats)
    const bool controller = blockIdx.x == 0 && threadIdx.x == 0;    // the one thread that reports to the host
    auto stamp = [&](int k) {
        if (A.trace != nullptr && controller && cs[S_WAVE] < A.trace_cap) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.trace[cs[S_WAVE] * 5 + k] = t;
        }
    };
    if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 90);     // kernel entered
    if (A.make_root && blockIdx.x == 0) {
        // The root state (signature "S", reference src/solveralgorithm.cpp:951-954) and its search node -- every variable at
        // its declared range at every offset -- written here instead of by three host-to-device copies.  Wave 0 is this one
        // node, and block 0 is the block that expands it (after the barrier at the top of the wave loop).
        for (int j = threadIdx.x; j < M.key_words; j += blockDim.x) A.state_key[j] = (j == 0 && M.sig_len != 0) ? -1 : 0;
        int32_t *node = A.frontier[A.cur0];
        if (threadIdx.x < 4) node[threadIdx.x] = threadIdx.x == 3 ? -1 : 0;
        u64 *nd = reinterpret_cast<u64 *>(node + 4);
        for (int i = threadIdx.x; i < M.V * M.k; i += blockDim.x) nd[i] = width_mask(M.width[i / M.k]);
        if (threadIdx.x == 0) {
            A.counters[C_STATES] = 1ull;
            // ... and its slot in the (fresh) state table, like rehash_kernel does for every known state
            uint32_t h = 0;
            for (int j = 0; j < M.key_words; j++) h ^= key_word_hash((j == 0 && M.sig_len != 0) ? -1 : 0, j);
            A.table[(long long)mix32(h) & A.table_mask] = 0;
        }
        __threadfence();
    }
    if (A.make_root && A.fin.deg != nullptr && A.fin.deg_counted) {
        // out-degrees are counted while the edges are appended (IngestArgs::deg): start from zero.  No edge is appended
        // before the grid barrier of wave 0 (its leaves, if the root is one, are not fused -- see `fuse` below).
        const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long long)gridDim.x * blockDim.x;
        for (long long i = gtid; i <= A.fin.cap_states; i += gsize) { A.fin.deg[i] = 0; A.fin.cursor[i] = 0; }
    }
    if (threadIdx.x < BS_COUNT) s_bstats[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        s_resident = -1;
        for (int i = 0; i < S_COUNT; i++) cs[i] = 0;
        cs[S_N_IN] = A.n_in0;
        cs[S_CUR] = A.cur0;
        cs[S_WAVES_LEFT] = A.waves_left0;
        cs[S_STATES] = A.make_root ? 1ll : (long long)tot[C_STATES];
        cs[S_EDGES] = (long long)tot[C_EDGES];
        s_an = A.make_root || A.ahead_policy != 1 ? 0ll : (long long)tot[C_AHEAD_NODES];
        s_af = A.make_root || A.ahead_policy != 1 ? 0ll : (long long)tot[C_AHEAD_FAILS];
    }
    // A solve starts with ONE node on one block.  The other blocks use the time to stage the root's constraint set (it is every
    // node's set in most models), so that their first node does not wait for that.  (Pulling small relation tables into every
    // SM's L1 here as well was measured: no effect.)
    if (A.make_root && blockIdx.x != 0) {
        __syncthreads();                        // s_resident = -1 is in place
        stage_set(M, smem, 0, &s_resident);
    }
    // Leave the kernel (all threads call it, every block with the same status).  The host reads set 0.
    auto leave = [&](int st, bool mid_wave) {
        if (blockIdx.x != 0) return;
        if (mid_wave && s_set != 0 && threadIdx.x >= C_OUT && threadIdx.x < C_COUNT)
            A.counters[threadIdx.x] = A.counters[s_set * kCounterStride + threadIdx.x];
        // host mirror of counter set 0 (the threads that may just have copied a word read that word back: program order)
        if (A.h_counters != nullptr && threadIdx.x < C_COUNT) A.h_counters[threadIdx.x] = tot[threadIdx.x];
        if (A.h_counters != nullptr && threadIdx.x >= C_AHEAD_NODES && threadIdx.x <= C_AHEAD_FAILS)
            A.h_counters[kHostAheadWord + threadIdx.x - C_AHEAD_NODES] = tot[threadIdx.x];
        if (threadIdx.x == 0) {
            SearchCtl c;
            c.status = st;
            c.n_in = cs[S_N_IN];
            c.cur = (int)cs[S_CUR];
            c.t_nodes = cs[S_NODES]; c.t_fails = cs[S_FAILS]; c.t_tuples = cs[S_TUPLES]; c.t_revisions = cs[S_REV];
            c.t_dominance = cs[S_DOM]; c.t_leaves = cs[S_LEAVES]; c.t_waves = cs[S_WAVES];
            c.waves_left = cs[S_WAVES_LEFT];
            c.overflow = (int)cs[S_OVERFLOW];
            c.t_max_in = cs[S_MAX_IN];
            c.finished = 0; c.dead_edges = 0; c.changed = 0; c.pushed = 0; c.pad = 0;
            *const_cast<SearchCtl *>(ctl) = c;
            if (A.h_ctl != nullptr) *A.h_ctl = c;
        }
    };
    // Warp 0: one parallel load of the wave's counters (frozen while it runs).  after_expand: publish expand's results
    // for the block's decisions, and if the wave has no leaves finish it right away; else: end of a wave with leaves.
    __shared__ long long s_leaves, s_out;
    __shared__ int s_ovf;
    auto wave_counters = [&](volatile unsigned long long *cnt, bool after_expand) {
        const int ln = threadIdx.x;
        // Every block reads the same few words at the same moment -- right after the grid barrier -- and an L2 slice serves
        // requests for one line one after the other: with a 64-bit load per counter and lane (30 per block, 4 440 in all) a
        // block that had waited at the barrier got its values 3.7 us later (the block that arrives last reads before the
        // others have noticed the release: 0.5 us).  So: 16-byte loads, only of what the block's decisions need (five
        // requests); the statistics are the controller block's business alone (two more).
        static_assert(C_STATES == 0 && C_EDGES == 1 && C_OUT == 2 && C_NEW == 3 && C_LEAVES == 4 && C_UNRESOLVED == 5 &&
                      C_OVERFLOW == 6 && C_NODES == 7 && C_FAILS == 8 && C_TUPLES == 9 && C_REVISIONS == 10 && C_DOMINANCE == 11 &&
                      C_AHEAD_NODES == 28 && C_AHEAD_FAILS == 29, "pairs of counters are loaded together");
        const ulonglong2 *pt = reinterpret_cast<const ulonglong2 *>(const_cast<const unsigned long long *>(tot));
        const ulonglong2 *pc = reinterpret_cast<const ulonglong2 *>(const_cast<const unsigned long long *>(cnt));
        ulonglong2 c = make_ulonglong2(0ull, 0ull);
        if (ln == 0) c = __ldcg(pt);                                    // C_STATES, C_EDGES
        else if (ln <= 3) c = __ldcg(pc + ln);                          // (C_OUT, C_NEW) (C_LEAVES, C_UNRESOLVED) (C_OVERFLOW, C_NODES)
        else if (ln == 4 && A.ahead_policy == 1) c = __ldcg(pt + 14);   // C_AHEAD_NODES, C_AHEAD_FAILS
        else if ((ln == 5 || ln == 6) && blockIdx.x == 0) c = __ldcg(pc + ln - 1);     // (C_FAILS, C_TUPLES) (C_REVISIONS, C_DOMINANCE)
        const long long v_states = (long long)__shfl_sync(0xffffffffu, c.x, 0);
        const long long v_edges = (long long)__shfl_sync(0xffffffffu, c.y, 0);
        const long long v_out = (long long)__shfl_sync(0xffffffffu, c.x, 1);
        const long long v_new = (long long)__shfl_sync(0xffffffffu, c.y, 1);
        const long long v_leaves = (long long)__shfl_sync(0xffffffffu, c.x, 2);
        const long long v_unres = (long long)__shfl_sync(0xffffffffu, c.y, 2);
        const long long v_ovf = (long long)__shfl_sync(0xffffffffu, c.x, 3);
        const long long v_nodes = (long long)__shfl_sync(0xffffffffu, c.y, 3);
        const long long v_an = (long long)__shfl_sync(0xffffffffu, c.x, 4);
        const long long v_af = (long long)__shfl_sync(0xffffffffu, c.y, 4);
        const long long v_fails = (long long)__shfl_sync(0xffffffffu, c.x, 5);     // (zero in every block but the controller's,
        const long long v_tuples = (long long)__shfl_sync(0xffffffffu, c.y, 5);    //  which alone reports the statistics: leave())
        const long long v_rev = (long long)__shfl_sync(0xffffffffu, c.x, 6);
        const long long v_dom = (long long)__shfl_sync(0xffffffffu, c.y, 6);
        if (ln != 0) return;
        if (after_expand) { s_an = v_an; s_af = v_af; }
        if (after_expand) {
            // (a fused wave has merged its leaves already; its frontier cannot overflow -- see the guard at wave start -- and if
            //  it did, re-running the wave would duplicate edges: the host is told that a pool overflowed)
            const long long pend = s_fuse ? 0 : v_leaves;
            if (s_fuse && (v_ovf & 1)) cs[S_OVERFLOW] |= 64;
            s_ovf = (int)(v_ovf & 1);
            s_leaves = pend;
            s_out = v_out;
            if ((v_ovf & 1) || v_out + pend > A.out_cap || pend > 0) return;    // leaving, or the leaf phase comes first
        }
        s_status = v_unres != 0 ? SEARCH_RESOLVE : SEARCH_RUN;
        if (v_unres != 0) return;               // the host finishes this wave and counts it
        cs[S_NODES] += v_nodes; cs[S_FAILS] += v_fails; cs[S_TUPLES] += v_tuples; cs[S_REV] += v_rev;
        cs[S_DOM] += v_dom;
        cs[S_LEAVES] += v_leaves;
        cs[S_WAVES] += 1;
        cs[S_OVERFLOW] |= v_ovf;
        cs[S_WAVES_LEFT] -= 1;
        cs[S_N_IN] = v_out + v_new;
        cs[S_CUR] ^= 1;
        cs[S_STATES] = v_states;
        cs[S_EDGES] = v_edges;
    };
    auto flush_block_stats = [&]() {
        __syncthreads();
        if (threadIdx.x < BS_COUNT && s_bstats[threadIdx.x] != 0u) {
            const int t = threadIdx.x;
            unsigned long long *dst = t == BS_AHEAD_NODES ? &A.counters[C_AHEAD_NODES] : t == BS_AHEAD_FAILS ? &A.counters[C_AHEAD_FAILS]
                                    : &ea.counters[t == BS_NODES ? C_NODES : t == BS_FAILS ? C_FAILS : t == BS_TUPLES ? C_TUPLES
                                                   : t == BS_REV ? C_REVISIONS : C_DOMINANCE];
            atomicAdd(dst, (unsigned long long)s_bstats[t]);
            s_bstats[t] = 0u;
        }
    };
    for (;;) {
        // ---- wave start: can this wave run without the host?
        if (threadIdx.x == 0) {
            const long long c_n_in = cs[S_N_IN], c_states = cs[S_STATES], c_edges = cs[S_EDGES];
            int st = SEARCH_RUN;
            if (c_n_in == 0) st = SEARCH_DONE;
            else if (cs[S_WAVES_LEFT] <= 0 || (A.max_frontier > 0 && c_n_in > A.max_frontier)) st = SEARCH_YIELD;
            else if (c_n_in > A.leaf_cap || c_n_in > A.unresolved_cap || c_states + c_n_in > A.state_cap ||
                     c_edges + c_n_in > A.edge_cap || 2 * (c_states + c_n_in) > A.table_mask + 1 || 2 * c_n_in > A.out_cap)
                st = SEARCH_GROW;
            s_status = st;
            if (st == SEARCH_RUN && c_n_in > cs[S_MAX_IN]) cs[S_MAX_IN] = c_n_in;
            const int set = (int)(cs[S_WAVE] % 3), cur = (int)cs[S_CUR];
            s_set = set;
            ea.in_nodes = A.frontier[cur];
            ea.n_in = c_n_in;
            ea.out_nodes = A.frontier[cur ^ 1];
            ea.out_cap = A.out_cap;
            ea.leaves = A.leaves;
            ea.leaf_cap = A.leaf_cap;
            ea.fan = branch_fan(M, c_n_in, A.out_cap);
            ea.counters = A.counters + set * kCounterStride;
            ea.dbg = A.trace ? A.trace + 5 * A.trace_cap : nullptr;     // block 0's timeline follows the per-wave stamps
            ea.dbg_cap = A.trace ? 4096 : 0;
            // Look-ahead propagators: dropped while kAheadSample or more nodes have run them, fewer than one in kAheadRatio of
            // those failed because of one and kAheadLeaves leaves have been seen (ahead_droppable) -- but for one WAVE in
            // kAheadWavePeriod, which runs them in all its nodes and keeps the two totals alive.  (Sampling single nodes of a
            // wave does not pay: a narrow wave lasts as long as its slowest node.)  The blocks agree (the totals are only
            // written during expand, and this is read between two grid barriers), though nothing depends on that.
            ea.ahead_stats = A.ahead_policy == 1 ? A.counters : nullptr;
            ea.block_stats = s_bstats;
            ea.skip_ahead = A.ahead_policy == 2 ||
                            (A.ahead_policy == 1 && !ahead_sample_wave(A.waves0 + cs[S_WAVES]) &&
                             ahead_droppable(s_an, s_af, A.leaves0 + cs[S_LEAVES]));
            // Narrow waves (a CTA or a warp per node): the warp that finds a leaf routes and merges it at once, while the
            // other nodes of the wave are still being propagated -- one grid barrier per wave, and the leaf's chain of
            // dependent L2 round trips is hidden behind the slowest node.  Only when the wave cannot overflow the output
            // frontier (a node makes at most max(64, fan) children or one new state; a fused wave cannot be run again) and
            // never on the quad-mode waves, whose leaf phase overlaps four chains per warp anyway.
            const int wmode = pick_expand_mode(M, c_n_in, gridDim.x);
            const long long per_node = ea.fan > 64 ? ea.fan : 64;
            const int fuse = st == SEARCH_RUN && A.fuse_leaves && wmode != EXPAND_QUAD && c_n_in * per_node <= A.out_cap &&
                             !(A.make_root && cs[S_WAVES] == 0);
            s_fuse = fuse;
            ea.fuse_route = nullptr;
            ea.fuse_ingest = nullptr;
            if (fuse) {
                s_ra.leaves = A.leaves;
                s_ra.list = nullptr;
                s_ra.count = 0;
                s_ra.capmap = A.capmap;
                s_ra.capvals = A.capvals;
                s_ra.capmap_mask = A.capmap_mask;
                s_ra.unresolved = A.unresolved;
                s_ra.unresolved_cap = A.unresolved_cap;
                s_ra.counters = ea.counters;
                s_ia.records = A.leaves;
                s_ia.count = 0;
                s_ia.table = A.table;
                s_ia.table_mask = A.table_mask;
                s_ia.state_key = A.state_key;
                s_ia.state_cap = A.state_cap;
                s_ia.edge_src = A.edge_src;
                s_ia.edge_dst = A.edge_dst;
                s_ia.edge_label = A.edge_label;
                s_ia.edge_cap = A.edge_cap;
                s_ia.out_nodes = ea.out_nodes;
                s_ia.out_base = 0;
                s_ia.out_cap = A.out_cap;
                s_ia.counters = ea.counters;
                s_ia.totals = A.counters;
                s_ia.n_segs = 0;
                s_ia.fused = 1;
                s_ia.block_stats = nullptr;     // (fused leaves count through expand's own statistics)
                s_ia.deg = A.fin.deg_counted ? A.fin.deg : nullptr;
                s_ia.deg_cap = A.fin.cap_states;
                ea.fuse_route = &s_ra;
                ea.fuse_ingest = &s_ia;
            }
        }
        __syncthreads();
        if (s_status == SEARCH_DONE) {
            if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 91);     // search done, finishing begins
            leave(SEARCH_DONE, false);
            // ---- small automaton: group the edges by source and apply the fail rule right here (no further launches)
            const long long ns = (long long)tot[C_STATES], ne = (long long)tot[C_EDGES];
            const FinishArgs &F = A.fin;
            if (F.deg == nullptr || ns > F.cap_states || ne > F.cap_edges) return;
            const int V = M.V, KW = M.key_words;
            const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long long)gridDim.x * blockDim.x;
            const int lane = threadIdx.x & 31;
            // (this launch's dynamic shared memory = expand_smem_bytes(M); nobody is using it any more)
            const size_t smem_bytes = node_bytes(M) * M.node_slots + scratch_bytes(M) * kExpandWarps + align8(sizeof(DevModel)) +
                                      align8((size_t)M.stage_bytes);
            if (F.deg_counted && (size_t)(ns + 1 + blockDim.x) * 4 <= smem_bytes) {
                // ---- degrees already counted: EVERY block scans them into its own shared memory, so positions are known
                // without a grid barrier; when no state is a dead end (nothing to trim) the edges go straight to their final
                // places, on the device and -- if the host gave pinned buffers -- on the host, and the kernel is done
                int32_t *first_s = reinterpret_cast<int32_t *>(smem);
                int32_t *partial = first_s + ns + 1;
                const int nt = blockDim.x, tid = threadIdx.x;
                const long long n = ns + 1, chunk = (n + nt - 1) / nt;
                const long long lo = min((long long)tid * chunk, n), hi = min(lo + chunk, n);
                int sum = 0;
                bool dead_end = false;
                for (long long i = lo; i < hi; i++) {
                    const int d = i < ns ? F.deg[i] : 0;
                    first_s[i] = d;
                    sum += d;
                    dead_end |= i < ns && d == 0;
                }
                partial[tid] = sum;
                __syncthreads();
                for (int off = 1; off < nt; off <<= 1) {
                    const int v = tid >= off ? partial[tid - off] : 0;
                    __syncthreads();
                    partial[tid] += v;
                    __syncthreads();
                }
                int run = partial[tid] - sum;
                for (long long i = lo; i < hi; i++) { const int d = first_s[i]; first_s[i] = run; run += d; }
                const bool any_dead_end = __syncthreads_or(dead_end && F.do_trim) != 0;     // (the barrier also publishes first_s)
                if (!any_dead_end && first_s[ns] == (int)ne) {
                    const bool push = F.h_src != nullptr && ns <= F.h_cap_states && ne <= F.h_cap_edges;
                    const long long gwarp = gtid >> 5, nwarps = gsize >> 5;
                    for (long long e = gwarp; e < ne; e += nwarps) {
                        const int s = A.edge_src[e];
                        int pos = 0;
                        if (lane == 0) pos = first_s[s] + atomicAdd(&F.cursor[s], 1);
                        pos = __shfl_sync(0xffffffffu, pos, 0);
                        if (lane == 0) {
                            const int d = A.edge_dst[e];
                            F.s_src[pos] = s; F.s_dst[pos] = d; F.alive[pos] = 1;
                            if (push) { F.h_src[pos] = s; F.h_dst[pos] = d; }
                        }
                        for (int v = lane; v < V; v += 32) {
                            const int32_t x = A.edge_label[e * V + v];
                            F.s_label[(long long)pos * V + v] = x;
                            if (push) F.h_label[(long long)pos * V + v] = x;
                        }
                    }
                    for (long long i = gtid; i < ns * KW; i += gsize) {
                        const long long st = i / KW;
                        const int j = (int)(i % KW), v = A.state_key[i];
                        if (j == 0) {
                            F.rows_cset[st] = v < 0 ? 0 : v;
                            F.failed[st] = 0;
                            if (push) { F.h_cset[st] = v < 0 ? 0 : v; F.h_failed[st] = 0; }
                        } else {
                            F.rows_sig[st * (KW - 1) + (j - 1)] = v;
                            if (push) F.h_sig[st * (KW - 1) + (j - 1)] = v;
                        }
                    }
                    if (gtid == 0) {
                        if (A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 92);       // block 0's share of the finishing done
                        ctl->changed = 0;
                        ctl->dead_edges = 0;
                        ctl->finished = 1;
                        ctl->pad = 1;               // (which finishing path ran: shown at verbosity 1)
                        if (A.h_ctl != nullptr) {
                            A.h_ctl->dead_edges = 0;
                            A.h_ctl->finished = 1;
                            A.h_ctl->pushed = push ? 1 : 0;
                            A.h_ctl->pad = 1;
                        }
                    }
                    return;
                }
                grid.sync();            // a dead end (or a count that does not add up): the general path below, from scratch
            }
            for (long long i = gtid; i <= ns; i += gsize) F.deg[i] = 0;
            if (gtid == 0) { ctl->changed = 0; ctl->dead_edges = 0; }
            grid.sync();
            for (long long e = gtid; e < ne; e += gsize) atomicAdd(&F.deg[A.edge_src[e]], 1);
            grid.sync();
            if (blockIdx.x == 0) {              // exclusive scan of deg[0 .. ns] by one CTA
                int32_t *partial = reinterpret_cast<int32_t *>(smem);
                const int nt = blockDim.x, tid = threadIdx.x;
                const long long n = ns + 1, chunk = (n + nt - 1) / nt;
                const long long lo = min((long long)tid * chunk, n), hi = min(lo + chunk, n);
                int sum = 0;
                for (long long i = lo; i < hi; i++) sum += F.deg[i];
                partial[tid] = sum;
                __syncthreads();
                for (int off = 1; off < nt; off <<= 1) {
                    const int v = tid >= off ? partial[tid - off] : 0;
                    __syncthreads();
                    partial[tid] += v;
                    __syncthreads();
                }
                int run = partial[tid] - sum;
                bool any_failed = false;
                for (long long i = lo; i < hi; i++) {
                    const int d = F.deg[i];
                    F.first[i] = run;
                    F.cursor[i] = run;
                    F.outdeg[i] = d;
                    if (i < ns) {
                        F.failed[i] = F.do_trim && d == 0;
                        any_failed |= F.do_trim && d == 0;
                    }
                    run += d;
                }
                if (any_failed) ctl->changed = 1;
            }
            grid.sync();
            {
                const long long gwarp = gtid >> 5, nwarps = gsize >> 5;
                for (long long e = gwarp; e < ne; e += nwarps) {
                    const int s = A.edge_src[e];
                    int pos = 0;
                    if (lane == 0) pos = atomicAdd(&F.cursor[s], 1);
                    pos = __shfl_sync(0xffffffffu, pos, 0);
                    if (lane == 0) { F.s_src[pos] = s; F.s_dst[pos] = A.edge_dst[e]; F.alive[pos] = 1; }
                    for (int v = lane; v < V; v += 32) F.s_label[(long long)pos * V + v] = A.edge_label[e * V + v];
                }
                for (long long i = gtid; i < ns * KW; i += gsize) {
                    const long long s = i / KW;
                    const int j = (int)(i % KW), v = A.state_key[i];
                    if (j == 0) F.rows_cset[s] = v < 0 ? 0 : v;
                    else F.rows_sig[s * (KW - 1) + (j - 1)] = v;
                }
            }
            grid.sync();
            while (ctl->changed) {              // fail rule: uniform over the grid (read between two barriers)
                grid.sync();
                if (gtid == 0) ctl->changed = 0;
                grid.sync();
                for (long long e = gtid; e < ne; e += gsize) {
                    if (!F.alive[e] || !F.failed[F.s_dst[e]]) continue;
                    F.alive[e] = 0;
                    atomicAdd((int *)&ctl->dead_edges, 1);
                    if (atomicSub(&F.outdeg[F.s_src[e]], 1) == 1) {
                        F.failed[F.s_src[e]] = 1;
                        ctl->changed = 1;
                    }
                }
                grid.sync();
            }
            // ---- push: no edge died, so the grouped arrays are final -- write them into the host's pinned buffers
            const bool push = F.h_src != nullptr && ctl->dead_edges == 0 && ns <= F.h_cap_states && ne <= F.h_cap_edges;
            if (push) {
                const long long nsig = ns * (KW - 1);
                for (long long i = gtid; i < ns; i += gsize) { F.h_cset[i] = F.rows_cset[i]; F.h_failed[i] = F.failed[i]; }
                for (long long i = gtid; i < nsig; i += gsize) F.h_sig[i] = F.rows_sig[i];
                for (long long e = gtid; e < ne; e += gsize) { F.h_src[e] = F.s_src[e]; F.h_dst[e] = F.s_dst[e]; }
                for (long long i = gtid; i < ne * V; i += gsize) F.h_label[i] = F.s_label[i];
            }
            if (gtid == 0) {
                ctl->finished = 1;
                if (A.h_ctl != nullptr) {
                    A.h_ctl->dead_edges = ctl->dead_edges;
                    A.h_ctl->finished = 1;
                    A.h_ctl->pushed = push ? 1 : 0;
                }
            }
            return;
        }
        if (s_status != SEARCH_RUN) { leave(s_status, false); return; }
        stamp(0);
        volatile unsigned long long *cnt = ea.counters;         // this wave's counter set
        const int mode = pick_expand_mode(M, ea.n_in, gridDim.x);
        if (mode == EXPAND_CTA) expand_body<true>(M, ea, smem, &s_resident);
        else if (mode == EXPAND_QUAD) expand_body_quad(M, ea, smem, &s_resident);
        else expand_body<false>(M, ea, smem, &s_resident);
        // the block's statistics of this pass: one reduction per counter and BLOCK (flush_warp_stats)
        flush_block_stats();
        if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 96);     // block 0 at the grid barrier
        grid.sync();
        if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 97);     // ... through it
        stamp(1);
        // the previous wave's counter set is no longer read by anybody: clear it for the wave after this one
        // (the same threads that copy the current set for the host when the kernel leaves in mid-wave: program order)
        if (blockIdx.x == 0 && threadIdx.x >= C_OUT && threadIdx.x < C_COUNT)
            A.counters[((s_set + 2) % 3) * kCounterStride + threadIdx.x] = 0ull;
        // Every block must take the same exit decisions, so they may only depend on values that no block is changing
        // while they are read: bit 0 of C_OVERFLOW, C_OUT and C_LEAVES are written by expand only (the leaf phase, which
        // faster blocks may already be in, appends through C_NEW and flags its own overflow bit).
        // Warp 0 reads all counters with one parallel load and publishes what the block needs; on a wave without leaves
        // that load also serves the end-of-wave bookkeeping (nothing changes any more), so such a wave costs one grid
        // barrier and one L2 round trip.
        if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 93);     // counters cleared
        if (threadIdx.x < 32) wave_counters(cnt, true);
        if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 94);     // wave counters read
        __syncthreads();
        if (controller && A.trace != nullptr) dbg_stamp(A.trace + 5 * A.trace_cap, 4096, 95);     // ... and published to the block
        // the frontier buffer overflowed: only scratch was written, the host grows it and the wave runs again
        if (s_ovf) { leave(SEARCH_RETRY, true); return; }
        const long long n_leaves = s_leaves;
        if (s_out + n_leaves > A.out_cap) {
            // no room for the first nodes of new states: the host grows the frontier and finishes the wave (route + ingest)
            leave(SEARCH_INGEST, true);
            return;
        }
        if (n_leaves > 0) {
            RouteArgs ra;
            ra.leaves = A.leaves;
            ra.list = nullptr;
            ra.count = 0;
            ra.capmap = A.capmap;
            ra.capvals = A.capvals;
            ra.capmap_mask = A.capmap_mask;
            ra.unresolved = A.unresolved;
            ra.unresolved_cap = A.unresolved_cap;
            ra.counters = ea.counters;
            IngestArgs ia;
            ia.records = A.leaves;
            ia.count = n_leaves;
            ia.table = A.table;
            ia.table_mask = A.table_mask;
            ia.state_key = A.state_key;
            ia.state_cap = A.state_cap;
            ia.edge_src = A.edge_src;
            ia.edge_dst = A.edge_dst;
            ia.edge_label = A.edge_label;
            ia.edge_cap = A.edge_cap;
            ia.out_nodes = ea.out_nodes;
            ia.out_base = s_out;
            ia.out_cap = A.out_cap;
            ia.counters = ea.counters;
            ia.totals = A.counters;
            ia.n_segs = 0;
            ia.fused = 0;
            ia.block_stats = s_bstats;
            ia.deg = A.fin.deg_counted ? A.fin.deg : nullptr;
            ia.deg_cap = A.fin.cap_states;
            if (n_leaves >= 4ll * kExpandWarps * gridDim.x || M.force_mode == EXPAND_QUAD + 1)
                leaf_body_quad<true, true>(M, ra, ia, n_leaves);    // four leaves per warp
            else leaf_body(M, ra, ia, n_leaves);            // route + ingest in one pass
            flush_block_stats();                            // (dominance hits of the leaf phase)
            grid.sync();
            stamp(2);
            if (threadIdx.x < 32) wave_counters(cnt, false);
            __syncthreads();
        } else {
            stamp(2);
        }
        if (s_status == SEARCH_RESOLVE) {
            // leaves with an unseen constraint-set transition are still pending: the host resolves and ingests them
            leave(SEARCH_RESOLVE, true);
            return;
        }
        stamp(3);
        stamp(4);
        if (threadIdx.x == 0) cs[S_WAVE]++;
    }
}

// Re-insert every state into a fresh table after growth (one thread per state).
__global__ void __launch_bounds__(256) rehash_kernel(const DevModel M, int32_t *table, long long mask,
                                                     const int32_t *state_key, long long n_states) {
    const int KW = M.key_words;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_states;
         s += (long long)gridDim.x * blockDim.x) {
        uint32_t h = 0;
        for (int j = 0; j < KW; j++) h ^= key_word_hash(state_key[s * KW + j], j);
        h = mix32(h);
        long long slot = (long long)h & mask;
        while (atomicCAS(&table[slot], -1, (int)s) != -1) slot = (slot + 1) & mask;
    }
}

// Fill relation tables: blockIdx.y = job (constraint), thread per entry (prefix tuple over the DECLARED domains),
// loop over the pivot's values.
__global__ void __launch_bounds__(128) build_table_kernel(const DevModel M, const int32_t *jobs, u64 *tables) {
    const DevCon con = M.cons[jobs[blockIdx.y]];
    const int n = con.n_scope, pv = con.pivot, entries = con.table_entries;
    int32_t cur[Limits::kMaxScope];
    int32_t stk[Limits::kMaxStack + 2];
    const Instr *code = M.code + con.code_off;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < entries; e += gridDim.x * blockDim.x) {
        int rest = e;
        for (int i = 0; i < n; i++) {
            if (i == pv) continue;
            const int v = M.scope[con.scope_off + i];
            const int w = M.width[v];
            cur[i] = M.lb[v] + rest % w;
            rest /= w;
        }
        const int vp = M.scope[con.scope_off + pv];
        u64 mask = 0ull;
        for (int b = 0; b < M.width[vp]; b++) {
            cur[pv] = M.lb[vp] + b;
            if (eval_tuple<1>(code, cur, stk, 0, M)) mask |= 1ull << b;
        }
        tables[con.table_off + e] = mask;
    }
}

__global__ void fill_kernel(int32_t *ptr, long long n, int32_t value) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ptr[i] = value;
}

}  // namespace

uint32_t capmap_hash(int cid, const int32_t *vals, int n) { return cap_hash(cid, vals, n); }

size_t expand_smem_bytes(const DevModel &m) {
    return node_bytes(m) * m.node_slots + scratch_bytes(m) * kExpandWarps + align8(sizeof(DevModel)) + align8((size_t)m.stage_bytes);
}

// Function attributes (dynamic shared memory above 48 KiB, carve-out) belong to the CONTEXT: one record per device.
constexpr int kMaxDevices = 64;
static int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d >= 0 && d < kMaxDevices ? d : 0;
}

static void configure_expand(size_t smem) {
    static size_t configured_dev[kMaxDevices] = {0};
    size_t &configured = configured_dev[current_device()];
    if (smem > configured) {
        if (smem > 48 * 1024) {
            cudaFuncSetAttribute(expand_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(expand_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        cudaFuncSetAttribute(expand_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(expand_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (smem > 48 * 1024) cudaFuncSetAttribute(expand_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(expand_quad_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = smem;
    }
}

int expand_max_grid(const DevModel &m, int sm_count) {
    const size_t smem = expand_smem_bytes(m);
    configure_expand(smem);
    static size_t cached_smem_dev[kMaxDevices];
    static int cached_per_sm_dev[kMaxDevices] = {0};
    static bool init_done = false;
    if (!init_done) {
        for (int i = 0; i < kMaxDevices; i++) cached_smem_dev[i] = ~(size_t)0;
        init_done = true;
    }
    size_t &cached_smem = cached_smem_dev[current_device()];
    int &cached_per_sm = cached_per_sm_dev[current_device()];
    if (smem == cached_smem) return cached_per_sm * sm_count;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, expand_kernel<false>, kExpandWarps * 32, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    cached_smem = smem;
    cached_per_sm = per_sm;
    return per_sm * sm_count;
}

void launch_expand(const DevModel &m, const ExpandArgs &a, int grid, int mode, cudaStream_t stream) {
    const size_t smem = expand_smem_bytes(m);
    configure_expand(smem);
    if (mode == EXPAND_CTA) expand_kernel<true><<<grid, kExpandWarps * 32, smem, stream>>>(m, a);
    else if (mode == EXPAND_QUAD) expand_quad_kernel<<<grid, kExpandWarps * 32, smem, stream>>>(m, a);
    else expand_kernel<false><<<grid, kExpandWarps * 32, smem, stream>>>(m, a);
}

int search_max_grid(const DevModel &m, int sm_count) {
    const size_t smem = expand_smem_bytes(m);
    static size_t configured_dev[kMaxDevices] = {0};
    size_t &configured = configured_dev[current_device()];
    if (smem > configured) {
        if (smem > 48 * 1024) {
            cudaFuncSetAttribute(search_kernel<kExpandCtasPerSm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(search_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        cudaFuncSetAttribute(search_kernel<kExpandCtasPerSm>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(search_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = smem;
    }
    static size_t cached_smem_dev[kMaxDevices];
    static int cached_per_sm_dev[kMaxDevices] = {0};
    static bool init_done = false;
    if (!init_done) {
        for (int i = 0; i < kMaxDevices; i++) cached_smem_dev[i] = ~(size_t)0;
        init_done = true;
    }
    size_t &cached_smem = cached_smem_dev[current_device()];
    int &cached_per_sm = cached_per_sm_dev[current_device()];
    if (smem == cached_smem) return cached_per_sm * sm_count;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel<kExpandCtasPerSm>, kExpandWarps * 32, smem) != cudaSuccess ||
        per_sm < 1)
        per_sm = 0;
    cached_smem = smem;
    cached_per_sm = per_sm;
    return per_sm * sm_count;
}

cudaError_t launch_search(const DevModel &m, const SearchArgs &a, int grid, int sm_count, cudaStream_t stream) {
    const size_t smem = expand_smem_bytes(m);
    DevModel mm = m;
    SearchArgs aa = a;
    void *params[] = {&mm, &aa};
    const void *fn = grid <= sm_count ? (const void *)search_kernel<1> : (const void *)search_kernel<kExpandCtasPerSm>;
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kExpandWarps * 32), params, smem, stream);
}

void launch_route(const DevModel &m, const RouteArgs &a, int grid, cudaStream_t stream) {
    route_kernel<<<grid, 256, 0, stream>>>(m, a);
}

void launch_ingest(const DevModel &m, const IngestArgs &a, int grid, cudaStream_t stream) {
    if (a.count <= 0) return;
    const size_t smem = a.n_segs > 0 ? (size_t)32 * m.rec_words * sizeof(int32_t) : 0;     // <= 48 KiB: checked by the caller
    ingest_kernel<<<grid, 256, smem, stream>>>(m, a);
}

void launch_scatter(const DevModel &m, const int32_t *leaves, long long n_leaves, const long long *dev_offsets,
                    unsigned long long *dev_fill, int32_t *outbox, int grid, cudaStream_t stream) {
    if (n_leaves <= 0) return;
    scatter_kernel<<<grid, 256, 0, stream>>>(m, leaves, n_leaves, dev_offsets, dev_fill, outbox);
}

void launch_gather(const int32_t *src, const int32_t *list, long long count, int rec_words, int32_t *dst, int grid,
                   cudaStream_t stream) {
    if (count <= 0) return;
    gather_kernel<<<grid, 256, 0, stream>>>(src, list, count, rec_words, dst);
}

void launch_rehash(const DevModel &m, int32_t *table, long long table_mask, const int32_t *state_key,
                   long long n_states, int grid, cudaStream_t stream) {
    if (n_states <= 0) return;
    rehash_kernel<<<grid, 256, 0, stream>>>(m, table, table_mask, state_key, n_states);
}

void launch_build_tables(const DevModel &m, const int32_t *dev_jobs, int n_jobs, int max_entries, unsigned long long *tables,
                         cudaStream_t stream) {
    if (n_jobs <= 0) return;
    dim3 grid((unsigned)std::min((max_entries + 127) / 128, 256), (unsigned)n_jobs);
    build_table_kernel<<<grid, 128, 0, stream>>>(m, dev_jobs, tables);
}

// CUDA loads kernels lazily, and loading one may have to wait for running kernels to finish.  A rank whose exchange
// kernel is waiting for a peer must never be the reason that peer cannot launch: a group loads every kernel up front.
void preload_search_kernels() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, expand_kernel<false>);
    cudaFuncGetAttributes(&fa, expand_kernel<true>);
    cudaFuncGetAttributes(&fa, expand_quad_kernel);
    cudaFuncGetAttributes(&fa, search_kernel<1>);
    cudaFuncGetAttributes(&fa, search_kernel<kExpandCtasPerSm>);
    cudaFuncGetAttributes(&fa, route_kernel);
    cudaFuncGetAttributes(&fa, ingest_kernel);
    cudaFuncGetAttributes(&fa, scatter_kernel);
    cudaFuncGetAttributes(&fa, gather_kernel);
    cudaFuncGetAttributes(&fa, rehash_kernel);
    cudaFuncGetAttributes(&fa, build_table_kernel);
    cudaFuncGetAttributes(&fa, fill_kernel);
    cudaGetLastError();
}

void launch_fill(int32_t *ptr, long long n, int32_t value, cudaStream_t stream) {
    if (n <= 0) return;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    fill_kernel<<<(int)blocks, 256, 0, stream>>>(ptr, n, value);
}

}  // namespace stcsp
