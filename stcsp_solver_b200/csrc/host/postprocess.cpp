// Automaton post-processing and output: liveness w.r.t. `until`, the two adversarial fixpoints,
// canonical numbering, solutions.dot and canonical text.
//
// Restates, on flat arrays, what the reference does on its pointer graph after the search:
//   graphTraverse          src/graph.cpp:357-418   final / valid flags, drop edges into invalid states
//   adversarialTraverse    src/graph.cpp:304-355   flag -a: every value of variable #5 must have an edge
//   adversarialTraverse2   src/graph.cpp:247-302   flag -z: some value of #6 covers every value of #5
//   renumberVertex/vertexOut/edgeOut/solverOut     src/graph.cpp:41-101,420-442, src/solveralgorithm.cpp:709-730
// The reference's vertex numbering and line order come from hash_map iteration; here vertices are
// numbered breadth-first from the root with out-edges in label order, which
// is the canonical form the parity tests compare.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cstdio>

#include "error.h"
#include "sha256.h"
#include "trace_ranges.h"
#include "stcsp_host.h"

namespace {

struct Work {
    const stcsp_automaton_t *a;
    int64_t n, m;
    int32_t nv;
    std::vector<uint8_t> valid, fin, alive;
    std::vector<int64_t> first_edge;        // CSR over edges sorted by src
    std::vector<std::vector<int32_t>> parents;

    const int32_t *label(int64_t e) const { return a->edge_label + e * nv; }

    void build_parents(bool only_valid_sources) {
        parents.assign(n, {});
        for (int64_t s = 0; s < n; s++) {
            if (only_valid_sources && !valid[s]) continue;
            int32_t last = -1;
            for (int64_t e = first_edge[s]; e < first_edge[s + 1]; e++) {
                if (!alive[e]) continue;
                int32_t d = a->edge_dst[e];
                if (d == s || d == last) continue;
                // one entry per (parent, child) bucket is enough; duplicates are harmless
                parents[d].push_back((int32_t)s);
                last = d;
            }
        }
    }
    void drop_edges_into_invalid(bool include_root) {
        for (int64_t s = 0; s < n; s++) {
            if (!(valid[s] || (include_root && s == 0))) continue;
            for (int64_t e = first_edge[s]; e < first_edge[s + 1]; e++)
                if (alive[e] && !valid[a->edge_dst[e]]) alive[e] = 0;
        }
    }
};

bool check_all_values(const Work &w, int64_t s, int var, int32_t lb, int32_t ub) {
    for (int32_t c = lb; c <= ub; c++) {
        bool exist = false;
        for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1] && !exist; e++)
            exist = w.alive[e] && w.label(e)[var] == c && w.valid[w.a->edge_dst[e]];
        if (!exist) return false;
    }
    return true;
}

// reference checkVertexOutEdge2: keeps only edges whose avatar value answers every opponent value.
bool check_some_value_covers(Work &w, int64_t s, int op, int ava, int32_t ava_lb, int32_t ava_ub, int64_t op_count) {
    std::map<int32_t, std::set<int32_t>> seen;
    for (int32_t v = ava_lb; v <= ava_ub; v++) seen[v];
    bool node_valid = false;
    for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1]; e++) {
        if (!w.alive[e] || !w.valid[w.a->edge_dst[e]]) continue;
        auto &set = seen[w.label(e)[ava]];
        set.insert(w.label(e)[op]);
        if ((int64_t)set.size() == op_count) node_valid = true;
    }
    if (node_valid) {
        for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1]; e++) {
            if (!w.alive[e] || !w.valid[w.a->edge_dst[e]]) continue;
            int32_t av = w.label(e)[ava];
            if (av >= ava_lb && av <= ava_ub && (int64_t)seen[av].size() != op_count) w.alive[e] = 0;
        }
    }
    return node_valid;
}

template <class Check>
void greatest_fixpoint(Work &w, Check check) {
    std::deque<int32_t> todo;
    std::vector<uint8_t> queued(w.n, 1);
    for (int64_t s = 0; s < w.n; s++) todo.push_back((int32_t)s);
    while (!todo.empty()) {
        int32_t s = todo.front();
        todo.pop_front();
        queued[s] = 0;
        if (check(s)) continue;
        w.valid[s] = 0;
        for (int32_t p : w.parents[s])
            if (w.valid[p] && !queued[p]) { queued[p] = 1; todo.push_back(p); }
    }
}

// Host threads for the passes that are linear in the automaton (sorting out-edges, filling the solution arrays,
// formatting text): a partialorder_20-size automaton has 63 M edges and 10 GB of canonical text.
int host_threads() {
    static const int n = [] {
        int t = (int)std::thread::hardware_concurrency();
        if (const char *e = getenv("STCSP_HOST_THREADS")) t = atoi(e);
        return std::max(1, std::min(t, 32));
    }();
    return n;
}

// f(begin, end) over [0, n) in chunks of `grain`, handed out dynamically; inline when the range is small.
template <class F>
void parallel_chunks(int64_t n, int64_t grain, F f) {
    const int T = (int)std::min<int64_t>(host_threads(), (n + grain - 1) / grain);
    if (T <= 1) { if (n > 0) f((int64_t)0, n); return; }
    std::atomic<int64_t> next{0};
    auto body = [&] {
        for (;;) {
            const int64_t b = next.fetch_add(grain);
            if (b >= n) return;
            f(b, std::min(n, b + grain));
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(body);
    body();
    for (auto &x : th) x.join();
}

struct SolutionStore {
    // plain arrays, not vectors: nothing zero-fills 10 GB of labels before they are written
    std::unique_ptr<int32_t[]> cset, sig, src, dst, label;
    std::unique_ptr<uint8_t[]> fin;
    std::vector<int64_t> first;             // CSR over the canonical edges: out-edges of vertex v = [first[v], first[v+1])
};

char *dup_string(const std::string &s) {
    char *p = (char *)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

std::string header_line(const stcsp_problem_t *p, const stcsp_solution_t *s, bool sig_only, const int32_t *sig_vars,
                        int32_t n_sig_vars) {
    std::string out = "#";
    if (!sig_only) {
        for (int32_t v = 0; v < p->n_vars; v++) out += std::string(" ") + (p->var_names ? p->var_names[v] : "?");
    } else {
        for (int32_t i = 0; i < n_sig_vars; i++) out += std::string(" ") + (p->var_names ? p->var_names[sig_vars[i]] : "?");
    }
    (void)s;
    return out;
}

struct Impl {
    SolutionStore store;
    std::vector<int32_t> sig_vars;
};

}  // namespace

extern "C" {

int stcsp_postprocess(const stcsp_problem_t *problem, const stcsp_automaton_t *a, int32_t adversarial1,
                      int32_t adversarial2, stcsp_solution_t *out) {
    if (!problem || !a || !out) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    stcsp::TraceRange range("stcsp_postprocess (fixpoints, canonical numbering)");
    memset(out, 0, sizeof *out);
    Work w;
    w.a = a;
    w.n = a->n_states;
    w.m = a->n_edges;
    w.nv = a->n_vars;
    w.valid.assign(w.n, 0);
    w.fin.assign(w.n, 0);
    w.alive.assign(w.m, 1);
    w.first_edge.assign(w.n + 1, 0);
    for (int64_t e = 0; e < w.m; e++) {
        if (e && a->edge_src[e] < a->edge_src[e - 1]) { stcsp::set_error("edges are not sorted by source"); return STCSP_ERR_INVALID; }
        w.first_edge[a->edge_src[e] + 1]++;
    }
    for (int64_t s = 0; s < w.n; s++) w.first_edge[s + 1] += w.first_edge[s];

    const int32_t want = STCSP_POST_LIVENESS | (adversarial1 ? STCSP_POST_ADVERSARIAL1 : 0) | (adversarial2 ? STCSP_POST_ADVERSARIAL2 : 0);
    out->adver1 = out->adver2 = -1;
    if (adversarial1 || adversarial2) {
        if (problem->n_vars < (adversarial2 ? 7 : 6)) {
            stcsp::set_error("adversarial modes use variables #5 and #6 (reference src/graph.cpp:275,329); the model has too few");
            return STCSP_ERR_INVALID;
        }
    }
    const bool have_flags = a->post_applied && a->state_valid && a->state_final && (a->edge_alive || w.m == 0);
    const int32_t applied = have_flags ? a->post_applied : 0;
    if (applied & ~want) {
        stcsp::set_error("the automaton was post-processed on the device with -a / -z flags this call does not ask for");
        return STCSP_ERR_INVALID;
    }
    if ((applied & STCSP_POST_ADVERSARIAL2) && adversarial1 && !(applied & STCSP_POST_ADVERSARIAL1)) {
        stcsp::set_error("the device ran -z without -a; -a cannot follow it (the reference runs -a first, src/solver.cpp:300-321)");
        return STCSP_ERR_INVALID;
    }
    if (applied) {
        // ---- the device ran these fixpoints (automaton.cu: liveness / -a / -z sweeps): take its flags; what it did not run
        // (the caller asked stcsp_gpu_solve for less than it asks for here) follows on the host below
        for (int64_t s = 0; s < w.n; s++) { w.valid[s] = a->state_valid[s]; w.fin[s] = a->state_final[s]; }
        for (int64_t e = 0; e < w.m; e++) w.alive[e] = a->edge_alive[e];
        if (applied & STCSP_POST_ADVERSARIAL1) out->adver1 = a->adver1;
        if (applied & STCSP_POST_ADVERSARIAL2) out->adver2 = a->adver2;
    } else {
        // ---- graphTraverse: final = all until flags set (the reference looks at numUntil flags, which
        // counts distinct right-hand variables), valid = can reach a final state.
        for (int64_t s = 1; s < w.n; s++) {
            bool f = true;
            for (int32_t c = a->n_sig_vars; f && c < a->n_sig_vars + a->n_until_vars; c++)
                f = a->state_sig[s * a->sig_len + c] == 1;
            w.fin[s] = w.valid[s] = f;
        }
        if (w.n > 0) w.fin[0] = w.valid[0] = a->root_final != 0;
        bool all_valid = true;
        for (int64_t s = 0; s < w.n && all_valid; s++) all_valid = w.valid[s] != 0;
        // (a model without `until` -- every shipped and generated benchmark instance -- makes every state final: nothing to
        //  propagate, and at partialorder_20 size the parent lists alone would be 63 M entries)
        if (!all_valid) {
            w.build_parents(false);
            std::vector<int32_t> stack;
            for (int64_t s = 1; s < w.n; s++)
                if (w.valid[s]) stack.push_back((int32_t)s);
            while (!stack.empty()) {
                int32_t s = stack.back();
                stack.pop_back();
                for (int32_t p : w.parents[s])
                    if (!w.valid[p]) { w.valid[p] = 1; stack.push_back(p); }
            }
            w.drop_edges_into_invalid(true);
        }
    }
    if (adversarial1 && !(applied & STCSP_POST_ADVERSARIAL1)) {
        w.build_parents(false);
        int32_t lb = problem->var_lb[5], ub = problem->var_ub[5];
        greatest_fixpoint(w, [&](int32_t s) { return check_all_values(w, s, 5, lb, ub); });
        w.drop_edges_into_invalid(false);
        out->adver1 = w.valid[0];
    }
    if (adversarial2 && !(applied & STCSP_POST_ADVERSARIAL2)) {
        w.build_parents(true);
        int32_t alb = problem->var_lb[6], aub = problem->var_ub[6];
        int64_t opn = (int64_t)problem->var_ub[5] - problem->var_lb[5] + 1;
        greatest_fixpoint(w, [&](int32_t s) { return check_some_value_covers(w, s, 5, 6, alb, aub, opn); });
        if (w.valid[0]) w.drop_edges_into_invalid(false);
        out->adver2 = w.valid[0];
    }

    // ---- canonical numbering of what graphOut would print
    auto *impl = new Impl();
    out->impl = impl;
    out->root_valid = w.n > 0 && w.valid[0];
    out->n_table_states = w.n;
    out->n_vars = a->n_vars;
    // vertex labels show numSignVar + numUntil values (src/graph.cpp:55), numUntil = distinct right-hand variables
    out->sig_len = a->n_sig_vars + a->n_until_vars;
    impl->sig_vars.assign(a->sig_vars, a->sig_vars + a->n_sig_vars);
    SolutionStore &st = impl->store;
    int64_t n_out = 0, m_out = 0;
    if (out->root_valid) {
        // out-edges of every state in label order: one index array over the surviving edges, ranges sorted in parallel
        std::vector<int64_t> pfirst(w.n + 1, 0);
        for (int64_t s = 0; s < w.n; s++) {
            int64_t c = 0;
            for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1]; e++) c += w.alive[e] != 0;
            pfirst[s + 1] = pfirst[s] + c;
        }
        std::unique_ptr<int64_t[]> perm(new int64_t[pfirst[w.n] + 1]);
        const int32_t nv = w.nv;
        const int32_t *labels = a->edge_label;
        parallel_chunks(w.n, 4096, [&](int64_t b, int64_t e_) {
            for (int64_t s = b; s < e_; s++) {
                int64_t *o = perm.get() + pfirst[s], *o0 = o;
                bool sorted = true;
                for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1]; e++) {
                    if (!w.alive[e]) continue;
                    if (o != o0 && sorted)
                        sorted = !std::lexicographical_compare(labels + e * nv, labels + e * nv + nv, labels + o[-1] * nv, labels + o[-1] * nv + nv);
                    *o++ = e;
                }
                if (!sorted)
                    std::sort(o0, o, [&](int64_t x, int64_t y) {
                        return std::lexicographical_compare(labels + x * nv, labels + x * nv + nv, labels + y * nv, labels + y * nv + nv);
                    });
            }
        });
        // breadth-first numbering from the root: the visiting order IS the queue
        std::vector<int32_t> number(w.n, -1), order;
        order.reserve(w.n);
        number[0] = 0;
        order.push_back(0);
        for (size_t head = 0; head < order.size(); head++) {
            const int32_t s = order[head];
            for (int64_t i = pfirst[s]; i < pfirst[s + 1]; i++) {
                const int32_t d = a->edge_dst[perm[i]];
                if (number[d] < 0) { number[d] = (int32_t)order.size(); order.push_back(d); }
            }
        }
        n_out = (int64_t)order.size();
        st.first.assign(n_out + 1, 0);
        for (int64_t i = 0; i < n_out; i++) st.first[i + 1] = st.first[i] + (pfirst[order[i] + 1] - pfirst[order[i]]);
        m_out = st.first[n_out];
        const int32_t SL = out->sig_len;
        st.cset.reset(new int32_t[n_out + 1]);
        st.fin.reset(new uint8_t[n_out + 1]);
        st.sig.reset(new int32_t[n_out * (int64_t)SL + 1]);
        st.src.reset(new int32_t[m_out + 1]);
        st.dst.reset(new int32_t[m_out + 1]);
        st.label.reset(new int32_t[m_out * (int64_t)nv + 1]);
        std::map<int32_t, int32_t> cmap;        // constraint sets in order of first appearance (a handful)
        for (int64_t i = 0; i < n_out; i++)
            st.cset[i] = cmap.emplace(a->state_cset[order[i]], (int32_t)cmap.size()).first->second;
        parallel_chunks(n_out, 2048, [&](int64_t b, int64_t e_) {
            for (int64_t i = b; i < e_; i++) {
                const int32_t s = order[i];
                st.fin[i] = w.fin[s];
                for (int32_t k = 0; k < SL; k++) st.sig[i * SL + k] = s == 0 ? 0 : a->state_sig[(int64_t)s * a->sig_len + k];
                int64_t o = st.first[i];
                for (int64_t j = pfirst[s]; j < pfirst[s + 1]; j++, o++) {
                    const int64_t e = perm[j];
                    st.src[o] = (int32_t)i;
                    st.dst[o] = number[a->edge_dst[e]];
                    memcpy(st.label.get() + o * nv, labels + e * nv, (size_t)nv * 4);
                }
            }
        });
    } else {
        st.first.assign(1, 0);
    }
    out->n_states = n_out;
    out->n_edges = m_out;
    out->state_cset = st.cset.get();
    out->state_final = st.fin.get();
    out->state_sig = st.sig.get();
    out->edge_src = st.src.get();
    out->edge_dst = st.dst.get();
    out->edge_label = st.label.get();
    return STCSP_OK;
}

void stcsp_solution_free(stcsp_solution_t *s) {
    if (!s) return;
    delete (Impl *)s->impl;
    memset(s, 0, sizeof *s);
}

}  // extern "C"

// Both texts are produced through a sink in chunks of a megabyte or so, so they can be streamed to a file or into
// SHA-256 without ever holding the whole text (82 MB at partialorder_14, ~10 GB at partialorder_20).  Chunks are
// formatted by host threads side by side and handed to the sink in order while the next ones are being formatted.
namespace {

struct TextBuf {
    std::vector<char> mem;
    size_t len = 0;
    char *room(size_t n) {                 // at least n more bytes
        if (len + n > mem.size()) mem.resize(std::max(mem.size() * 2, len + n + (1 << 16)));
        return mem.data() + len;
    }
    void put(const char *p, size_t n) { memcpy(room(n), p, n); len += n; }
    void put(const std::string &t) { put(t.data(), t.size()); }
};

inline char *put_int(char *p, int64_t v) {
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    if (v < 0) *p++ = '-';
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char *put_lit(char *p, const char *lit, size_t n) { memcpy(p, lit, n); return p + n; }
#define LIT(p, s) put_lit(p, s, sizeof(s) - 1)

// fmt(chunk index, buffer) for chunks 0 .. n_chunks-1; the sink sees the buffers in chunk order
template <class Sink, class Fmt>
void emit_chunks(int64_t n_chunks, Fmt fmt, Sink &sink) {
    const int T = (int)std::min<int64_t>(host_threads(), n_chunks);
    if (T <= 1) {
        TextBuf b;
        for (int64_t c = 0; c < n_chunks; c++) { b.len = 0; fmt(c, b); sink(b.mem.data(), b.len); }
        return;
    }
    std::vector<TextBuf> bufs[2] = {std::vector<TextBuf>(T), std::vector<TextBuf>(T)};
    int64_t filled[2] = {0, 0};
    for (int64_t base = 0, r = 0; base < n_chunks || filled[(r + 1) & 1]; base += T, r++) {
        const int cur = (int)(r & 1), prev = cur ^ 1;
        const int64_t cnt = std::max<int64_t>(0, std::min<int64_t>(T, n_chunks - base));
        std::vector<std::thread> th;
        for (int64_t t = 0; t < cnt; t++)
            th.emplace_back([&, t] { bufs[cur][t].len = 0; fmt(base + t, bufs[cur][t]); });
        for (int64_t t = 0; t < filled[prev]; t++) sink(bufs[prev][t].mem.data(), bufs[prev][t].len);   // the round before, in order
        filled[prev] = 0;
        for (auto &x : th) x.join();
        filled[cur] = cnt;
    }
}

const std::vector<int64_t> &edge_first(const stcsp_solution_t *s) { return ((const Impl *)s->impl)->store.first; }

// vertices [v0, v1) cut into chunks of about `weight` lines (a vertex with its out-edges stays in one chunk)
std::vector<int64_t> vertex_chunks(const stcsp_solution_t *s, int64_t weight) {
    const std::vector<int64_t> &first = edge_first(s);
    std::vector<int64_t> cut{0};
    int64_t acc = 0;
    for (int64_t v = 0; v < s->n_states; v++) {
        acc += 1 + first[v + 1] - first[v];
        if (acc >= weight) { cut.push_back(v + 1); acc = 0; }
    }
    if (cut.back() != s->n_states) cut.push_back(s->n_states);
    return cut;
}

template <class Sink>
void emit_dot(const stcsp_problem_t *p, const stcsp_solution_t *s, Sink &&sink) {
    const Impl *impl = (const Impl *)s->impl;
    std::string head = "# Number of nodes = " + std::to_string(s->n_table_states) + "\n";
    head += header_line(p, s, false, nullptr, 0) + "\n";
    head += header_line(p, s, true, impl->sig_vars.data(), (int32_t)impl->sig_vars.size()) + "\n";
    head += "digraph \"StCSP\" {\n";
    sink(head.data(), head.size());
    const std::vector<int64_t> &first = edge_first(s);
    const std::vector<int64_t> cut = vertex_chunks(s, 1 << 14);
    const size_t vmax = 64 + 14 * (size_t)std::max(1, s->sig_len), emax = 48 + 14 * (size_t)std::max(1, s->n_vars);
    // the reference prints numSignVar + numUntil(distinct variables) values per vertex label
    emit_chunks((int64_t)cut.size() - 1, [&](int64_t c, TextBuf &b) {
        for (int64_t v = cut[c]; v < cut[c + 1]; v++) {
            char *q = b.room(vmax), *q0 = q;
            q = put_int(q, v);
            q = s->state_final[v] ? LIT(q, " [shape=doublecircle, label=\"") : LIT(q, " [shape=circle, label=\"");
            q = put_int(q, s->state_cset[v]);
            q = LIT(q, ": ");
            if (v == 0) *q++ = 'S';
            else
                for (int32_t k = 0; k < s->sig_len; k++) {
                    if (k) q = LIT(q, ", ");
                    q = put_int(q, s->state_sig[v * s->sig_len + k]);
                }
            q = LIT(q, "\"];\n");
            b.len += q - q0;
            for (int64_t e = first[v]; e < first[v + 1]; e++) {
                q = q0 = b.room(emax);
                q = put_int(q, v);
                q = LIT(q, " -> ");
                q = put_int(q, s->edge_dst[e]);
                q = LIT(q, " [label=\"");
                const int32_t *lab = s->edge_label + e * s->n_vars;
                for (int32_t k = 0; k < s->n_vars; k++) {
                    if (k) q = LIT(q, ", ");
                    q = put_int(q, lab[k]);
                }
                q = LIT(q, "\"];\n");
                b.len += q - q0;
            }
        }
    }, sink);
    sink("}\n", 2);
}

template <class Sink>
void emit_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s, Sink &&sink) {
    if (!s->root_valid) { sink("EMPTY", 5); return; }
    const Impl *impl = (const Impl *)s->impl;
    std::string head = header_line(p, s, false, nullptr, 0) + "\n";
    head += header_line(p, s, true, impl->sig_vars.data(), (int32_t)impl->sig_vars.size()) + "\n";
    sink(head.data(), head.size());
    const size_t vmax = 64 + 12 * (size_t)std::max(1, s->sig_len), emax = 48 + 12 * (size_t)std::max(1, s->n_vars);
    const int64_t VC = 1 << 14, EC = 1 << 14;
    emit_chunks((s->n_states + VC - 1) / VC, [&](int64_t c, TextBuf &b) {
        for (int64_t v = c * VC; v < std::min(s->n_states, (c + 1) * VC); v++) {
            char *q = b.room(vmax), *q0 = q;
            q = LIT(q, "V ");
            q = put_int(q, v);
            q = s->state_final[v] ? LIT(q, " F ") : LIT(q, " N ");
            q = put_int(q, s->state_cset[v]);
            *q++ = ' ';
            if (v == 0) *q++ = 'S';
            else
                for (int32_t k = 0; k < s->sig_len; k++) {
                    if (k) *q++ = ' ';
                    q = put_int(q, s->state_sig[v * s->sig_len + k]);
                }
            *q++ = '\n';
            b.len += q - q0;
        }
    }, sink);
    emit_chunks((s->n_edges + EC - 1) / EC, [&](int64_t c, TextBuf &b) {
        for (int64_t e = c * EC; e < std::min(s->n_edges, (c + 1) * EC); e++) {
            char *q = b.room(emax), *q0 = q;
            q = LIT(q, "E ");
            q = put_int(q, s->edge_src[e]);
            *q++ = ' ';
            q = put_int(q, s->edge_dst[e]);
            const int32_t *lab = s->edge_label + e * s->n_vars;
            for (int32_t k = 0; k < s->n_vars; k++) {
                *q++ = ' ';
                q = put_int(q, lab[k]);
            }
            *q++ = '\n';
            b.len += q - q0;
        }
    }, sink);
}

struct FileSink {
    FILE *f;
    bool ok = true;
    void operator()(const char *p, size_t n) { ok = ok && fwrite(p, 1, n, f) == n; }
};

template <class Emit>
int write_file(const char *path, Emit &&emit) {
    FILE *f = fopen(path, "wb");
    if (!f) { stcsp::set_error(std::string("cannot write ") + path); return STCSP_ERR_INVALID; }
    std::vector<char> big(1 << 20);
    setvbuf(f, big.data(), _IOFBF, big.size());
    FileSink sink{f};
    emit(sink);
    const bool closed = fclose(f) == 0;
    if (!sink.ok || !closed) { stcsp::set_error(std::string("short write to ") + path); return STCSP_ERR_INVALID; }
    return STCSP_OK;
}

}  // namespace

extern "C" {

char *stcsp_solution_dot(const stcsp_problem_t *p, const stcsp_solution_t *s) {
    std::string out;
    emit_dot(p, s, [&](const char *t, size_t n) { out.append(t, n); });
    return dup_string(out);
}

char *stcsp_solution_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s) {
    std::string out;
    emit_canonical(p, s, [&](const char *t, size_t n) { out.append(t, n); });
    return dup_string(out);
}

int stcsp_solution_write_dot(const stcsp_problem_t *p, const stcsp_solution_t *s, const char *path) {
    if (!p || !s || !path) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    stcsp::TraceRange range("stcsp_solution_write_dot");
    return write_file(path, [&](FileSink &sink) { emit_dot(p, s, sink); });
}

int stcsp_solution_write_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s, const char *path) {
    if (!p || !s || !path) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    return write_file(path, [&](FileSink &sink) { emit_canonical(p, s, sink); });
}

int stcsp_solution_canonical_sha256(const stcsp_problem_t *p, const stcsp_solution_t *s, char out_hex[65]) {
    if (!p || !s || !out_hex) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    stcsp::TraceRange range("stcsp_solution_canonical_sha256");
    stcsp::Sha256 sha;
    emit_canonical(p, s, [&](const char *t, size_t n) { sha.update(t, n); });
    const std::string hex = sha.hex();
    memcpy(out_hex, hex.c_str(), 65);
    return STCSP_OK;
}

void stcsp_string_free(char *s) { free(s); }

}  // extern "C"
