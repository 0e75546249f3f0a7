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
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <cstdio>

#include "error.h"
#include "sha256.h"
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

struct SolutionStore {
    std::vector<int32_t> cset, sig, src, dst, label;
    std::vector<uint8_t> fin;
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
    if (out->root_valid) {
        std::vector<int32_t> number(w.n, -1), order;
        std::deque<int32_t> queue;
        number[0] = 0;
        queue.push_back(0);
        std::vector<std::vector<int64_t>> sorted_edges;
        auto label_less = [&](int64_t x, int64_t y) {
            return std::lexicographical_compare(w.label(x), w.label(x) + w.nv, w.label(y), w.label(y) + w.nv);
        };
        while (!queue.empty()) {
            int32_t s = queue.front();
            queue.pop_front();
            order.push_back(s);
            std::vector<int64_t> es;
            for (int64_t e = w.first_edge[s]; e < w.first_edge[s + 1]; e++)
                if (w.alive[e]) es.push_back(e);
            std::sort(es.begin(), es.end(), label_less);
            for (int64_t e : es) {
                int32_t d = a->edge_dst[e];
                if (number[d] < 0) { number[d] = (int32_t)order.size() + (int32_t)queue.size(); queue.push_back(d); }
            }
            sorted_edges.push_back(std::move(es));
        }
        std::map<int32_t, int32_t> cmap;
        for (size_t i = 0; i < order.size(); i++) {
            int32_t s = order[i];
            int32_t c = (int32_t)cmap.emplace(a->state_cset[s], (int32_t)cmap.size()).first->second;
            st.cset.push_back(c);
            st.fin.push_back(w.fin[s]);
            for (int32_t k = 0; k < out->sig_len; k++) st.sig.push_back(s == 0 ? 0 : a->state_sig[(int64_t)s * a->sig_len + k]);
            for (int64_t e : sorted_edges[i]) {
                st.src.push_back((int32_t)i);
                st.dst.push_back(number[a->edge_dst[e]]);
                st.label.insert(st.label.end(), w.label(e), w.label(e) + w.nv);
            }
        }
    }
    out->n_states = (int64_t)st.cset.size();
    out->n_edges = (int64_t)st.src.size();
    out->state_cset = st.cset.data();
    out->state_final = st.fin.data();
    out->state_sig = st.sig.data();
    out->edge_src = st.src.data();
    out->edge_dst = st.dst.data();
    out->edge_label = st.label.data();
    return STCSP_OK;
}

void stcsp_solution_free(stcsp_solution_t *s) {
    if (!s) return;
    delete (Impl *)s->impl;
    memset(s, 0, sizeof *s);
}

}  // extern "C"

// Both texts are produced line by line through a sink, so they can be streamed to a file or into SHA-256 without
// ever holding the whole text (82 MB at partialorder_14, ~10 GB at partialorder_20).
namespace {

template <class Sink>
void emit_dot(const stcsp_problem_t *p, const stcsp_solution_t *s, Sink &&sink) {
    const Impl *impl = (const Impl *)s->impl;
    std::string line;
    line = "# Number of nodes = " + std::to_string(s->n_table_states) + "\n";
    sink(line);
    sink(header_line(p, s, false, nullptr, 0) + "\n");
    sink(header_line(p, s, true, impl->sig_vars.data(), (int32_t)impl->sig_vars.size()) + "\n");
    sink(std::string("digraph \"StCSP\" {\n"));
    int64_t e = 0;
    // the reference prints numSignVar + numUntil(distinct variables) values per vertex label
    for (int64_t v = 0; v < s->n_states; v++) {
        line = std::to_string(v) + " [shape=" + (s->state_final[v] ? "doublecircle" : "circle") + ", label=\"" +
               std::to_string(s->state_cset[v]) + ": ";
        if (v == 0) line += "S";
        else
            for (int32_t k = 0; k < s->sig_len; k++) {
                if (k) line += ", ";
                line += std::to_string(s->state_sig[v * s->sig_len + k]);
            }
        line += "\"];\n";
        sink(line);
        for (; e < s->n_edges && s->edge_src[e] == v; e++) {
            line = std::to_string(v) + " -> " + std::to_string(s->edge_dst[e]) + " [label=\"";
            for (int32_t k = 0; k < s->n_vars; k++) {
                if (k) line += ", ";
                line += std::to_string(s->edge_label[e * s->n_vars + k]);
            }
            line += "\"];\n";
            sink(line);
        }
    }
    sink(std::string("}\n"));
}

template <class Sink>
void emit_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s, Sink &&sink) {
    if (!s->root_valid) { sink(std::string("EMPTY")); return; }
    const Impl *impl = (const Impl *)s->impl;
    sink(header_line(p, s, false, nullptr, 0) + "\n");
    sink(header_line(p, s, true, impl->sig_vars.data(), (int32_t)impl->sig_vars.size()) + "\n");
    std::string line;
    for (int64_t v = 0; v < s->n_states; v++) {
        line = "V " + std::to_string(v) + " " + (s->state_final[v] ? "F" : "N") + " " + std::to_string(s->state_cset[v]) + " ";
        if (v == 0) line += "S";
        else
            for (int32_t k = 0; k < s->sig_len; k++) {
                if (k) line += " ";
                line += std::to_string(s->state_sig[v * s->sig_len + k]);
            }
        line += "\n";
        sink(line);
    }
    char buf[16];
    for (int64_t e = 0; e < s->n_edges; e++) {
        line = "E " + std::to_string(s->edge_src[e]) + " " + std::to_string(s->edge_dst[e]);
        const int32_t *lab = s->edge_label + e * s->n_vars;
        for (int32_t k = 0; k < s->n_vars; k++) {
            snprintf(buf, sizeof buf, " %d", lab[k]);
            line += buf;
        }
        line += "\n";
        sink(line);
    }
}

struct FileSink {
    FILE *f;
    bool ok = true;
    void operator()(const std::string &l) { ok = ok && fwrite(l.data(), 1, l.size(), f) == l.size(); }
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
    emit_dot(p, s, [&](const std::string &l) { out += l; });
    return dup_string(out);
}

char *stcsp_solution_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s) {
    std::string out;
    emit_canonical(p, s, [&](const std::string &l) { out += l; });
    return dup_string(out);
}

int stcsp_solution_write_dot(const stcsp_problem_t *p, const stcsp_solution_t *s, const char *path) {
    if (!p || !s || !path) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    return write_file(path, [&](FileSink &sink) { emit_dot(p, s, sink); });
}

int stcsp_solution_write_canonical(const stcsp_problem_t *p, const stcsp_solution_t *s, const char *path) {
    if (!p || !s || !path) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    return write_file(path, [&](FileSink &sink) { emit_canonical(p, s, sink); });
}

int stcsp_solution_canonical_sha256(const stcsp_problem_t *p, const stcsp_solution_t *s, char out_hex[65]) {
    if (!p || !s || !out_hex) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    stcsp::Sha256 sha;
    emit_canonical(p, s, [&](const std::string &l) { sha.update(l.data(), l.size()); });
    const std::string hex = sha.hex();
    memcpy(out_hex, hex.c_str(), 65);
    return STCSP_OK;
}

void stcsp_string_free(char *s) { free(s); }

}  // extern "C"
