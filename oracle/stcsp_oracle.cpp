/* TEST INFRASTRUCTURE -- CPU restatement of the reference's search-and-propagate path.
 *
 * This file is the ORACLE for the CUDA path in stcsp_solver_b200/csrc: a plain, single-threaded
 * restatement of the reference's own algorithm (depth-first search over time points, binary
 * domain splitting, bounds GAC with nested-loop support search, tree-walking evaluator,
 * hash-free state lookup).  It is only ever used by tests/, by __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference leg; the product never links or calls it.
 *
 * PARITY PINNING: the reference ships no golden vectors or tests.  This
 * restatement is pinned against the reference ITSELF: oracle/_ref/stcsp_ref (the unmodified
 * reference sources + a stand-in parser, oracle/build_ref.sh) was run on all 26 shipped models
 * and on the feature probes of oracle/make_goldens.py; tests/test_oracle.py checks that this file
 * reproduces the same canonical automaton AND the same search statistics (numNodes, numFails,
 * numDominance of the reference's stat line, src/solveralgorithm.cpp:1001), which it can only do
 * by exploring the same search tree with the same propagation strength.
 *
 * Each function cites the reference lines it follows.  Input and output are the flat structs of
 * include/stcsp_b200.h, so the oracle and the CUDA library are called with the same problem.
 *
 * Defined behaviour where the reference traps: `/` or `%` by zero (and INT_MIN / -1) poisons the
 * evaluation exactly like an out-of-range array index (the tuple does not satisfy the constraint).
 */
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <pthread.h>

#include "stcsp_b200.h"

namespace {

/* ------------------------------------------------------------------------------ expression trees */
struct Node;                                   /* reference ConstraintNode, src/constraint.h:24-31 */
typedef std::shared_ptr<Node> NodeP;
struct Node {
    int op, arg;
    NodeP a, b, c;                             /* operands in evaluation order */
    Node(int o, int g) : op(o), arg(g) {}
};

int arity(int op) {
    if (op == STCSP_OP_CONST || op == STCSP_OP_VAR) return 0;
    if (op >= STCSP_OP_ARR && op <= STCSP_OP_AT) return 1;
    if (op == STCSP_OP_IF) return 3;
    if ((op >= STCSP_OP_LT && op <= STCSP_OP_MOD) || (op >= STCSP_CON_LT && op <= STCSP_CON_UNTIL)) return 2;
    return -1;
}

NodeP build(const stcsp_tok_t *t, int n) {
    std::vector<NodeP> st;
    for (int i = 0; i < n; i++) {
        int ar = arity(t[i].op);
        if (ar < 0 || (int)st.size() < ar) throw std::runtime_error("oracle: malformed token list");
        NodeP e = std::make_shared<Node>(t[i].op, t[i].arg);
        if (ar == 3) { e->c = st.back(); st.pop_back(); }
        if (ar >= 2) { e->b = st.back(); st.pop_back(); }
        if (ar >= 1) { e->a = st.back(); st.pop_back(); }
        st.push_back(e);
    }
    if (st.size() != 1) throw std::runtime_error("oracle: token list is not one tree");
    return st[0];
}

bool node_eq(const NodeP &x, const NodeP &y) {            /* constraintNodeEq, src/constraint.cpp:551-561 */
    if (!x && !y) return true;
    if (!x || !y) return false;
    return x->op == y->op && x->arg == y->arg && node_eq(x->a, y->a) && node_eq(x->b, y->b) && node_eq(x->c, y->c);
}

bool has_first(const NodeP &e) {                          /* constraintNodeHasFirst, src/constraint.cpp:240-250 */
    if (!e) return false;
    if (e->op == STCSP_OP_FIRST || e->op == STCSP_OP_AT) return true;
    return has_first(e->a) || has_first(e->b) || has_first(e->c);
}

/* -------------------------------------------------------------------------------- solver state */
enum { K_NEXT = 0, K_POINT = 1, K_UNTIL = 2, K_AT = 3 };  /* src/constraint.h:33-36 */

struct Con {                                              /* reference Constraint, src/constraint.h:38-49 */
    NodeP root;
    int kind;
    bool hasFirst;
    std::vector<int> arcs;                                /* variables in discovery order (one arc each) */
    std::vector<int> vars;                                /* the same list reversed (src/constraint.cpp:307-314) */
};

struct CSet {                                             /* one entry of solver->seenConstraints */
    std::vector<Con> cons;
    std::vector<std::vector<int> > varCons;               /* var -> constraints it occurs in, queue order */
};

struct Dom {                                              /* currLB/currUB windows, src/variable.h:19-20 */
    std::vector<int> lb, ub;                              /* [var * k + offset] */
};

struct Vertex {
    int cset;
    std::vector<int> sig;
    bool fail;
};

struct Stats {
    int64_t gac_calls, validates, node_visits, revisions, leaves, splits, max_depth;
};

struct Oracle {
    const stcsp_problem_t *p;
    int V, k;
    std::vector<std::vector<int> > arrays;
    std::vector<CSet> seen;                               /* solver->seenConstraints */
    bool solverHasFirst;
    std::vector<int> isSignature;                         /* per variable */
    int numSignVar, numUntilVars;
    std::vector<int> untilSeen;
    /* automaton */
    std::vector<Vertex> vertices;
    std::map<std::pair<int, std::vector<int> >, int> table;   /* VertexTable, src/graph.h:64 */
    std::vector<int> esrc, edst;
    std::vector<int> elabel;
    /* statistics of the reference's stat line */
    int64_t numFails, numDominance;
    Stats st;
    std::vector<int> propagateValue;                      /* Variable::propagateValue */
    double deadline;                                      /* seconds since epoch, 0 = none */
    bool timed_out;

    /* ---- constraint classification: solverConstraintQueuePush, src/constraint.cpp:254-318 */
    void link_vars(const NodeP &e, std::vector<int> &out) {
        if (!e) return;
        if (e->op == STCSP_OP_VAR) {
            if (std::find(out.begin(), out.end(), e->arg) == out.end()) out.push_back(e->arg);
            return;
        }
        link_vars(e->a, out); link_vars(e->b, out); link_vars(e->c, out);
    }
    void push(CSet &q, const NodeP &root) {
        Con c;
        c.root = root;
        if (root->op == STCSP_CON_UNTIL) {
            c.kind = K_UNTIL;
            int y = root->b->arg;
            if (!untilSeen[y]) { untilSeen[y] = 1; numUntilVars++; }
        } else if (root->b && root->b->op == STCSP_OP_NEXT) {
            c.kind = K_NEXT;
            int x = root->a->arg;
            if (!isSignature[x]) { isSignature[x] = 1; numSignVar++; }
        } else if (root->b && root->b->op == STCSP_OP_AT) {
            c.kind = K_AT;
        } else {
            c.kind = K_POINT;
        }
        c.hasFirst = has_first(root);
        if (c.hasFirst) solverHasFirst = true;
        link_vars(root, c.arcs);
        c.vars.assign(c.arcs.rbegin(), c.arcs.rend());
        int id = (int)q.cons.size();
        for (size_t i = 0; i < c.arcs.size(); i++) q.varCons[c.arcs[i]].push_back(id);
        q.cons.push_back(c);
    }

    /* ---- lifted constant folding: constraintNodeValue, src/constraint.cpp:335-439 (quirks kept) */
    bool fold(const NodeP &e, int &out) {                 /* false = unknown */
        int l, r;
        switch (e->op) {
            case STCSP_OP_VAR: case STCSP_OP_NEXT: case STCSP_OP_AT: return false;
            case STCSP_OP_CONST: out = e->arg; return true;
            case STCSP_OP_FIRST: return fold(e->a, out);
            case STCSP_OP_ARR:
                if (!fold(e->a, l)) return false;
                if (l < 0 || l >= (int)arrays[e->arg].size()) return false;
                out = arrays[e->arg][l]; return true;
            case STCSP_OP_ABS: if (!fold(e->a, l)) return false; out = l < 0 ? (int)(0u - (unsigned)l) : l; return true;
            case STCSP_OP_IF: if (!fold(e->a, l)) return false; return fold(l ? e->b : e->c, out);
            case STCSP_OP_NOT: if (!fold(e->a, l)) return false; out = (l == 0); return true;
            case STCSP_OP_AND: if (!fold(e->a, l)) return false; if (l == 0) { out = 0; return true; } return fold(e->b, out);
            case STCSP_OP_OR: if (!fold(e->a, l)) return false; if (l != 1) { out = 1; return true; } return fold(e->b, out);
            default: break;
        }
        if (!e->a || !e->b) return false;
        bool lk = fold(e->a, l), rk = fold(e->b, r);
        if (!lk || !rk) return false;
        switch (e->op) {
            case STCSP_OP_LT: out = l < r; return true;
            case STCSP_OP_GT: out = l < r; return true;            /* src/constraint.cpp:425 */
            case STCSP_OP_LE: out = l <= r; return true;
            case STCSP_OP_GE: out = l >= r; return true;
            case STCSP_OP_EQ: out = l == r; return true;
            case STCSP_OP_NE: out = l != r; return true;
            case STCSP_OP_ADD: out = (int)((unsigned)l + (unsigned)r); return true;
            case STCSP_OP_SUB: out = (int)((unsigned)l - (unsigned)r); return true;
            case STCSP_OP_MUL: out = (int)((unsigned)l * (unsigned)r); return true;
            case STCSP_OP_DIV: if (r == 0 || (l == INT_MIN && r == -1)) return false; out = l / r; return true;
            case STCSP_OP_MOD: if (r == 0 || (l == INT_MIN && r == -1)) return false; out = l % r; return true;
            default: return false;
        }
    }
    bool tautology(const NodeP &root) {                   /* src/constraint.cpp:443-462 */
        int l, r;
        if (!root->a || !root->b) return false;
        bool lk = fold(root->a, l), rk = fold(root->b, r);
        if (!lk || !rk) return false;
        switch (root->op) {
            case STCSP_CON_LT: return l < r;
            case STCSP_CON_GT: return l > r;
            case STCSP_CON_LE: return l <= r;
            case STCSP_CON_GE: return l >= r;
            case STCSP_CON_EQ: return l == r;
            case STCSP_CON_NE: return l != r;
            case STCSP_CON_IMPLY: return l <= r;
            case STCSP_CON_UNTIL: return r == 1;
            default: return false;
        }
    }

    /* ---- leaf rewrite: constraintNodeTranslate{,First,AT}, src/constraint.cpp:466-548 */
    NodeP subst(const NodeP &e, const Dom &d) {           /* variables -> the values just taken */
        if (!e) return e;
        if (e->op == STCSP_OP_VAR) return std::make_shared<Node>(STCSP_OP_CONST, d.lb[e->arg * k]);
        NodeP n = std::make_shared<Node>(e->op, e->arg);
        n->a = subst(e->a, d); n->b = subst(e->b, d); n->c = subst(e->c, d);
        return n;
    }
    NodeP translate(const NodeP &e, const Dom &d) {
        if (!e) return e;
        if (e->op == STCSP_OP_FIRST) {
            NodeP s = subst(e->a, d);
            int v;
            if (fold(s, v)) return std::make_shared<Node>(STCSP_OP_CONST, v);
            return s;                                     /* reference logs an error and keeps the subtree */
        }
        if (e->op == STCSP_CON_EQ && e->b && e->b->op == STCSP_OP_AT) {
            NodeP n = std::make_shared<Node>(STCSP_CON_EQ, 0);
            n->a = std::make_shared<Node>(STCSP_OP_VAR, e->a->arg);
            NodeP y = std::make_shared<Node>(STCSP_OP_VAR, e->b->a->arg);
            if (e->b->arg == 1) {
                n->b = std::make_shared<Node>(STCSP_OP_FIRST, 0);
                n->b->a = y;
            } else {
                n->b = std::make_shared<Node>(STCSP_OP_AT, e->b->arg - 1);
                n->b->a = y;
            }
            return n;
        }
        NodeP n = std::make_shared<Node>(e->op, e->arg);
        n->a = translate(e->a, d); n->b = translate(e->b, d); n->c = translate(e->c, d);
        return n;
    }

    /* ---- evaluator: solverValidateRe, src/solveralgorithm.cpp:336-424 */
    int eval(const Node *e, bool &valid) {
        st.node_visits++;
        int l, r, t;
        switch (e->op) {
            case STCSP_OP_VAR: return propagateValue[e->arg];
            case STCSP_OP_CONST: return e->arg;
            case STCSP_OP_ARR:
                t = eval(e->a.get(), valid);
                if (t < 0 || t >= (int)arrays[e->arg].size()) { valid = false; return 0; }
                return arrays[e->arg][t];
            case STCSP_OP_ABS: t = eval(e->a.get(), valid); return t < 0 ? (int)(0u - (unsigned)t) : t;
            case STCSP_OP_IF: t = eval(e->a.get(), valid); return t ? eval(e->b.get(), valid) : eval(e->c.get(), valid);
            case STCSP_OP_FIRST: case STCSP_OP_AT: return eval(e->a.get(), valid);
            case STCSP_OP_NOT: t = eval(e->a.get(), valid); return t == 0 ? 1 : 0;
            case STCSP_OP_AND: t = eval(e->a.get(), valid); return t ? eval(e->b.get(), valid) : 0;
            case STCSP_OP_OR: t = eval(e->a.get(), valid); return t ? 1 : eval(e->b.get(), valid);
            case STCSP_CON_IMPLY:
                l = eval(e->a.get(), valid);
                if (l == 0) return 1;
                r = eval(e->b.get(), valid);
                return l <= r;
            default: break;
        }
        if (!e->a || !e->b) return 0;                     /* e.g. a stray NEXT: evaluates to 0 */
        l = eval(e->a.get(), valid);
        r = eval(e->b.get(), valid);
        if (!valid) return 0;
        switch (e->op) {
            case STCSP_CON_LT: case STCSP_OP_LT: return l < r;
            case STCSP_CON_GT: case STCSP_OP_GT: return l > r;
            case STCSP_CON_LE: case STCSP_OP_LE: return l <= r;
            case STCSP_CON_GE: case STCSP_OP_GE: return l >= r;
            case STCSP_CON_EQ: case STCSP_OP_EQ: return l == r;
            case STCSP_CON_NE: case STCSP_OP_NE: return l != r;
            case STCSP_OP_ADD: return (int)((unsigned)l + (unsigned)r);
            case STCSP_OP_SUB: return (int)((unsigned)l - (unsigned)r);
            case STCSP_OP_MUL: return (int)((unsigned)l * (unsigned)r);
            case STCSP_OP_DIV: if (r == 0 || (l == INT_MIN && r == -1)) { valid = false; return 0; } return l / r;
            case STCSP_OP_MOD: if (r == 0 || (l == INT_MIN && r == -1)) { valid = false; return 0; } return l % r;
            default: return 0;
        }
    }
    struct TimeUp {};
    bool validate(const Con &c) {                         /* src/solveralgorithm.cpp:428-431 */
        st.validates++;
        if (deadline > 0 && (st.validates & 0xFFFFF) == 0) {      /* a single propagation may run for hours */
            double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            if (now > deadline) { timed_out = true; throw TimeUp(); }
        }
        bool valid = true;
        return eval(c.root.get(), valid) != 0;
    }

    /* ---- support search: findSupportRe / findSupport, src/solveralgorithm.cpp:435-470 */
    bool findSupportRe(const Con &c, int var, int point, int index, int numVar, const Dom &d) {
        int thisVar = c.vars[index];
        if (index == numVar - 1) {
            if (thisVar == var) return validate(c);
            bool supported = false;
            for (int v = d.lb[thisVar * k + point], u = d.ub[thisVar * k + point]; !supported && v <= u; v++) {
                propagateValue[thisVar] = v;
                supported = validate(c);
            }
            return supported;
        }
        if (thisVar == var) return findSupportRe(c, var, point, index + 1, numVar, d);
        bool supported = false;
        for (int v = d.lb[thisVar * k + point], u = d.ub[thisVar * k + point]; !supported && v <= u; v++) {
            propagateValue[thisVar] = v;
            supported = findSupportRe(c, var, point, index + 1, numVar, d);
        }
        return supported;
    }
    bool findSupport(const Con &c, int var, int point, const Dom &d) {
        return findSupportRe(c, var, point, 0, (int)c.vars.size(), d);
    }

    /* ---- bounds revision of one arc: enforcePointConsistencyAt, src/solveralgorithm.cpp:476-523 */
    bool pointAt(const Con &c, int var, bool &change, int point, Dom &d) {
        bool supported = false;
        int lb = d.lb[var * k + point], ub = d.ub[var * k + point];
        for (int v = lb; !supported && v <= ub; v++) {
            propagateValue[var] = v;
            supported = findSupport(c, var, point, d);
        }
        if (supported) {
            if (lb != propagateValue[var]) { change = true; d.lb[var * k + point] = propagateValue[var]; }
            supported = false;
            lb = d.lb[var * k + point];
            for (int v = ub; !supported && v > lb; v--) {
                propagateValue[var] = v;
                supported = findSupport(c, var, point, d);
            }
            if (!supported) {
                if (ub != d.lb[var * k + point]) { change = true; d.ub[var * k + point] = d.lb[var * k + point]; }
                supported = true;
            } else if (ub != propagateValue[var]) {
                change = true;
                d.ub[var * k + point] = propagateValue[var];
            }
        }
        return supported;
    }
    bool enforcePoint(const Con &c, int var, bool &change, Dom &d) {      /* :527-539 */
        if (c.hasFirst) return pointAt(c, var, change, 0, d);
        bool consistent = true;
        for (int pt = 0; consistent && pt < k; pt++) consistent = pointAt(c, var, change, pt, d);
        return consistent;
    }
    bool enforceNext(const Con &c, int var, bool &change, Dom &d) {       /* :544-593 */
        int x = c.root->a->arg, y = c.root->b->a->arg;
        bool consistent = true;
        if (var == y) {
            for (int pt = 1; consistent && pt < k; pt++) {
                if (d.lb[y * k + pt] < d.lb[x * k + pt - 1]) { change = true; d.lb[y * k + pt] = d.lb[x * k + pt - 1]; }
                if (d.ub[y * k + pt] > d.ub[x * k + pt - 1]) { change = true; d.ub[y * k + pt] = d.ub[x * k + pt - 1]; }
                if (d.lb[y * k + pt] > d.ub[y * k + pt]) consistent = false;
            }
        } else {
            for (int pt = 0; consistent && pt < k - 1; pt++) {
                if (d.lb[x * k + pt] < d.lb[y * k + pt + 1]) { change = true; d.lb[x * k + pt] = d.lb[y * k + pt + 1]; }
                if (d.ub[x * k + pt] > d.ub[y * k + pt + 1]) { change = true; d.ub[x * k + pt] = d.ub[y * k + pt + 1]; }
                if (d.lb[x * k + pt] > d.ub[x * k + pt]) consistent = false;
            }
        }
        return consistent;
    }
    bool enforceUntil(const Con &c, int expired, const Dom &d) {          /* :598-614 */
        int l = c.root->a->arg, r = c.root->b->arg;
        if (expired) return true;
        if (d.lb[l * k] == d.ub[l * k] && d.lb[r * k] == d.ub[r * k])
            if (d.lb[r * k] != 1 && d.lb[l * k] != 1) return false;
        return true;
    }

    /* ---- propagation to fixpoint: generalisedArcConsistent, src/solveralgorithm.cpp:617-706 */
    bool gac(const CSet &q, const std::vector<int> &expire, Dom &d) {
        st.gac_calls++;
        struct ArcRef { int con, var; };
        std::deque<ArcRef> queue;
        std::vector<std::vector<char> > inq(q.cons.size());
        for (size_t c = 0; c < q.cons.size(); c++) {
            inq[c].assign(q.cons[c].arcs.size(), 1);
            for (size_t a = 0; a < q.cons[c].arcs.size(); a++) queue.push_back(ArcRef{(int)c, (int)a});
        }
        bool consistent = true;
        int untilIndex;
        while (consistent && !queue.empty()) {
            ArcRef ar = queue.front();
            queue.pop_front();
            inq[ar.con][ar.var] = 0;
            const Con &c = q.cons[ar.con];
            int var = c.arcs[ar.var];
            bool change = false;
            st.revisions++;
            if (c.kind == K_NEXT) consistent = enforceNext(c, var, change, d);
            else if (c.kind == K_POINT) consistent = enforcePoint(c, var, change, d);
            else if (c.kind == K_UNTIL) {
                untilIndex = 0;
                for (int j = 0; j < ar.con; j++) untilIndex += q.cons[j].kind == K_UNTIL;
                consistent = enforceUntil(c, expire[untilIndex], d);
            }
            if (consistent && change) {
                const std::vector<int> &cs = q.varCons[var];
                for (size_t i = 0; i < cs.size(); i++) {
                    if (cs[i] == ar.con) continue;
                    for (size_t a = 0; a < q.cons[cs[i]].arcs.size(); a++)
                        if (!inq[cs[i]][a]) { inq[cs[i]][a] = 1; queue.push_back(ArcRef{cs[i], (int)a}); }
                }
            }
        }
        return consistent;
    }

    int firstUnbound(const Dom &d) {                      /* solverGetFirstUnboundVar, src/solver.cpp:41-53 */
        for (int v = 0; v < V; v++)
            if (d.lb[v * k] < d.ub[v * k]) return v;
        return -1;
    }

    bool out_of_time() {
        if (deadline > 0 && (st.gac_calls & 63) == 0) {
            double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            if (now > deadline) timed_out = true;
        }
        return timed_out;
    }

    /* ---- the search: solverSolveRe, src/solveralgorithm.cpp:733-942 */
    bool solveRe(int vertex, int cset, const std::vector<int> &expire, const Dom &d, int depth) {
        if (depth > st.max_depth) st.max_depth = depth;
        if (out_of_time()) return false;
        bool ok = false;
        int var = firstUnbound(d);
        if (var < 0) {
            st.leaves++;
            int ncset = cset;
            if (solverHasFirst) {                         /* :755-805 */
                CSet q;
                q.varCons.resize(V);
                for (size_t c = 0; c < seen[cset].cons.size(); c++) {
                    NodeP t = translate(seen[cset].cons[c].root, d);
                    if (!tautology(t)) push(q, t);
                }
                bool found = false;
                for (size_t s = 0; !found && s < seen.size(); s++) {
                    if (seen[s].cons.size() != q.cons.size()) continue;
                    bool eq = true;
                    for (size_t c = 0; eq && c < q.cons.size(); c++) eq = node_eq(q.cons[c].root, seen[s].cons[c].root);
                    if (eq) { ncset = (int)s; found = true; }
                }
                if (!found) { ncset = (int)seen.size(); seen.push_back(q); }
            }
            std::vector<int> sig;                         /* :810-837 */
            for (int v = 0; v < V; v++)
                if (isSignature[v]) sig.push_back(d.lb[v * k]);
            std::vector<int> nexpire;
            {
                const CSet &nq = seen[ncset];
                size_t u = 0;
                for (size_t c = 0; c < nq.cons.size(); c++) {
                    if (nq.cons[c].kind != K_UNTIL) continue;
                    int flag;
                    if (u < expire.size() && expire[u] == 1) flag = 1;
                    else if (d.lb[nq.cons[c].root->b->arg * k] == 1) flag = 1;
                    else flag = 0;
                    nexpire.push_back(flag);
                    sig.push_back(flag);
                    u++;
                }
            }
            std::pair<int, std::vector<int> > key(ncset, sig);
            std::map<std::pair<int, std::vector<int> >, int>::iterator it = table.find(key);
            int temp;
            if (it == table.end()) {                      /* :842-864 */
                temp = (int)vertices.size();
                Vertex nv;
                nv.cset = ncset; nv.sig = sig; nv.fail = false;
                vertices.push_back(nv);
                table[key] = temp;
                Dom nd = d;                               /* variableAdvanceOneTimeStep, src/variable.cpp:94-108 */
                for (int v = 0; v < V; v++) {
                    for (int i = 0; i < k - 1; i++) { nd.lb[v * k + i] = d.lb[v * k + i + 1]; nd.ub[v * k + i] = d.ub[v * k + i + 1]; }
                    nd.lb[v * k + k - 1] = p->var_lb[v];
                    nd.ub[v * k + k - 1] = p->var_ub[v];
                }
                if (gac(seen[ncset], nexpire, nd)) ok = solveRe(temp, ncset, nexpire, nd, depth + 1);
                else { numFails++; ok = false; }
            } else if (vertices[it->second].fail) {
                temp = it->second;
                ok = false;
            } else {
                temp = it->second;
                numDominance++;
                ok = true;
            }
            if (timed_out) return false;
            if (ok) {                                     /* edgeNew, src/graph.cpp:78-89 */
                esrc.push_back(vertex);
                edst.push_back(temp);
                for (int v = 0; v < V; v++) elabel.push_back(d.lb[v * k]);
            } else {
                vertices[temp].fail = true;
            }
        } else {
            st.splits++;
            int lo = d.lb[var * k], hi = d.ub[var * k];
            int mid = lo + (hi - lo) / 2;                 /* variableSplitLower/Upper, src/variable.cpp:52-67 */
            {
                Dom nd = d;
                nd.ub[var * k] = mid;
                if (gac(seen[cset], expire, nd)) ok |= solveRe(vertex, cset, expire, nd, depth + 1);
                else numFails++;
            }
            if (timed_out) return false;
            {
                Dom nd = d;
                nd.lb[var * k] = mid + 1;
                if (gac(seen[cset], expire, nd)) ok |= solveRe(vertex, cset, expire, nd, depth + 1);
                else numFails++;
            }
        }
        return ok;
    }

    /* ---- entry: solverSolve, src/solveralgorithm.cpp:945-1005 (up to graphTraverse) */
    void run() {
        V = p->n_vars;
        k = p->prefix_k > 0 ? p->prefix_k : 2;
        arrays.resize(p->n_arrays);
        for (int a = 0; a < p->n_arrays; a++)
            arrays[a].assign(p->arr_values + p->arr_offsets[a], p->arr_values + p->arr_offsets[a + 1]);
        isSignature.assign(V, 0);
        untilSeen.assign(V, 0);
        propagateValue.assign(V, 0);
        numSignVar = numUntilVars = 0;
        solverHasFirst = false;
        numFails = numDominance = 0;
        memset(&st, 0, sizeof st);
        timed_out = false;
        CSet q0;
        q0.varCons.resize(V);
        for (int c = 0; c < p->n_constraints; c++)
            push(q0, build(p->con_tokens + p->con_offsets[c], p->con_offsets[c + 1] - p->con_offsets[c]));
        seen.push_back(q0);
        Vertex root;
        root.cset = 0; root.fail = false;
        vertices.push_back(root);
        table[std::make_pair(0, std::vector<int>())] = 0;
        Dom d;
        d.lb.resize(V * k); d.ub.resize(V * k);
        for (int v = 0; v < V; v++)
            for (int i = 0; i < k; i++) { d.lb[v * k + i] = p->var_lb[v]; d.ub[v * k + i] = p->var_ub[v]; }
        std::vector<int> expire;
        for (size_t c = 0; c < q0.cons.size(); c++)
            if (q0.cons[c].kind == K_UNTIL) expire.push_back(0);
        try {
            if (gac(seen[0], expire, d)) solveRe(0, 0, expire, d, 1);
            else numFails++;
        } catch (const TimeUp &) {
            timed_out = true;
        }
    }
};

struct ThreadArg {
    Oracle *o;
    std::string err;
};

void *thread_main(void *arg) {
    ThreadArg *ta = (ThreadArg *)arg;
    try { ta->o->run(); } catch (const std::exception &e) { ta->err = e.what(); }
    return NULL;
}

thread_local std::string g_err;

template <class T>
T *dup_vec(const std::vector<T> &v) {
    T *p = (T *)malloc(sizeof(T) * (v.size() ? v.size() : 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

}  // namespace

extern "C" {

typedef struct stcsp_oracle_stats {
    int64_t num_nodes;        /* stat line "nodes": vertex-table size incl. failed states */
    int64_t num_fails;        /* stat line "fails" */
    int64_t num_dominance;    /* stat line "dominance" */
    int64_t gac_calls;        /* search nodes */
    int64_t validates, node_visits, revisions, leaves, splits, max_depth;
    int64_t constraint_sets;
    double solve_s;           /* wall seconds of the search alone */
    int32_t timed_out;
} stcsp_oracle_stats_t;

const char *stcsp_oracle_last_error(void) { return g_err.c_str(); }

/* Same contract as stcsp_gpu_solve (include/stcsp_b200.h); time_limit_s <= 0 means none. */
int stcsp_oracle_solve(const stcsp_problem_t *problem, double time_limit_s, stcsp_automaton_t *out,
                       stcsp_oracle_stats_t *stats) {
    Oracle o;
    o.p = problem;
    o.deadline = 0;
    if (time_limit_s > 0)
        o.deadline = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() + time_limit_s;
    /* The reference recurses once per search node along a path of the automaton (25 675 frames on
     * digitinvader8) and lifts RLIMIT_STACK for it (src/stcsp.y:184-191); run on a big stack. */
    ThreadArg ta;
    ta.o = &o;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_t th;
    auto t0 = std::chrono::steady_clock::now();
    bool started = false;
    for (int shift = 33; !started && shift >= 26; shift -= 2) {
        pthread_attr_setstacksize(&attr, (size_t)1 << shift);
        started = pthread_create(&th, &attr, thread_main, &ta) == 0;
    }
    if (!started) { g_err = "pthread_create failed"; return STCSP_ERR_INVALID; }
    pthread_join(th, NULL);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!ta.err.empty()) { g_err = ta.err; return STCSP_ERR_INVALID; }

    if (stats) {
        stats->num_nodes = (int64_t)o.vertices.size();
        stats->num_fails = o.numFails;
        stats->num_dominance = o.numDominance;
        stats->gac_calls = o.st.gac_calls;
        stats->validates = o.st.validates;
        stats->node_visits = o.st.node_visits;
        stats->revisions = o.st.revisions;
        stats->leaves = o.st.leaves;
        stats->splits = o.st.splits;
        stats->max_depth = o.st.max_depth;
        stats->constraint_sets = (int64_t)o.seen.size();
        stats->solve_s = secs;
        stats->timed_out = o.timed_out;
    }
    if (o.timed_out) { g_err = "time limit"; return STCSP_ERR_TIMEOUT; }
    if (!out) return STCSP_OK;

    memset(out, 0, sizeof *out);
    int V = o.V;
    std::vector<int> sigVars;
    for (int v = 0; v < V; v++)
        if (o.isSignature[v]) sigVars.push_back(v);
    int nUntil = 0;
    for (size_t c = 0; c < o.seen[0].cons.size(); c++) nUntil += o.seen[0].cons[c].kind == K_UNTIL;
    out->n_vars = V;
    out->n_sig_vars = (int32_t)sigVars.size();
    out->n_until = nUntil;
    out->n_until_vars = o.numUntilVars;
    out->sig_len = out->n_sig_vars + nUntil;
    out->sig_vars = dup_vec(sigVars);
    out->root_final = nUntil == 0;
    out->n_constraint_sets = (int32_t)o.seen.size();
    out->n_states = (int64_t)o.vertices.size();
    std::vector<int32_t> sig((size_t)out->n_states * out->sig_len, 0), cset(out->n_states);
    std::vector<uint8_t> failed(out->n_states);
    for (int64_t s = 0; s < out->n_states; s++) {
        cset[s] = o.vertices[s].cset;
        failed[s] = o.vertices[s].fail;
        for (size_t i = 0; i < o.vertices[s].sig.size(); i++) sig[s * out->sig_len + i] = o.vertices[s].sig[i];
    }
    out->state_sig = dup_vec(sig);
    out->state_cset = dup_vec(cset);
    out->state_failed = dup_vec(failed);
    /* edges sorted by (src, label) */
    size_t m = o.esrc.size();
    std::vector<size_t> order(m);
    for (size_t i = 0; i < m; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
        if (o.esrc[x] != o.esrc[y]) return o.esrc[x] < o.esrc[y];
        return std::lexicographical_compare(o.elabel.begin() + x * V, o.elabel.begin() + (x + 1) * V,
                                            o.elabel.begin() + y * V, o.elabel.begin() + (y + 1) * V);
    });
    std::vector<int32_t> es(m), ed(m), el(m * V);
    for (size_t i = 0; i < m; i++) {
        es[i] = o.esrc[order[i]];
        ed[i] = o.edst[order[i]];
        if (o.vertices[ed[i]].fail) { g_err = "oracle invariant broken: edge into a failed state"; return STCSP_ERR_INVALID; }
        memcpy(&el[i * V], &o.elabel[order[i] * V], sizeof(int32_t) * V);
    }
    out->n_edges = (int64_t)m;
    out->edge_src = dup_vec(es);
    out->edge_dst = dup_vec(ed);
    out->edge_label = dup_vec(el);
    out->n_search_nodes = o.st.gac_calls;
    out->n_fails = o.numFails;
    out->n_leaves = o.st.leaves;
    out->n_dominance = o.numDominance;
    out->n_tuples = o.st.validates;
    out->solve_ms = out->wall_ms = secs * 1e3;
    return STCSP_OK;
}

void stcsp_oracle_automaton_free(stcsp_automaton_t *a) {
    if (!a) return;
    free(a->sig_vars); free(a->state_sig); free(a->state_cset); free(a->state_failed);
    free(a->edge_src); free(a->edge_dst); free(a->edge_label);
    memset(a, 0, sizeof *a);
}

}  // extern "C"
