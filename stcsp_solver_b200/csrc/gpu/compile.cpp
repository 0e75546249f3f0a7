// Constraint-set table: structural identity, rewriting at time advance, compilation to device tables.
#include "compile.h"

#include <algorithm>
#include <stdexcept>

namespace stcsp {

namespace {

constexpr int kScalarScope = 4;     // widest scope a single thread revises (see scalar_revise in kernels.cu)

struct Emitter {
    const std::vector<int32_t> &scope;
    std::vector<Instr> &out;
    size_t base;

    int32_t here() const { return (int32_t)(out.size() - base); }
    int32_t emit(int32_t op, int32_t arg = 0) {
        out.push_back(Instr{op, arg});
        return (int32_t)out.size() - 1;
    }
    void patch(int32_t at) { out[at].arg = here(); }
    int32_t slot(int32_t var) const {
        for (size_t i = 0; i < scope.size(); i++)
            if (scope[i] == var) return (int32_t)i;
        throw std::runtime_error("internal: variable not in constraint scope");
    }

    // returns the stack depth needed by the subtree (it leaves exactly one value)
    int gen(const Expr &e) {
        switch (e.op) {
            case STCSP_OP_CONST: emit(BC_PUSHC, e.arg); return 1;
            case STCSP_OP_VAR: emit(BC_PUSHV, slot(e.arg)); return 1;
            case STCSP_OP_FIRST: case STCSP_OP_AT: return gen(*e.kid[0]);      // transparent (:364-367)
            case STCSP_OP_NEXT: emit(BC_PUSHC, 0); return 1;                     // a stray next evaluates to 0
            case STCSP_OP_ARR: { int d = gen(*e.kid[0]); emit(BC_ARR, e.arg); return d; }
            case STCSP_OP_ABS: { int d = gen(*e.kid[0]); emit(BC_ABS); return d; }
            case STCSP_OP_NOT: { int d = gen(*e.kid[0]); emit(BC_NOT); return d; }
            case STCSP_OP_IF: {
                int d = gen(*e.kid[0]);
                int32_t jz = emit(BC_JZ);
                d = std::max(d, gen(*e.kid[1]));
                int32_t jmp = emit(BC_JMP);
                patch(jz);
                d = std::max(d, gen(*e.kid[2]));
                patch(jmp);
                return d;
            }
            case STCSP_OP_AND: case STCSP_OP_OR: {
                int d = gen(*e.kid[0]);
                int32_t sc = emit(e.op == STCSP_OP_AND ? BC_AND_SC : BC_OR_SC);
                d = std::max(d, gen(*e.kid[1]));
                patch(sc);
                return d;
            }
            case STCSP_CON_IMPLY: {
                int d = gen(*e.kid[0]);
                int32_t sc = emit(BC_IMPLY_SC);
                d = std::max(d, 1 + gen(*e.kid[1]));
                emit(BC_IMPLY_FIN);
                patch(sc);
                return d;
            }
            default: break;
        }
        int32_t op;
        switch (e.op) {
            case STCSP_OP_LT: case STCSP_CON_LT: op = BC_LT; break;
            case STCSP_OP_GT: case STCSP_CON_GT: op = BC_GT; break;
            case STCSP_OP_LE: case STCSP_CON_LE: op = BC_LE; break;
            case STCSP_OP_GE: case STCSP_CON_GE: op = BC_GE; break;
            case STCSP_OP_EQ: case STCSP_CON_EQ: op = BC_EQ; break;
            case STCSP_OP_NE: case STCSP_CON_NE: op = BC_NE; break;
            case STCSP_OP_ADD: op = BC_ADD; break;
            case STCSP_OP_SUB: op = BC_SUB; break;
            case STCSP_OP_MUL: op = BC_MUL; break;
            case STCSP_OP_DIV: op = BC_DIV; break;
            case STCSP_OP_MOD: op = BC_MOD; break;
            default: throw std::runtime_error("unsupported operator inside a pointwise constraint");
        }
        int d = gen(*e.kid[0]);
        d = std::max(d, 1 + gen(*e.kid[1]));
        emit(op);
        return d;
    }
};

void collect_vars(const Expr &e, std::vector<int32_t> &out) {
    if (e.op == STCSP_OP_VAR) {
        if (std::find(out.begin(), out.end(), e.arg) == out.end()) out.push_back(e.arg);
        return;
    }
    for (const auto &k : e.kid) collect_vars(*k, out);
}

void scan_first(const Expr &e, bool &has_first, bool &has_at, std::vector<int32_t> &cap) {
    if (e.op == STCSP_OP_FIRST) {
        has_first = true;
        collect_vars(*e.kid[0], cap);
        return;
    }
    if (e.op == STCSP_OP_AT) has_at = true;
    for (const auto &k : e.kid) scan_first(*k, has_first, has_at, cap);
}

// reference constraintNodeTranslateFirst, src/constraint.cpp:466-481
ExprPtr substitute(const Expr &e, const int32_t *values) {
    if (e.op == STCSP_OP_VAR) return mk(STCSP_OP_CONST, values[e.arg]);
    auto n = mk(e.op, e.arg);
    for (const auto &k : e.kid) n->kid.push_back(substitute(*k, values));
    return n;
}

// reference constraintNodeTranslate / constraintNodeTranslateAT, src/constraint.cpp:483-537
ExprPtr translate(const Expr &e, const int32_t *values, const std::vector<Array> &arrays) {
    if (e.op == STCSP_OP_FIRST) {
        ExprPtr s = substitute(*e.kid[0], values);
        Lifted v = fold_value(*s, arrays);
        if (!v.unknown) return mk(STCSP_OP_CONST, v.value);
        return s;           // the reference logs "cannot be completely evaluated" and keeps the subtree
    }
    if (e.op == STCSP_CON_EQ && e.kid[1]->op == STCSP_OP_AT) {
        int32_t x = e.kid[0]->arg, y = e.kid[1]->kid[0]->arg, n = e.kid[1]->arg;
        if (n == 1) return mk2(STCSP_CON_EQ, mk(STCSP_OP_VAR, x), mk1(STCSP_OP_FIRST, mk(STCSP_OP_VAR, y)));
        return mk2(STCSP_CON_EQ, mk(STCSP_OP_VAR, x), mk1(STCSP_OP_AT, mk(STCSP_OP_VAR, y), n - 1));
    }
    auto n = mk(e.op, e.arg);
    for (const auto &k : e.kid) n->kid.push_back(translate(*k, values, arrays));
    return n;
}

}  // namespace

int compile_expr(const Expr &root, const std::vector<int32_t> &scope, std::vector<Instr> &out) {
    Emitter em{scope, out, out.size()};
    int depth = em.gen(root);
    em.emit(BC_END);
    return depth;
}

void SetTable::init(const stcsp_problem_t &p) {
    if (p.abi_version != STCSP_ABI_VERSION) throw std::runtime_error("abi_version mismatch");
    if (p.n_vars <= 0) throw std::runtime_error("problem has no variables");
    k_ = p.prefix_k > 0 ? p.prefix_k : 2;
    if (k_ > 4) throw std::invalid_argument("unsupported: prefix_k > 4");
    lb_.assign(p.var_lb, p.var_lb + p.n_vars);
    width_.resize(p.n_vars);
    for (int32_t v = 0; v < p.n_vars; v++) {
        int64_t w = (int64_t)p.var_ub[v] - p.var_lb[v] + 1;
        if (w < 1) throw std::runtime_error("variable with empty declared domain");
        if (w > Limits::kMaxWidth)
            throw std::invalid_argument("unsupported: variable " + std::to_string(v) + " has a domain of " +
                                        std::to_string(w) + " values (limit 64: one bitset word per domain)");
        width_[v] = (int32_t)w;
    }
    arrays_.resize(p.n_arrays);
    arr_off.assign(1, 0);
    for (int32_t a = 0; a < p.n_arrays; a++) {
        arrays_[a].elements.assign(p.arr_values + p.arr_offsets[a], p.arr_values + p.arr_offsets[a + 1]);
        arr_val.insert(arr_val.end(), arrays_[a].elements.begin(), arrays_[a].elements.end());
        arr_off.push_back((int32_t)arr_val.size());
    }
    std::vector<Constraint> cons;
    is_sig_.assign(p.n_vars, 0);
    std::vector<uint8_t> until_var(p.n_vars, 0);
    for (int32_t c = 0; c < p.n_constraints; c++) {
        Constraint con;
        con.root = unflatten(p.con_tokens + p.con_offsets[c], p.con_offsets[c + 1] - p.con_offsets[c]);
        if (!is_constraint_op(con.root->op)) throw std::runtime_error("constraint " + std::to_string(c) + " does not end in a constraint operator");
        classify(con);
        for (int32_t v : con.scope)
            if (v < 0 || v >= p.n_vars) throw std::runtime_error("variable index out of range");
        if (con.kind == ConKind::Next) {
            if (con.root->kid[0]->op != STCSP_OP_VAR || con.root->kid[1]->kid[0]->op != STCSP_OP_VAR)
                throw std::runtime_error("malformed `x == next y` constraint");
            is_sig_[con.root->kid[0]->arg] = 1;
        } else if (con.kind == ConKind::Until) {
            if (con.root->kid[0]->op != STCSP_OP_VAR || con.root->kid[1]->op != STCSP_OP_VAR)
                throw std::runtime_error("malformed `x until y` constraint");
            n_until_++;
            if (!until_var[con.root->kid[1]->arg]) { until_var[con.root->kid[1]->arg] = 1; n_until_vars_++; }
        } else if (con.kind == ConKind::At) {
            if (con.root->kid[0]->op != STCSP_OP_VAR || con.root->kid[1]->kid[0]->op != STCSP_OP_VAR)
                throw std::runtime_error("malformed `x == y@n` constraint");
        }
        cons.push_back(std::move(con));
    }
    if (n_until_ > Limits::kMaxUntil) throw std::invalid_argument("unsupported: more than 30 until constraints");
    for (int32_t v = 0; v < p.n_vars; v++)
        if (is_sig_[v]) sig_vars_.push_back(v);
    add_set(std::move(cons));
}

int32_t SetTable::add_set(std::vector<Constraint> cons) {
    int32_t s = (int32_t)sets_.size();
    sets_.emplace_back();
    sets_[s].cons = std::move(cons);
    dev_sets.emplace_back();
    compile_set(s);
    resolve_static(s);
    dirty_ = true;
    return s;
}

int32_t SetTable::find_or_add(std::vector<Constraint> cons) {
    for (size_t s = 0; s < sets_.size(); s++) {          // reference constraintQueueEq
        const auto &have = sets_[s].cons;
        if (have.size() != cons.size()) continue;
        bool eq = true;
        for (size_t c = 0; eq && c < cons.size(); c++) eq = have[c].root->equals(*cons[c].root);
        if (eq) return (int32_t)s;
    }
    return add_set(std::move(cons));
}

void SetTable::compile_set(int32_t s) {
    HostSet &hs = sets_[s];
    DevSet ds{};
    const int32_t V = n_vars();
    ds.prop_off = (int32_t)dev_props.size();
    ds.con_off = (int32_t)dev_cons.size();
    ds.scope_off = (int32_t)dev_scope.size();
    ds.code_off = (int32_t)dev_code.size();
    ds.until_off = (int32_t)dev_aux.size();
    std::vector<int32_t> until_right, next_pairs;
    int32_t until_idx = 0;
    for (auto &c : hs.cons) {
        scan_first(*c.root, hs.has_first, hs.has_at, hs.cap_vars);
        if (c.kind == ConKind::At) continue;              // lazy: never enforced (src/solveralgorithm.cpp:658-662)
        if (c.kind == ConKind::Until) until_right.push_back(c.root->kid[1]->arg);
        if (c.scope.empty()) continue;                    // no arcs, never revised
        DevCon dc{};
        int32_t ci = (int32_t)dev_cons.size();
        if (c.kind == ConKind::Next) {
            dc.kind = DK_NEXT;
            dc.pivot = -1;
            dc.x = c.root->kid[0]->arg;
            dc.y = c.root->kid[1]->kid[0]->arg;
            next_pairs.push_back(dc.x);
            next_pairs.push_back(dc.y);
            dev_cons.push_back(dc);
            dev_props.push_back(DevProp{ci, 0});
        } else if (c.kind == ConKind::Until) {
            dc.kind = DK_UNTIL;
            dc.pivot = -1;
            dc.x = c.root->kid[0]->arg;
            dc.y = c.root->kid[1]->arg;
            dc.until_idx = until_idx;
            dev_cons.push_back(dc);
            dev_props.push_back(DevProp{ci, 0});
        } else {
            if ((int)c.scope.size() > Limits::kMaxScope)
                throw std::invalid_argument("unsupported: a constraint mentions more than 32 variables");
            dc.kind = DK_POINT;
            dc.n_scope = (int32_t)c.scope.size();
            dc.scope_off = (int32_t)dev_scope.size();
            dev_scope.insert(dev_scope.end(), c.scope.begin(), c.scope.end());
            dc.code_off = (int32_t)dev_code.size();
            int depth = compile_expr(*c.root, c.scope, dev_code);
            dc.code_len = (int32_t)dev_code.size() - dc.code_off;
            assign_table(dc, c);
            if (depth > Limits::kMaxStack) throw std::invalid_argument("unsupported: expression nesting needs a deeper evaluator stack");
            max_stack_ = std::max(max_stack_, (int32_t)depth);
            ds.max_stack = std::max(ds.max_stack, (int32_t)depth);
            max_scope_ = std::max(max_scope_, dc.n_scope);
            dev_cons.push_back(dc);
            int32_t offsets = c.has_first ? 1 : k_;      // src/solveralgorithm.cpp:527-539
            for (int32_t o = 0; o < offsets; o++) dev_props.push_back(DevProp{ci, o});
        }
        if (c.kind == ConKind::Until) until_idx++;
    }
    // until constraints without variables cannot exist; count every until for the flag layout
    ds.n_until = (int32_t)until_right.size();
    dev_aux.insert(dev_aux.end(), until_right.begin(), until_right.end());
    ds.next_off = (int32_t)dev_aux.size();
    ds.n_next = (int32_t)next_pairs.size() / 2;
    dev_aux.insert(dev_aux.end(), next_pairs.begin(), next_pairs.end());
    ds.n_cap = (int32_t)hs.cap_vars.size();
    ds.cap_off = (int32_t)dev_aux.size();
    dev_aux.insert(dev_aux.end(), hs.cap_vars.begin(), hs.cap_vars.end());

    ds.n_prop = (int32_t)dev_props.size() - ds.prop_off;
    ds.n_con = (int32_t)dev_cons.size() - ds.con_off;
    ds.n_scope = (int32_t)dev_scope.size() - ds.scope_off;
    ds.n_code = (int32_t)dev_code.size() - ds.code_off;
    // the propagation loop runs the lowest dirty index first: cheap propagators before expensive ones
    // A wide scope is scanned faster by 32 lanes than by one thread -- unless many wide tables are dirty together, in
    // which case one thread each, side by side, wins (digitinvader: 22 of them; partial order: 2).
    int wide_tables = 0;
    for (int32_t q = ds.prop_off; q < (int32_t)dev_props.size(); q++) {
        const DevCon &dc = dev_cons[dev_props[q].con];
        wide_tables += dc.kind == DK_POINT && dc.pivot >= 0 && dc.n_scope > kScalarScope;
    }
    const bool wide_scalar = wide_tables >= 6;
    auto cost = [&](const DevProp &pr) -> long long {
        const DevCon &dc = dev_cons[pr.con];
        if (dc.kind != DK_POINT) return 0;
        // scalar class first (cheapest table first), then the cooperative tables, then the bytecode enumerations
        if (dc.pivot >= 0)
            return (dc.n_scope <= kScalarScope || wide_scalar) ? 1 + dc.table_entries : (1ll << 39) + dc.table_entries;
        return (1ll << 40) + dc.n_scope;
    };
    std::stable_sort(dev_props.begin() + ds.prop_off, dev_props.end(),
                     [&](const DevProp &a, const DevProp &b) { return cost(a) < cost(b); });
    ds.n_cheap = 0;
    for (int32_t q = 0; q < ds.n_prop; q++)
        if (cost(dev_props[ds.prop_off + q]) < (1ll << 39)) ds.n_cheap = q + 1;
    ds.n_words = std::max(1, (ds.n_prop + 31) / 32);
    max_props_ = std::max(max_props_, ds.n_prop);
    ds.wake_off = (int32_t)dev_wake.size();
    // V*k rows of wake masks, then one more row: the pointwise propagators at look-ahead offsets (>= 1)
    dev_wake.resize(dev_wake.size() + ((size_t)V * k_ + 1) * ds.n_words, 0u);
    auto wake = [&](int32_t var, int32_t off, int32_t q) {
        dev_wake[ds.wake_off + ((size_t)var * k_ + off) * ds.n_words + q / 32] |= 1u << (q % 32);
    };
    for (int32_t q = 0; q < ds.n_prop; q++) {
        const DevProp &pr = dev_props[ds.prop_off + q];
        const DevCon &dc = dev_cons[pr.con];
        if (dc.kind == DK_POINT && pr.offset >= 1)
            dev_wake[ds.wake_off + (size_t)V * k_ * ds.n_words + q / 32] |= 1u << (q % 32);
        if (dc.kind == DK_POINT) {
            for (int32_t i = 0; i < dc.n_scope; i++) wake(dev_scope[dc.scope_off + i], pr.offset, q);
        } else if (dc.kind == DK_NEXT) {
            for (int32_t o = 0; o + 1 < k_; o++) { wake(dc.x, o, q); wake(dc.y, o + 1, q); }
        } else {
            wake(dc.x, 0, q);
            wake(dc.y, 0, q);
        }
    }
    ds.static_next = -1;
    dev_sets[s] = ds;
}

// Relation table of a pointwise constraint: allocated here, filled on the device.
void SetTable::assign_table(DevCon &dc, const Constraint &c) {
    dc.pivot = -1;
    dc.table_entries = 0;
    dc.table_off = 0;
    const int n = dc.n_scope;
    dev_stride.resize(dev_scope.size(), 0);
    if (n < 1) return;
    std::vector<stcsp_tok_t> toks;
    flatten_expr(*c.root, toks);
    std::string key((const char *)toks.data(), toks.size() * sizeof(stcsp_tok_t));
    auto it = table_cache_.find(key);
    if (it == table_cache_.end()) {
        int pivot = 0;
        for (int i = 1; i < n; i++)
            if (width_[c.scope[i]] > width_[c.scope[pivot]]) pivot = i;
        long long entries = 1;
        for (int i = 0; i < n && entries <= Limits::kMaxTableEntries; i++)
            if (i != pivot) entries *= width_[c.scope[i]];
        if (entries > Limits::kMaxTableEntries || table_words + entries > Limits::kMaxTableWords) return;
        TableRef ref;
        ref.off = table_words;
        ref.entries = (int32_t)entries;
        ref.pivot = pivot;
        ref.strides.assign(n, 0);
        int32_t stride = 1;
        for (int i = 0; i < n; i++) {
            if (i == pivot) continue;
            ref.strides[i] = stride;
            stride *= width_[c.scope[i]];
        }
        table_words += entries;
        it = table_cache_.emplace(key, ref).first;
        dc.pivot = ref.pivot;
        dc.table_entries = ref.entries;
        dc.table_off = ref.off;
        for (int i = 0; i < n; i++) dev_stride[dc.scope_off + i] = ref.strides[i];
        table_jobs.push_back(TableJob{(int32_t)dev_cons.size()});      // dc is appended right after this call
        return;
    }
    const TableRef &ref = it->second;
    dc.pivot = ref.pivot;
    dc.table_entries = ref.entries;
    dc.table_off = ref.off;
    for (int i = 0; i < n; i++) dev_stride[dc.scope_off + i] = ref.strides[i];
}

size_t SetTable::max_stage_bytes() const {
    auto a8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
    size_t best = 0;
    for (const DevSet &ds : dev_sets) {
        size_t b = a8((size_t)ds.n_prop * sizeof(DevProp)) + a8((size_t)ds.n_con * sizeof(DevCon)) +
                   2 * a8((size_t)ds.n_scope * 4) + a8(((size_t)n_vars() * k_ + 1) * ds.n_words * 4) + 2 * a8((size_t)n_vars() * 4) +
                   a8((size_t)ds.n_code * sizeof(Instr));
        best = std::max(best, b);
    }
    return best;
}

void SetTable::resolve_static(int32_t s) {
    if (!sets_[s].has_first && !sets_[s].has_at) {
        sets_[s].static_next = s;                         // rewriting is the identity
    } else if (sets_[s].cap_vars.empty()) {
        std::vector<int32_t> zeros(n_vars(), 0);          // no variable is read
        int32_t nxt = successor(s, zeros.data());
        sets_[s].static_next = nxt;
    }
    dev_sets[s].static_next = sets_[s].static_next;
}

int32_t SetTable::successor(int32_t s, const int32_t *values) {
    std::vector<Constraint> out;
    const size_t n = sets_[s].cons.size();
    for (size_t c = 0; c < n; c++) {
        ExprPtr t = translate(*sets_[s].cons[c].root, values, arrays_);
        if (is_tautology(*t, arrays_)) continue;          // constraintTranslate, src/constraint.cpp:540-548
        Constraint con;
        con.root = std::move(t);
        classify(con);
        out.push_back(std::move(con));
    }
    return find_or_add(std::move(out));
}

}  // namespace stcsp
