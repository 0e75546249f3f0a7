// Expression helpers, constraint classification, constant folding and the flat C-ABI view.
#include "model.h"

#include <cstdlib>
#include <sstream>

namespace stcsp {

ExprPtr Expr::clone() const {
    auto e = std::make_unique<Expr>(op, arg);
    e->kid.reserve(kid.size());
    for (const auto &k : kid) e->kid.push_back(k->clone());
    return e;
}

// reference constraintNodeEq (src/constraint.cpp:551-561): token, num and var must agree; the
// array pointer is not compared there, but an array reference is never rewritten, so comparing
// `arg` for ARR nodes as well is equivalent.
bool Expr::equals(const Expr &o) const {
    if (op != o.op || arg != o.arg || kid.size() != o.kid.size()) return false;
    for (size_t i = 0; i < kid.size(); i++)
        if (!kid[i]->equals(*o.kid[i])) return false;
    return true;
}

ExprPtr mk(int32_t op, int32_t arg) { return std::make_unique<Expr>(op, arg); }
ExprPtr mk1(int32_t op, ExprPtr a, int32_t arg) {
    auto e = mk(op, arg);
    e->kid.push_back(std::move(a));
    return e;
}
ExprPtr mk2(int32_t op, ExprPtr a, ExprPtr b) {
    auto e = mk(op);
    e->kid.push_back(std::move(a));
    e->kid.push_back(std::move(b));
    return e;
}
ExprPtr mk3(int32_t op, ExprPtr a, ExprPtr b, ExprPtr c) {
    auto e = mk(op);
    e->kid.push_back(std::move(a));
    e->kid.push_back(std::move(b));
    e->kid.push_back(std::move(c));
    return e;
}

int arity_of(int32_t op) {
    switch (op) {
        case STCSP_OP_CONST: case STCSP_OP_VAR: return 0;
        case STCSP_OP_ARR: case STCSP_OP_ABS: case STCSP_OP_NOT: case STCSP_OP_FIRST:
        case STCSP_OP_NEXT: case STCSP_OP_AT: return 1;
        case STCSP_OP_IF: return 3;
        default:
            if ((op >= STCSP_OP_LT && op <= STCSP_OP_MOD) || (op >= STCSP_CON_LT && op <= STCSP_CON_UNTIL)) return 2;
            if (op == OP_FBY) return 2;
            return -1;
    }
}

bool is_constraint_op(int32_t op) { return op >= STCSP_CON_LT && op <= STCSP_CON_UNTIL; }

int32_t Model::find_var(const std::string &name) const {
    for (size_t i = 0; i < vars.size(); i++)
        if (vars[i].name == name) return (int32_t)i;
    return -1;
}
int32_t Model::find_array(const std::string &name) const {
    for (size_t i = 0; i < arrays.size(); i++)
        if (arrays[i].name == name) return (int32_t)i;
    return -1;
}
int32_t Model::add_var(const std::string &name, int32_t lb, int32_t ub) {
    if (lb > ub)   // reference variableNew, src/variable.cpp:15-18
        throw ParseError("Invalid domain [" + std::to_string(lb) + ", " + std::to_string(ub) + "] in variable " + name);
    vars.push_back(Variable{name, lb, ub});
    return (int32_t)vars.size() - 1;
}
int32_t Model::add_aux(int32_t lb, int32_t ub) {
    return add_var("_V" + std::to_string(n_aux++), lb, ub);
}

bool expr_has_first(const Expr &e) {   // reference constraintNodeHasFirst, src/constraint.cpp:240-250
    if (e.op == STCSP_OP_FIRST || e.op == STCSP_OP_AT) return true;
    for (const auto &k : e.kid)
        if (expr_has_first(*k)) return true;
    return false;
}

static void link_vars(const Expr &e, std::vector<int32_t> &scope) {
    if (e.op == STCSP_OP_VAR) {
        for (int32_t v : scope)
            if (v == e.arg) return;
        scope.push_back(e.arg);
        return;
    }
    for (const auto &k : e.kid) link_vars(*k, scope);
}

void classify(Constraint &c) {
    const Expr &r = *c.root;
    if (r.op == STCSP_CON_UNTIL) c.kind = ConKind::Until;
    else if (r.kid.size() == 2 && r.kid[1]->op == STCSP_OP_NEXT) c.kind = ConKind::Next;
    else if (r.kid.size() == 2 && r.kid[1]->op == STCSP_OP_AT) c.kind = ConKind::At;
    else c.kind = ConKind::Point;
    c.has_first = expr_has_first(r);
    c.scope.clear();
    link_vars(r, c.scope);
}

static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

Lifted fold_value(const Expr &e, const std::vector<Array> &arrays) {
    Lifted r;
    auto known = [](int32_t v) { Lifted x; x.unknown = false; x.value = v; return x; };
    switch (e.op) {
        case STCSP_OP_VAR: return r;
        case STCSP_OP_CONST: return known(e.arg);
        case STCSP_OP_NEXT: return r;
        case STCSP_OP_FIRST: return fold_value(*e.kid[0], arrays);
        case STCSP_OP_AT: {          // the reference leaves the value undefined here; it is never used
            return r;
        }
        case STCSP_OP_ARR: {
            Lifted i = fold_value(*e.kid[0], arrays);
            if (i.unknown) return r;
            const auto &el = arrays[e.arg].elements;
            if (i.value < 0 || i.value >= (int32_t)el.size()) return r;   // reference reads out of bounds
            return known(el[i.value]);
        }
        case STCSP_OP_ABS: {
            Lifted v = fold_value(*e.kid[0], arrays);
            if (v.unknown) return r;
            return known(v.value < 0 ? wrap_sub(0, v.value) : v.value);
        }
        case STCSP_OP_IF: {
            Lifted c = fold_value(*e.kid[0], arrays);
            if (c.unknown) return r;
            return fold_value(*e.kid[c.value ? 1 : 2], arrays);
        }
        case STCSP_OP_NOT: {
            Lifted v = fold_value(*e.kid[0], arrays);
            if (v.unknown) return r;
            return known(v.value == 0 ? 1 : 0);
        }
        case STCSP_OP_AND: {
            Lifted l = fold_value(*e.kid[0], arrays);
            if (l.unknown) return r;
            if (l.value == 0) return known(0);
            return fold_value(*e.kid[1], arrays);
        }
        case STCSP_OP_OR: {          // quirk kept: anything but exactly 1 on the left yields 1
            Lifted l = fold_value(*e.kid[0], arrays);
            if (l.unknown) return r;
            if (l.value != 1) return known(1);
            return fold_value(*e.kid[1], arrays);
        }
        default: break;
    }
    if (e.kid.size() != 2) return r;
    Lifted a = fold_value(*e.kid[0], arrays), b = fold_value(*e.kid[1], arrays);
    if (a.unknown || b.unknown) return r;
    switch (e.op) {
        case STCSP_OP_LT: return known(a.value < b.value);
        case STCSP_OP_GT: return known(a.value < b.value);    // quirk kept (src/constraint.cpp:425)
        case STCSP_OP_LE: return known(a.value <= b.value);
        case STCSP_OP_GE: return known(a.value >= b.value);
        case STCSP_OP_EQ: return known(a.value == b.value);
        case STCSP_OP_NE: return known(a.value != b.value);
        case STCSP_OP_ADD: return known(wrap_add(a.value, b.value));
        case STCSP_OP_SUB: return known(wrap_sub(a.value, b.value));
        case STCSP_OP_MUL: return known(wrap_mul(a.value, b.value));
        case STCSP_OP_DIV:
            if (b.value == 0 || (a.value == INT32_MIN && b.value == -1)) return r;   // reference traps
            return known(a.value / b.value);
        case STCSP_OP_MOD:
            if (b.value == 0 || (a.value == INT32_MIN && b.value == -1)) return r;
            return known(a.value % b.value);
        default: return r;           // constraint-level tokens: undefined in the reference
    }
}

bool is_tautology(const Expr &root, const std::vector<Array> &arrays) {   // src/constraint.cpp:443-462
    if (root.kid.size() != 2) return false;
    Lifted a = fold_value(*root.kid[0], arrays), b = fold_value(*root.kid[1], arrays);
    if (a.unknown || b.unknown) return false;
    switch (root.op) {
        case STCSP_CON_LT: return a.value < b.value;
        case STCSP_CON_GT: return a.value > b.value;
        case STCSP_CON_LE: return a.value <= b.value;
        case STCSP_CON_GE: return a.value >= b.value;
        case STCSP_CON_EQ: return a.value == b.value;
        case STCSP_CON_NE: return a.value != b.value;
        case STCSP_CON_IMPLY: return a.value <= b.value;
        case STCSP_CON_UNTIL: return b.value == 1;
        default: return false;
    }
}

static const char *op_text(int32_t op) {
    switch (op) {
        case STCSP_OP_ABS: return "abs"; case STCSP_OP_NOT: return "not";
        case STCSP_OP_FIRST: return "first"; case STCSP_OP_NEXT: return "next";
        case STCSP_OP_LT: return "lt"; case STCSP_OP_GT: return "gt"; case STCSP_OP_LE: return "le";
        case STCSP_OP_GE: return "ge"; case STCSP_OP_EQ: return "eq"; case STCSP_OP_NE: return "ne";
        case STCSP_OP_AND: return "and"; case STCSP_OP_OR: return "or";
        case STCSP_OP_ADD: return "+"; case STCSP_OP_SUB: return "-"; case STCSP_OP_MUL: return "*";
        case STCSP_OP_DIV: return "/"; case STCSP_OP_MOD: return "%";
        case STCSP_CON_LT: return "<"; case STCSP_CON_GT: return ">"; case STCSP_CON_LE: return "<=";
        case STCSP_CON_GE: return ">="; case STCSP_CON_EQ: return "=="; case STCSP_CON_NE: return "!=";
        case STCSP_CON_IMPLY: return "->"; case STCSP_CON_UNTIL: return "until";
        case OP_FBY: return "fby";
        default: return "?";
    }
}

std::string to_string(const Expr &e, const Model &m) {
    switch (e.op) {
        case STCSP_OP_CONST: return std::to_string(e.arg);
        case STCSP_OP_VAR: return m.vars[e.arg].name;
        case STCSP_OP_ARR: return m.arrays[e.arg].name + "[" + to_string(*e.kid[0], m) + "]";
        case STCSP_OP_AT: return "(" + to_string(*e.kid[0], m) + " @ " + std::to_string(e.arg) + ")";
        case STCSP_OP_IF:
            return "(if " + to_string(*e.kid[0], m) + " then " + to_string(*e.kid[1], m) + " else " +
                   to_string(*e.kid[2], m) + ")";
        default: break;
    }
    if (e.kid.size() == 1) return std::string(op_text(e.op)) + "(" + to_string(*e.kid[0], m) + ")";
    if (e.kid.size() == 2) {
        std::string s = to_string(*e.kid[0], m) + " " + op_text(e.op) + " " + to_string(*e.kid[1], m);
        return is_constraint_op(e.op) ? s : "(" + s + ")";
    }
    return "?";
}

std::string dump_model(const Model &m) {
    std::ostringstream os;
    os << "k " << m.prefix_k << "\n";
    for (const auto &v : m.vars) os << "var " << v.name << " [" << v.lb << ", " << v.ub << "]\n";
    for (const auto &a : m.arrays) {
        os << "arr " << a.name << " {";
        for (size_t i = 0; i < a.elements.size(); i++) os << (i ? ", " : "") << a.elements[i];
        os << "}\n";
    }
    static const char *kinds[] = {"NEXT", "POINT", "UNTIL", "AT"};
    for (const auto &c : m.cons) {
        os << kinds[(int)c.kind] << (c.has_first ? "* " : " ") << to_string(*c.root, m) << " ; scope";
        for (int32_t v : c.scope) os << " " << m.vars[v].name;
        os << "\n";
    }
    return os.str();
}

void flatten_expr(const Expr &e, std::vector<stcsp_tok_t> &out) {
    for (const auto &k : e.kid) flatten_expr(*k, out);
    out.push_back(stcsp_tok_t{e.op, e.arg});
}

std::unique_ptr<FlatProblem> flatten(const Model &m) {
    auto f = std::make_unique<FlatProblem>();
    for (const auto &v : m.vars) {
        f->lb.push_back(v.lb);
        f->ub.push_back(v.ub);
        f->names.push_back(v.name);
    }
    for (const auto &n : f->names) f->name_ptrs.push_back(n.c_str());
    f->arr_offsets.push_back(0);
    for (const auto &a : m.arrays) {
        f->arr_values.insert(f->arr_values.end(), a.elements.begin(), a.elements.end());
        f->arr_offsets.push_back((int32_t)f->arr_values.size());
    }
    f->con_offsets.push_back(0);
    for (const auto &c : m.cons) {
        flatten_expr(*c.root, f->tokens);
        f->con_offsets.push_back((int32_t)f->tokens.size());
    }
    stcsp_problem_t &p = f->c;
    p.abi_version = STCSP_ABI_VERSION;
    p.prefix_k = m.prefix_k;
    p.n_vars = (int32_t)m.vars.size();
    p.var_lb = f->lb.data();
    p.var_ub = f->ub.data();
    p.var_names = f->name_ptrs.data();
    p.n_arrays = (int32_t)m.arrays.size();
    p.arr_offsets = f->arr_offsets.data();
    p.arr_values = f->arr_values.data();
    p.n_constraints = (int32_t)m.cons.size();
    p.con_offsets = f->con_offsets.data();
    p.con_tokens = f->tokens.data();
    return f;
}

ExprPtr unflatten(const stcsp_tok_t *tok, int32_t n) {
    std::vector<ExprPtr> st;
    for (int32_t i = 0; i < n; i++) {
        int ar = arity_of(tok[i].op);
        if (ar < 0 || tok[i].op == OP_FBY) throw std::runtime_error("unknown operator in token list");
        if ((int)st.size() < ar) throw std::runtime_error("malformed postfix token list (stack underflow)");
        auto e = mk(tok[i].op, tok[i].arg);
        e->kid.resize(ar);
        for (int k = ar - 1; k >= 0; k--) {
            e->kid[k] = std::move(st.back());
            st.pop_back();
        }
        st.push_back(std::move(e));
    }
    if (st.size() != 1) throw std::runtime_error("malformed postfix token list (not a single tree)");
    return std::move(st.back());
}

}  // namespace stcsp
