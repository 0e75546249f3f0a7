// Host-side model IR of the St-CSP: variables, arrays and the normalised constraint queue.
//
// Mirrors what the reference keeps in `Solver` after solverParse (reference src/solver.h:22-49):
// varQueue (src/variable.h:10-26), arrayQueue (src/variable.h:54-59) and constrQueue
// (src/constraint.h:38-49).  Expression trees are value-semantic `Expr` objects instead of
// malloc'ed ConstraintNode graphs; the flat postfix form of include/stcsp_b200.h is produced by
// `flatten()`.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "stcsp_b200.h"

namespace stcsp {

// Operators: the stcsp_op values of the C ABI plus front-end-only ones that never survive
// normalisation.
enum : int32_t {
    OP_FBY = 100,       // a fby b          (reference token FBY)
};

struct Expr;
using ExprPtr = std::unique_ptr<Expr>;

struct Expr {
    int32_t op = 0;
    int32_t arg = 0;            // CONST value / VAR index / ARR index / AT offset
    std::vector<ExprPtr> kid;   // 0..3 operands, evaluation order

    Expr() = default;
    Expr(int32_t o, int32_t a) : op(o), arg(a) {}
    ExprPtr clone() const;
    bool equals(const Expr &o) const;           // structural (token, num, var) equality
};

ExprPtr mk(int32_t op, int32_t arg = 0);
ExprPtr mk1(int32_t op, ExprPtr a, int32_t arg = 0);
ExprPtr mk2(int32_t op, ExprPtr a, ExprPtr b);
ExprPtr mk3(int32_t op, ExprPtr a, ExprPtr b, ExprPtr c);

int arity_of(int32_t op);                       // operands of a stcsp_op (0, 1, 2, 3), -1 if unknown
bool is_constraint_op(int32_t op);

struct Variable {
    std::string name;
    int32_t lb = 0, ub = 0;
};

struct Array {
    std::string name;
    std::vector<int32_t> elements;
};

enum class ConKind : int32_t { Next = 0, Point = 1, Until = 2, At = 3 };   // reference constraint.h:33-36

struct Constraint {
    ExprPtr root;
    // derived by classify():
    ConKind kind = ConKind::Point;
    bool has_first = false;
    std::vector<int32_t> scope;     // variables in first-occurrence order (reference constraintVarLinkRe)
};

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Model {
    int32_t prefix_k = 2;
    std::vector<Variable> vars;
    std::vector<Array> arrays;
    std::vector<Constraint> cons;
    int32_t n_aux = 0;

    int32_t find_var(const std::string &name) const;    // -1 if absent
    int32_t find_array(const std::string &name) const;
    int32_t add_var(const std::string &name, int32_t lb, int32_t ub);
    int32_t add_aux(int32_t lb, int32_t ub);            // `_V<n>` (reference solverAuxVarNew)
};

// Front end: text -> Model (parse + normalise + classify).  Throws ParseError.
Model parse_model(const std::string &text, int32_t prefix_k);

// Derive kind / has_first / scope of one constraint (reference solverConstraintQueuePush,
// src/constraint.cpp:254-318).
void classify(Constraint &c);
bool expr_has_first(const Expr &e);

// Lifted-int constant folding used by tautology checks and by the `first` rewrite
// (reference constraintNodeValue src/constraint.cpp:335-439, including its quirks).
struct Lifted {
    bool unknown = true;
    int32_t value = 0;
};
Lifted fold_value(const Expr &e, const std::vector<Array> &arrays);
bool is_tautology(const Expr &root, const std::vector<Array> &arrays);

// Text rendering (fully parenthesised) for dumps and tests.
std::string to_string(const Expr &e, const Model &m);
std::string dump_model(const Model &m);

// Flat C-ABI view.  Owns the storage the stcsp_problem_t points into.
struct FlatProblem {
    stcsp_problem_t c{};
    std::vector<int32_t> lb, ub, arr_offsets, arr_values, con_offsets;
    std::vector<stcsp_tok_t> tokens;
    std::vector<std::string> names;
    std::vector<const char *> name_ptrs;
};
void flatten_expr(const Expr &e, std::vector<stcsp_tok_t> &out);
std::unique_ptr<FlatProblem> flatten(const Model &m);

// Inverse of flatten for one constraint: postfix tokens -> tree.  Throws std::runtime_error on a
// malformed list.
ExprPtr unflatten(const stcsp_tok_t *tok, int32_t n);

}  // namespace stcsp
