// Host half of the GPU path: constraint sets, their rewriting at time advance, and their
// compilation into the flat tables the kernels read.
//
//   reference                                        here
//   ConstraintNode trees (src/constraint.h:24-31) -> postfix bytecode with short-circuit jumps
//   Constraint / arcs (src/constraint.h:38-49)    -> DevCon + propagator list + wake masks
//   seenConstraints (src/solver.h:45)             -> SetTable (structural identity, src/constraint.cpp:564-576)
//   constraintTranslate (src/constraint.cpp:466-548) -> SetTable::successor
#pragma once

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../host/model.h"

namespace stcsp {

// ---- bytecode ------------------------------------------------------------------------------------
// One instruction = {op, arg}.  A value stack; `valid` is the sticky poison flag of the reference's
// evaluator (array index out of range, src/solveralgorithm.cpp:344-351; here also division by zero).
enum ByteOp : int32_t {
    BC_END = 0,
    BC_PUSHC,       // push arg
    BC_PUSHV,       // push value of scope slot arg
    BC_ARR,         // idx = pop; push arrays[arg][idx] or poison
    BC_ABS, BC_NOT,
    BC_LT, BC_GT, BC_LE, BC_GE, BC_EQ, BC_NE, BC_ADD, BC_SUB, BC_MUL, BC_DIV, BC_MOD,   // r = pop, l = pop; poison -> 0
    BC_JZ,          // c = pop; if c == 0 jump to arg
    BC_JMP,         // jump to arg
    BC_AND_SC,      // if top == 0 jump to arg (top stays 0) else pop
    BC_OR_SC,       // if top != 0 { top = 1; jump to arg } else pop
    BC_IMPLY_SC,    // if top == 0 { top = 1; jump to arg } (else keep left)
    BC_IMPLY_FIN,   // r = pop, l = pop; push l <= r   (no poison check, src/solveralgorithm.cpp:385-392)
};

struct Instr {
    int32_t op, arg;
};

enum DevKind : int32_t { DK_NEXT = 0, DK_POINT = 1, DK_UNTIL = 2 };

struct DevCon {            // one enforceable constraint of one constraint set
    int32_t kind;
    int32_t n_scope;       // POINT: variables in scope
    int32_t scope_off;     // into scope pool
    int32_t code_off;      // into code pool
    int32_t code_len;
    int32_t x, y;          // NEXT: x == next y.  UNTIL: x until y
    int32_t until_idx;     // UNTIL: index among the set's until constraints
    // POINT with a relation table (small product of declared widths): entry [sum bitpos(slot) * stride(slot)]
    // over the non-pivot slots = bitmask of the pivot variable's values completing a satisfying tuple
    int32_t pivot;         // scope slot of the pivot variable, -1: no table (bytecode enumeration)
    int32_t table_entries;
    long long table_off;   // into the table pool (u64 words)
};

struct DevProp {           // propagator = (constraint, time offset); NEXT/UNTIL use offset 0
    int32_t con;           // index into the con pool (absolute)
    int32_t offset;
};

struct DevSet {            // one constraint set
    int32_t prop_off, n_prop;
    int32_t wake_off;      // wake masks [(var * k + offset) * n_words + w]
    int32_t n_words;       // 32-bit words per propagator mask
    int32_t n_cap, cap_off;        // variables captured by `first` at time advance
    int32_t static_next;   // successor set when n_cap == 0
    int32_t n_until, until_off;    // right-hand variable of each until constraint (into aux pool)
    int32_t n_next, next_off;      // (x, y) pairs of the NEXT constraints (into aux pool)
    int32_t max_stack;
    int32_t n_cheap;       // propagators [0, n_cheap) are revised by one thread each (NEXT / UNTIL / tables over <= 4
                           // variables); the rest need a warp (wide tables, then bytecode enumerations)
    // contiguous ranges of this set in the con / scope / code pools (staged into shared memory by the kernels)
    int32_t con_off, n_con;
    int32_t scope_off, n_scope;
    int32_t code_off, n_code;
};

struct HostSet {
    std::vector<Constraint> cons;          // the set as the reference would hold it (identity + rewriting)
    bool has_first = false, has_at = false;
    std::vector<int32_t> cap_vars;
    int32_t static_next = -1;
};

struct Limits {
    static constexpr int kMaxScope = 32;        // variables per POINT constraint
    static constexpr int kMaxStack = 24;        // evaluator stack depth
    static constexpr int kMaxUntil = 30;        // until constraints (flags packed in one word)
    static constexpr int kMaxWidth = 64;        // values per domain (one 64-bit word)
    static constexpr int kMaxTableEntries = 1 << 22;    // per relation table (32 MiB: still L2-resident on B200)
    static constexpr long long kMaxTableWords = 1ll << 27;   // whole pool (1 GiB)
};

struct TableJob {          // a relation table the device still has to fill (build_tables_kernel)
    int32_t con;           // absolute constraint index
};

class SetTable {
  public:
    // Throws std::runtime_error (unsupported / malformed).
    void init(const stcsp_problem_t &p);

    int32_t n_vars() const { return (int32_t)lb_.size(); }
    int32_t k() const { return k_; }
    int32_t n_sets() const { return (int32_t)sets_.size(); }
    const std::vector<int32_t> &sig_vars() const { return sig_vars_; }
    int32_t n_until() const { return n_until_; }
    int32_t n_until_vars() const { return n_until_vars_; }
    int32_t max_scope() const { return max_scope_; }
    int32_t max_stack() const { return max_stack_; }
    int32_t max_props() const { return max_props_; }
    // shared-memory bytes needed to stage the metadata of the largest set (see stage_set in kernels.cu)
    size_t max_stage_bytes() const;
    const std::vector<int32_t> &lb() const { return lb_; }
    const std::vector<int32_t> &width() const { return width_; }
    const HostSet &host_set(int32_t s) const { return sets_[s]; }

    // Successor of set `s` when the time point's assignment is `values` (reference lines
    // src/solveralgorithm.cpp:755-805).  May create a new set (then dirty() becomes true).
    int32_t successor(int32_t s, const int32_t *values);

    // device pools (re-uploaded whenever dirty)
    bool dirty() const { return dirty_; }
    void clear_dirty() { dirty_ = false; }
    std::vector<DevSet> dev_sets;
    std::vector<DevCon> dev_cons;
    std::vector<DevProp> dev_props;
    std::vector<int32_t> dev_scope;
    std::vector<int32_t> dev_stride;        // parallel to dev_scope: table stride of each scope slot (0 for the pivot)
    std::vector<Instr> dev_code;
    std::vector<TableJob> table_jobs;       // tables allocated since the last take_table_jobs()
    long long table_words = 0;              // size of the device table pool in u64 words
    std::vector<uint32_t> dev_wake;
    std::vector<int32_t> dev_aux;
    std::vector<int32_t> arr_off, arr_val;

    // Relation tables already resident on the device (from an earlier solve of the same model): seeding them
    // before init() means no table is rebuilt.  export_tables() hands the directory back.
    struct TableRef { long long off; int32_t entries, pivot; std::vector<int32_t> strides; };
    typedef std::map<std::string, TableRef> TableDirectory;
    void seed_tables(const TableDirectory &dir, long long words) { table_cache_ = dir; table_words = words; }
    const TableDirectory &export_tables() const { return table_cache_; }

  private:
    int32_t add_set(std::vector<Constraint> cons);
    int32_t find_or_add(std::vector<Constraint> cons);
    void compile_set(int32_t s);
    void assign_table(DevCon &dc, const Constraint &c);
    void resolve_static(int32_t s);

    TableDirectory table_cache_;      // structurally equal constraints share one table
    int32_t k_ = 2;
    std::vector<int32_t> lb_, width_;
    std::vector<Array> arrays_;
    std::vector<HostSet> sets_;
    std::vector<int32_t> sig_vars_;
    std::vector<uint8_t> is_sig_;
    int32_t n_until_ = 0, n_until_vars_ = 0;
    int32_t max_scope_ = 1, max_stack_ = 1, max_props_ = 1;
    bool dirty_ = true;
};

// Compile one POINT constraint tree to bytecode; returns the maximum stack depth.
int compile_expr(const Expr &root, const std::vector<int32_t> &scope, std::vector<Instr> &out);

}  // namespace stcsp
