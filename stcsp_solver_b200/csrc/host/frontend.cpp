// .csp front end: scanner, recursive-descent parser and normaliser.
//
// The language is the one defined by the reference's lex/yacc sources (scanner rules
// src/stcsp.l:18-82, grammar src/stcsp.y:56-174).  lex and yacc do not exist
// in this build environment, so the scanner and parser are written by hand.  The normaliser
// restates reference constraintNormalise (src/solveralgorithm.cpp:60-332): it flattens the
// temporal operators into auxiliary variables `_V<n>` and primitive constraints
// `x == next y`, `first x == first y`, `x == y@n`, `x until y`.  The order in which auxiliaries
// and their defining constraints are created is parity-critical: it fixes the variable order
// and therefore the columns of every edge label of the automaton.
#include <climits>
#include <cstring>
#include <fstream>
#include <sstream>

#include "model.h"

namespace stcsp {
namespace {

// ------------------------------------------------------------------------------------------ scanner
enum Tok : int {
    T_EOF = 0, T_VAR = 300, T_OBJ, T_ARR, T_IDENT, T_CONST,
    T_LE_CON, T_GE_CON, T_EQ_CON, T_NE_CON, T_IMPLY, T_UNTIL,
    T_LT, T_GT, T_LE, T_GE, T_EQ, T_NE, T_AND, T_OR, T_NOT, T_ABS,
    T_FIRST, T_NEXT, T_FBY, T_IF, T_THEN, T_ELSE, T_AT
};

struct Lexeme {
    int tok = T_EOF;
    int32_t num = 0;
    std::string str;
};

class Scanner {
  public:
    explicit Scanner(const std::string &s) : src_(s) {}
    int line() const { return line_; }

    Lexeme next() {
        Lexeme lx;
        for (;;) {
            if (pos_ >= src_.size()) return lx;
            char c = src_[pos_];
            char d = pos_ + 1 < src_.size() ? src_[pos_ + 1] : '\0';
            if (c == '\n') { line_++; pos_++; continue; }
            if (c == ' ' || c == '\t' || c == '\v' || c == '\f') { pos_++; continue; }
            if (c == '/' && d == '/') {                     // needs its terminating newline (stcsp.l:20)
                size_t e = src_.find('\n', pos_);
                if (e != std::string::npos) { pos_ = e + 1; continue; }
            }
            if (c == '/' && d == '*') {                     // body may not contain " * / (stcsp.l:21)
                size_t q = pos_ + 2;
                while (q < src_.size() && src_[q] != '"' && src_[q] != '*' && src_[q] != '/') q++;
                if (q + 1 < src_.size() && src_[q] == '*' && src_[q + 1] == '/') { pos_ = q + 2; continue; }
            }
            if (c == '\'') {                                // quote comment to end of line (stcsp.l:79)
                while (pos_ < src_.size() && src_[pos_] != '\n') pos_++;
                continue;
            }
            if (is_letter(c)) {
                size_t e = pos_;
                while (e < src_.size() && (is_letter(src_[e]) || is_digit(src_[e]))) e++;
                lx.str = src_.substr(pos_, e - pos_);
                pos_ = e;
                lx.tok = keyword(lx.str);
                return lx;
            }
            if (is_digit(c) || (c == '-' && is_digit(d))) { // [-]?{D}+ through atoi (stcsp.l:77)
                size_t e = pos_ + 1;
                while (e < src_.size() && is_digit(src_[e])) e++;
                lx.tok = T_CONST;
                lx.num = (int32_t)atoi(src_.substr(pos_, e - pos_).c_str());
                pos_ = e;
                return lx;
            }
            if (c == '<' && d == '=') { pos_ += 2; lx.tok = T_LE_CON; return lx; }
            if (c == '>' && d == '=') { pos_ += 2; lx.tok = T_GE_CON; return lx; }
            if (c == '=' && d == '=') { pos_ += 2; lx.tok = T_EQ_CON; return lx; }
            if (c == '!' && d == '=') { pos_ += 2; lx.tok = T_NE_CON; return lx; }
            if (c == '-' && d == '>') { pos_ += 2; lx.tok = T_IMPLY; return lx; }
            if (c == '@') { pos_++; lx.tok = T_AT; return lx; }
            pos_++;
            lx.tok = (unsigned char)c;                      // anything else is returned as itself
            return lx;
        }
    }

  private:
    static bool is_letter(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
    static bool is_digit(char c) { return c >= '0' && c <= '9'; }
    static int keyword(const std::string &w) {
        static const struct { const char *kw; int tok; } table[] = {
            {"var", T_VAR}, {"obj", T_OBJ}, {"arr", T_ARR}, {"until", T_UNTIL},
            {"lt", T_LT}, {"gt", T_GT}, {"le", T_LE}, {"ge", T_GE}, {"eq", T_EQ}, {"ne", T_NE},
            {"and", T_AND}, {"or", T_OR}, {"not", T_NOT}, {"abs", T_ABS},
            {"first", T_FIRST}, {"next", T_NEXT}, {"fby", T_FBY},
            {"if", T_IF}, {"then", T_THEN}, {"else", T_ELSE}};
        for (const auto &k : table)
            if (w == k.kw) return k.tok;
        return T_IDENT;
    }
    const std::string &src_;
    size_t pos_ = 0;
    int line_ = 1;
};

// ------------------------------------------------------------------------------------------- parser
// Raw syntax tree: names are still strings (a name is resolved when its statement is processed,
// like the reference's solverParse does, src/solver.cpp:138-159).
struct Syn;
using SynPtr = std::unique_ptr<Syn>;
struct Syn {
    int32_t op = 0;             // stcsp_op / OP_FBY
    int32_t num = 0;
    std::string name;           // VAR / ARR
    std::vector<SynPtr> kid;
};

struct Statement {
    enum Kind { VarDecl, ArrDecl, Obj, Con } kind = Con;
    std::string name;
    int32_t lb = 0, ub = 0;
    std::vector<int32_t> elements;
    SynPtr con;
};

class Parser {
  public:
    explicit Parser(const std::string &text) : sc_(text) { advance(); }

    std::vector<Statement> program() {
        std::vector<Statement> out;
        while (la_.tok != T_EOF) out.push_back(statement());
        return out;
    }

  private:
    [[noreturn]] void fail() {                              // yyerror, src/stcsp.y:221-224
        throw ParseError("Line " + std::to_string(sc_.line()) + ": syntax error");
    }
    void advance() { la_ = sc_.next(); }
    void expect(int tok) { if (la_.tok != tok) fail(); advance(); }
    std::string ident() { if (la_.tok != T_IDENT) fail(); std::string s = la_.str; advance(); return s; }
    int32_t constant() { if (la_.tok != T_CONST) fail(); int32_t v = la_.num; advance(); return v; }

    static SynPtr node(int32_t op, SynPtr a = nullptr, SynPtr b = nullptr, SynPtr c = nullptr) {
        auto n = std::make_unique<Syn>();
        n->op = op;
        if (a) n->kid.push_back(std::move(a));
        if (b) n->kid.push_back(std::move(b));
        if (c) n->kid.push_back(std::move(c));
        return n;
    }

    Statement statement() {
        Statement st;
        if (la_.tok == T_VAR) {
            advance();
            st.kind = Statement::VarDecl;
            st.name = ident();
            expect(':'); expect('[');
            st.lb = constant();
            expect(',');
            st.ub = constant();
            expect(']'); expect(';');
            return st;
        }
        if (la_.tok == T_ARR) {
            advance();
            st.kind = Statement::ArrDecl;
            st.name = ident();
            expect(':'); expect('{');
            st.elements.push_back(constant());
            while (la_.tok == ',') { advance(); st.elements.push_back(constant()); }
            expect('}'); expect(';');
            return st;
        }
        if (la_.tok == T_OBJ) {
            advance();
            st.kind = Statement::Obj;
            st.name = ident();
            expect(';');
            return st;
        }
        SynPtr l = expression();
        int32_t op;
        switch (la_.tok) {
            case '<': op = STCSP_CON_LT; break;
            case '>': op = STCSP_CON_GT; break;
            case T_LE_CON: op = STCSP_CON_LE; break;
            case T_GE_CON: op = STCSP_CON_GE; break;
            case T_EQ_CON: op = STCSP_CON_EQ; break;
            case T_NE_CON: op = STCSP_CON_NE; break;
            case T_UNTIL: op = STCSP_CON_UNTIL; break;
            case T_IMPLY: op = STCSP_CON_IMPLY; break;
            default: fail();
        }
        advance();
        SynPtr r = expression();
        expect(';');
        st.kind = Statement::Con;
        st.con = node(op, std::move(l), std::move(r));
        return st;
    }

    SynPtr expression() {                                   // `not` binds loosest (stcsp.y:105-108)
        if (la_.tok == T_NOT) { advance(); return node(STCSP_OP_NOT, expression()); }
        return or_expr();
    }
    SynPtr or_expr() {
        SynPtr n = and_expr();
        while (la_.tok == T_OR) { advance(); n = node(STCSP_OP_OR, std::move(n), and_expr()); }
        return n;
    }
    SynPtr and_expr() {
        SynPtr n = eq_expr();
        while (la_.tok == T_AND) { advance(); n = node(STCSP_OP_AND, std::move(n), eq_expr()); }
        return n;
    }
    SynPtr eq_expr() {
        SynPtr n = rel_expr();
        while (la_.tok == T_EQ || la_.tok == T_NE) {
            int32_t op = la_.tok == T_EQ ? STCSP_OP_EQ : STCSP_OP_NE;
            advance();
            n = node(op, std::move(n), rel_expr());
        }
        return n;
    }
    SynPtr rel_expr() {
        SynPtr n = add_expr();
        for (;;) {
            int32_t op;
            if (la_.tok == T_LT) op = STCSP_OP_LT;
            else if (la_.tok == T_GT) op = STCSP_OP_GT;
            else if (la_.tok == T_LE) op = STCSP_OP_LE;
            else if (la_.tok == T_GE) op = STCSP_OP_GE;
            else return n;
            advance();
            n = node(op, std::move(n), add_expr());
        }
    }
    SynPtr add_expr() {
        SynPtr n = mul_expr();
        while (la_.tok == '+' || la_.tok == '-') {
            int32_t op = la_.tok == '+' ? STCSP_OP_ADD : STCSP_OP_SUB;
            advance();
            n = node(op, std::move(n), mul_expr());
        }
        return n;
    }
    SynPtr mul_expr() {
        SynPtr n = at_expr();
        while (la_.tok == '*' || la_.tok == '/' || la_.tok == '%') {
            int32_t op = la_.tok == '*' ? STCSP_OP_MUL : la_.tok == '/' ? STCSP_OP_DIV : STCSP_OP_MOD;
            advance();
            n = node(op, std::move(n), at_expr());
        }
        return n;
    }
    SynPtr at_expr() {
        SynPtr n = fby_expr();
        if (la_.tok == T_AT) {
            advance();
            int32_t t = constant();
            n = node(STCSP_OP_AT, std::move(n));
            n->num = t;
        }
        return n;
    }
    SynPtr fby_expr() {                                     // right associative
        SynPtr n = unary_expr();
        if (la_.tok == T_FBY) { advance(); n = node(OP_FBY, std::move(n), fby_expr()); }
        return n;
    }
    SynPtr unary_expr() {
        if (la_.tok == T_FIRST) { advance(); return node(STCSP_OP_FIRST, unary_expr()); }
        if (la_.tok == T_NEXT) { advance(); return node(STCSP_OP_NEXT, unary_expr()); }
        if (la_.tok == T_ABS) { advance(); return node(STCSP_OP_ABS, unary_expr()); }
        if (la_.tok == T_IF) {
            advance();
            SynPtr c = expression();
            expect(T_THEN);
            SynPtr t = expression();
            expect(T_ELSE);
            SynPtr e = unary_expr();
            return node(STCSP_OP_IF, std::move(c), std::move(t), std::move(e));
        }
        return primary_expr();
    }
    SynPtr primary_expr() {
        if (la_.tok == T_IDENT) {
            std::string name = ident();
            if (la_.tok == '[') {
                advance();
                SynPtr idx = expression();
                expect(']');
                SynPtr n = node(STCSP_OP_ARR, std::move(idx));
                n->name = name;
                return n;
            }
            SynPtr n = node(STCSP_OP_VAR);
            n->name = name;
            return n;
        }
        if (la_.tok == T_CONST) {
            SynPtr n = node(STCSP_OP_CONST);
            n->num = constant();
            return n;
        }
        if (la_.tok == '(') {
            advance();
            SynPtr n = expression();
            expect(')');
            return n;
        }
        fail();
    }

    Scanner sc_;
    Lexeme la_;
};

// --------------------------------------------------------------------------------------- normaliser
struct Bounds {
    int32_t lb = 0, ub = 0;
};

inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

class Normaliser {
  public:
    explicit Normaliser(Model &m) : m_(m) {}

    // reference constraintNodeParse, src/constraint.cpp:58-89: resolve names.
    ExprPtr resolve(const Syn &s) {
        ExprPtr e;
        if (s.op == STCSP_OP_VAR) {
            int32_t v = m_.find_var(s.name);
            if (v < 0) throw ParseError("Variable '" + s.name + "' has not been defined.");
            e = mk(STCSP_OP_VAR, v);
        } else if (s.op == STCSP_OP_ARR) {
            int32_t a = m_.find_array(s.name);
            if (a < 0) throw ParseError("Variable '" + s.name + "' has not been defined.");
            e = mk(STCSP_OP_ARR, a);
        } else if (s.op == STCSP_OP_CONST || s.op == STCSP_OP_AT) {
            e = mk(s.op, s.num);
        } else {
            e = mk(s.op);
        }
        for (const auto &k : s.kid) e->kid.push_back(resolve(*k));
        return e;
    }

    // reference solverAddConstr, src/solveralgorithm.cpp:16-24
    void add_statement(const Syn &s) {
        Bounds b;
        ExprPtr root = norm(resolve(s), b);
        if (!is_tautology(*root, m_.arrays)) push(std::move(root));
    }

  private:
    void push(ExprPtr root) {                               // solverConstraintQueuePush
        Constraint c;
        c.root = std::move(root);
        classify(c);
        m_.cons.push_back(std::move(c));
    }
    ExprPtr var(int32_t v) { return mk(STCSP_OP_VAR, v); }
    Bounds var_bounds(int32_t v) const { return Bounds{m_.vars[v].lb, m_.vars[v].ub}; }
    void add_eq_next(int32_t x, int32_t y) { push(mk2(STCSP_CON_EQ, var(x), mk1(STCSP_OP_NEXT, var(y)))); }
    void add_eq_node(int32_t x, ExprPtr e) { push(mk2(STCSP_CON_EQ, var(x), std::move(e))); }
    void add_first_eq_first(int32_t x, int32_t y) {
        push(mk2(STCSP_CON_EQ, mk1(STCSP_OP_FIRST, var(x)), mk1(STCSP_OP_FIRST, var(y))));
    }
    void add_eq_at(int32_t x, int32_t y, int32_t n) { push(mk2(STCSP_CON_EQ, var(x), mk1(STCSP_OP_AT, var(y), n))); }

    // Operand of fby / @ / until-free helper: an identifier is used as is, anything else is
    // normalised and named by a fresh auxiliary `y == <expr>`.
    int32_t name_operand(ExprPtr e, Bounds &b) {
        if (e->op == STCSP_OP_VAR) {
            b = var_bounds(e->arg);
            return e->arg;
        }
        ExprPtr n = norm(std::move(e), b);
        int32_t y = m_.add_aux(b.lb, b.ub);
        add_eq_node(y, std::move(n));
        return y;
    }

    ExprPtr norm(ExprPtr e, Bounds &b) {
        switch (e->op) {
            case STCSP_OP_FIRST: return norm_first(std::move(e), b);
            case STCSP_OP_NEXT: return norm_next(std::move(e), b);
            case OP_FBY: {                                  // src/solveralgorithm.cpp:161-188
                Bounds by, bz;
                int32_t y = name_operand(std::move(e->kid[0]), by);
                int32_t z = name_operand(std::move(e->kid[1]), bz);
                b.lb = std::min(by.lb, bz.lb);
                b.ub = std::max(by.ub, bz.ub);
                int32_t x = m_.add_aux(b.lb, b.ub);
                add_first_eq_first(x, y);
                add_eq_next(z, x);
                return var(x);
            }
            case STCSP_OP_AT: {                             // src/solveralgorithm.cpp:189-227
                if (e->kid[0]->op == STCSP_OP_CONST) {
                    b.lb = b.ub = e->kid[0]->arg;
                    return std::move(e->kid[0]);
                }
                if (e->kid[0]->op == STCSP_OP_NEXT)
                    throw ParseError("unsupported: '@' applied to a 'next' expression "
                                     "(the reference's handling of this form is undefined, "
                                     "src/solveralgorithm.cpp:195-209)");
                Bounds by;
                int32_t y = name_operand(std::move(e->kid[0]), by);
                b = by;
                int32_t x = m_.add_aux(b.lb, b.ub);
                add_eq_at(x, y, e->arg);
                return var(x);
            }
            case STCSP_OP_VAR: b = var_bounds(e->arg); return e;
            case STCSP_OP_CONST: b.lb = b.ub = e->arg; return e;
            case STCSP_OP_ARR: {                            // src/solveralgorithm.cpp:237-247
                Bounds bi;
                e->kid[0] = norm(std::move(e->kid[0]), bi);
                const auto &el = m_.arrays[e->arg].elements;
                b.lb = b.ub = el[0];
                for (int32_t v : el) { b.lb = std::min(b.lb, v); b.ub = std::max(b.ub, v); }
                return e;
            }
            case STCSP_CON_UNTIL: {                         // src/solveralgorithm.cpp:248-261
                bool l_ident = e->kid[0]->op == STCSP_OP_VAR, r_ident = e->kid[1]->op == STCSP_OP_VAR;
                Bounds bl, br;
                ExprPtr l = norm(std::move(e->kid[0]), bl);
                ExprPtr r = norm(std::move(e->kid[1]), br);
                if (!l_ident) {
                    int32_t x = m_.add_aux(0, 1);
                    add_eq_node(x, std::move(l));
                    l = var(x);
                }
                if (!r_ident) {
                    int32_t y = m_.add_aux(0, 1);
                    add_eq_node(y, std::move(r));
                    r = var(y);
                }
                e->kid[0] = std::move(l);
                e->kid[1] = std::move(r);
                b = Bounds{0, 1};
                return e;
            }
            default: break;
        }
        // Generic operators: operands left to right, then interval arithmetic for the bounds
        // (src/solveralgorithm.cpp:268-327).
        std::vector<Bounds> kb(e->kid.size());
        for (size_t i = 0; i < e->kid.size(); i++) e->kid[i] = norm(std::move(e->kid[i]), kb[i]);
        switch (e->op) {
            case STCSP_OP_ABS: {
                Bounds r = kb[0];
                if (r.lb < 0 && r.ub < 0) b = Bounds{wsub(0, r.ub), wsub(0, r.lb)};
                else if (r.lb < 0 && r.ub > 0) b = Bounds{0, std::max(wsub(0, r.lb), r.ub)};
                else b = r;                                 // incl. the reference's [lb<0, 0] case
                break;
            }
            case STCSP_OP_IF:                               // then/else union (THEN node in the reference)
                b = Bounds{std::min(kb[1].lb, kb[2].lb), std::max(kb[1].ub, kb[2].ub)};
                break;
            case STCSP_OP_ADD: b = Bounds{wadd(kb[0].lb, kb[1].lb), wadd(kb[0].ub, kb[1].ub)}; break;
            case STCSP_OP_SUB: b = Bounds{wsub(kb[0].lb, kb[1].ub), wsub(kb[0].ub, kb[1].lb)}; break;
            case STCSP_OP_MUL: {
                Bounds l = kb[0], r = kb[1];
                if (l.lb >= 0 && r.lb >= 0) b = Bounds{wmul(l.lb, r.lb), wmul(l.ub, r.ub)};
                else if (l.lb >= 0 && r.ub >= 0 && r.lb < 0) b = Bounds{wmul(l.ub, r.lb), wmul(l.ub, r.ub)};
                else if (l.ub >= 0 && l.lb < 0 && r.lb >= 0) b = Bounds{wmul(l.lb, r.lb), wmul(l.ub, r.ub)};
                else b = Bounds{wmul(l.ub, r.ub), wmul(l.lb, r.lb)};
                break;
            }
            case STCSP_OP_DIV: case STCSP_OP_MOD: b = Bounds{INT_MIN, INT_MAX}; break;
            default: b = Bounds{0, 1}; break;               // comparisons, and/or/not, constraint operators
        }
        return e;
    }

    ExprPtr norm_first(ExprPtr e, Bounds &b) {              // src/solveralgorithm.cpp:78-121
        Expr &r = *e->kid[0];
        if (r.op == STCSP_OP_CONST) {
            b.lb = b.ub = r.arg;
            return std::move(e->kid[0]);
        }
        if (r.op == STCSP_OP_VAR) {
            b = var_bounds(r.arg);
            return e;
        }
        if (r.op == OP_FBY) {                               // first (a fby b) == first a
            e->kid[0] = std::move(r.kid[0]);
            return norm(std::move(e), b);
        }
        ExprPtr n = norm(std::move(e->kid[0]), b);
        if (n->op == STCSP_OP_FIRST) n = std::move(n->kid[0]);   // first first e == first e
        e->kid[0] = std::move(n);
        return e;
    }

    ExprPtr norm_next(ExprPtr e, Bounds &b) {               // src/solveralgorithm.cpp:122-160
        Expr &r = *e->kid[0];
        if (r.op == STCSP_OP_CONST) {
            b.lb = b.ub = r.arg;
            return std::move(e->kid[0]);
        }
        if (r.op == STCSP_OP_VAR) {
            b = var_bounds(r.arg);
            int32_t x = m_.add_aux(b.lb, b.ub);
            add_eq_next(x, r.arg);
            return var(x);
        }
        if (r.op == OP_FBY) return norm(std::move(r.kid[1]), b);   // next (a fby b) == b
        ExprPtr n = norm(std::move(e->kid[0]), b);
        if (n->op == STCSP_OP_FIRST) return n;              // next first e == first e
        if (n->op == STCSP_OP_CONST || n->op == STCSP_OP_VAR) {
            e->kid[0] = std::move(n);
            return norm(std::move(e), b);
        }
        int32_t x = m_.add_aux(b.lb, b.ub);
        add_eq_node(x, std::move(n));
        int32_t y = m_.add_aux(b.lb, b.ub);
        add_eq_next(y, x);
        return var(y);
    }

    Model &m_;
};

}  // namespace

Model parse_model(const std::string &text, int32_t prefix_k) {
    Model m;
    m.prefix_k = prefix_k > 0 ? prefix_k : 2;
    Parser parser(text);
    std::vector<Statement> stmts = parser.program();        // the whole file must parse first
    Normaliser nz(m);
    for (const auto &st : stmts) {                          // reference solverParse, src/solver.cpp:138-159
        switch (st.kind) {
            case Statement::VarDecl: m.add_var(st.name, st.lb, st.ub); break;
            case Statement::ArrDecl: m.arrays.push_back(Array{st.name, st.elements}); break;
            case Statement::Obj:                            // reaches "Unknown token" in the reference
                throw ParseError("Unknown token: obj statements are not supported");
            case Statement::Con: nz.add_statement(*st.con); break;
        }
    }
    return m;
}

}  // namespace stcsp
