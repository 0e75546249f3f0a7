/* TEST INFRASTRUCTURE (oracle/_ref build only) -- not part of the product.
 *
 * Stand-in for the reference's generated scanner/parser.  The reference builds its
 * front end with lex + yacc (reference src/stcsp.l, src/stcsp.y); neither tool exists
 * in this image, so this file is a hand-written scanner and recursive-descent parser
 * for the same language.  It builds the reference's own `Node` AST through the
 * reference's `nodeNew` (src/node.cpp:8-20) with the shapes the grammar actions
 * produce (src/stcsp.y:56-174) and then hands it to the reference's `solve(Node*)`
 * (src/solver.cpp:195) exactly where `yyparse` would (src/stcsp.y:57).  Everything
 * after that call is the unmodified reference, compiled from /root/reference/src.
 *
 * Compiled as gnu++98 together with the reference sources (oracle/build_ref.sh).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/resource.h>
#include "node.h"
#include "y.tab.h"

void solve(Node *node);                       /* reference src/solver.cpp:195 */

int line_num = 1;                             /* reference src/stcsp.y:28-30 */
int my_argc = 0;
char **my_argv = NULL;

/* calloc shim: the reference mallocs structs that contain std::vector members
 * (src/graph.cpp:137-141, src/variable.cpp:161-169); zero-filled storage makes those
 * members valid empty vectors.  Linked with -Wl,--wrap=malloc so only the calls made by
 * the reference objects are redirected; the algorithm is untouched. */
extern "C" void *__wrap_malloc(size_t n) { return calloc(1, n ? n : 1); }

namespace {

const int TOK_EOF = 0;

struct Lexeme {
    int tok;
    int num;
    std::string str;
};

struct Scanner {
    std::string src;
    size_t pos;
    Scanner() : pos(0) {}

    static bool isL(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
    static bool isD(char c) { return c >= '0' && c <= '9'; }

    /* One token, longest match first, like flex on stcsp.l:18-82. */
    Lexeme next() {
        Lexeme lx;
        lx.tok = TOK_EOF;
        lx.num = 0;
        for (;;) {
            if (pos >= src.size()) return lx;
            char c = src[pos];
            char c1 = pos + 1 < src.size() ? src[pos + 1] : '\0';
            if (c == '\n') { line_num++; pos++; continue; }
            if (c == ' ' || c == '\t' || c == '\v' || c == '\f') { pos++; continue; }
            if (c == '/' && c1 == '/') {                    /* "//"[^\n]*\n  (needs the newline) */
                size_t e = src.find('\n', pos);
                if (e != std::string::npos) { pos = e + 1; continue; }
                /* no trailing newline: rule does not match; '/' is returned as itself */
            }
            if (c == '/' && c1 == '*') {                    /* "/*"[^"*\/"]*"*\/" */
                size_t q = pos + 2;
                while (q < src.size() && src[q] != '"' && src[q] != '*' && src[q] != '/') q++;
                if (q + 1 < src.size() && src[q] == '*' && src[q + 1] == '/') { pos = q + 2; continue; }
            }
            if (c == '\'') {                                /* "'"[^\n]*  */
                while (pos < src.size() && src[pos] != '\n') pos++;
                continue;
            }
            if (isL(c)) {
                size_t e = pos;
                while (e < src.size() && (isL(src[e]) || isD(src[e]))) e++;
                std::string w = src.substr(pos, e - pos);
                pos = e;
                static const struct { const char *kw; int tok; } kws[] = {
                    {"var", VAR}, {"obj", OBJ}, {"arr", ARR}, {"until", UNTIL_CON},
                    {"lt", LT_OP}, {"gt", GT_OP}, {"le", LE_OP}, {"ge", GE_OP}, {"eq", EQ_OP}, {"ne", NE_OP},
                    {"and", AND_OP}, {"or", OR_OP}, {"not", NOT_OP}, {"abs", ABS},
                    {"first", FIRST}, {"next", NEXT}, {"fby", FBY}, {"if", IF}, {"then", THEN}, {"else", ELSE},
                    {NULL, 0}};
                for (int i = 0; kws[i].kw; i++)
                    if (w == kws[i].kw) { lx.tok = kws[i].tok; return lx; }
                lx.tok = IDENTIFIER;
                lx.str = w;
                return lx;
            }
            if (isD(c) || (c == '-' && isD(c1))) {          /* [-]?{D}+ via atoi */
                size_t e = pos + 1;
                while (e < src.size() && isD(src[e])) e++;
                lx.tok = CONSTANT;
                lx.num = atoi(src.substr(pos, e - pos).c_str());
                pos = e;
                return lx;
            }
            if (c == '<' && c1 == '=') { pos += 2; lx.tok = LE_CON; return lx; }
            if (c == '>' && c1 == '=') { pos += 2; lx.tok = GE_CON; return lx; }
            if (c == '=' && c1 == '=') { pos += 2; lx.tok = EQ_CON; return lx; }
            if (c == '!' && c1 == '=') { pos += 2; lx.tok = NE_CON; return lx; }
            if (c == '-' && c1 == '>') { pos += 2; lx.tok = IMPLY_CON; return lx; }
            if (c == '@') { pos++; lx.tok = AT; return lx; }
            pos++;
            lx.tok = (unsigned char)c;                      /* any other character is itself */
            return lx;
        }
    }
};

void syntaxError() {                                        /* stcsp.y:221-224 */
    fprintf(stdout, "Line %d: %s\n", line_num, "syntax error");
    exit(1);
}

struct Parser {
    Scanner sc;
    Lexeme la;

    void advance() { la = sc.next(); }
    void expect(int tok) { if (la.tok != tok) syntaxError(); advance(); }
    static char *dup(const std::string &s) { return strdup(s.c_str()); }
    static Node *basic(int token, Node *l, Node *r) { return nodeNew(token, NULL, 0, 0, l, r); }

    Node *program() {
        /* statement_list is right recursive: STATEMENT(left = stmt, right = rest). */
        std::vector<Node *> stmts;
        while (la.tok != TOK_EOF) stmts.push_back(statement());
        Node *list = NULL;
        for (size_t i = stmts.size(); i-- > 0;) list = basic(STATEMENT, stmts[i], list);
        return list;
    }

    Node *statement() {
        if (la.tok == VAR) {
            advance();
            if (la.tok != IDENTIFIER) syntaxError();
            char *name = dup(la.str); advance();
            expect(':'); expect('[');
            if (la.tok != CONSTANT) syntaxError();
            int lb = la.num; advance();
            expect(',');
            if (la.tok != CONSTANT) syntaxError();
            int ub = la.num; advance();
            expect(']'); expect(';');
            return nodeNew(VAR, name, 0, 0, NULL, nodeNew(RANGE, NULL, lb, ub, NULL, NULL));
        }
        if (la.tok == ARR) {
            advance();
            if (la.tok != IDENTIFIER) syntaxError();
            char *name = dup(la.str); advance();
            expect(':'); expect('{');
            if (la.tok != CONSTANT) syntaxError();
            Node *content = nodeNew(LIST, NULL, la.num, 0, NULL, NULL); advance();
            while (la.tok == ',') {
                advance();
                if (la.tok != CONSTANT) syntaxError();
                content = nodeNew(LIST, NULL, la.num, 0, content, NULL); advance();
            }
            expect('}'); expect(';');
            return nodeNew(ARR, name, 0, 0, NULL, content);
        }
        if (la.tok == OBJ) {
            advance();
            if (la.tok != IDENTIFIER) syntaxError();
            char *name = dup(la.str); advance();
            expect(';');
            return nodeNew(OBJ, name, 0, 0, NULL, NULL);
        }
        Node *l = expression();
        int op = la.tok;
        if (!(op == '<' || op == '>' || op == LE_CON || op == GE_CON || op == EQ_CON || op == NE_CON ||
              op == UNTIL_CON || op == IMPLY_CON)) syntaxError();
        advance();
        Node *r = expression();
        expect(';');
        return basic(op, l, r);
    }

    Node *expression() {                                    /* logical_not_expression */
        if (la.tok == NOT_OP) { advance(); return basic(NOT_OP, NULL, expression()); }
        return orExpr();
    }
    Node *orExpr() {
        Node *n = andExpr();
        while (la.tok == OR_OP) { advance(); n = basic(OR_OP, n, andExpr()); }
        return n;
    }
    Node *andExpr() {
        Node *n = eqExpr();
        while (la.tok == AND_OP) { advance(); n = basic(AND_OP, n, eqExpr()); }
        return n;
    }
    Node *eqExpr() {
        Node *n = relExpr();
        while (la.tok == EQ_OP || la.tok == NE_OP) { int op = la.tok; advance(); n = basic(op, n, relExpr()); }
        return n;
    }
    Node *relExpr() {
        Node *n = addExpr();
        while (la.tok == LT_OP || la.tok == GT_OP || la.tok == LE_OP || la.tok == GE_OP) {
            int op = la.tok; advance(); n = basic(op, n, addExpr());
        }
        return n;
    }
    Node *addExpr() {
        Node *n = mulExpr();
        while (la.tok == '+' || la.tok == '-') { int op = la.tok; advance(); n = basic(op, n, mulExpr()); }
        return n;
    }
    Node *mulExpr() {
        Node *n = atExpr();
        while (la.tok == '*' || la.tok == '/' || la.tok == '%') { int op = la.tok; advance(); n = basic(op, n, atExpr()); }
        return n;
    }
    Node *atExpr() {
        Node *n = fbyExpr();
        if (la.tok == AT) {
            advance();
            if (la.tok != CONSTANT) syntaxError();
            n = nodeNew(AT, NULL, la.num, 0, n, NULL);
            advance();
        }
        return n;
    }
    Node *fbyExpr() {                                       /* right associative */
        Node *n = unaryExpr();
        if (la.tok == FBY) { advance(); n = basic(FBY, n, fbyExpr()); }
        return n;
    }
    Node *unaryExpr() {
        if (la.tok == FIRST) { advance(); return basic(FIRST, NULL, unaryExpr()); }
        if (la.tok == NEXT) { advance(); return basic(NEXT, NULL, unaryExpr()); }
        if (la.tok == ABS) { advance(); return basic(ABS, NULL, unaryExpr()); }
        if (la.tok == IF) {
            advance();
            Node *c = expression();
            expect(THEN);
            Node *t = expression();
            expect(ELSE);
            Node *e = unaryExpr();
            return basic(IF, c, basic(THEN, t, e));
        }
        return primaryExpr();
    }
    Node *primaryExpr() {
        if (la.tok == IDENTIFIER) {
            char *name = dup(la.str); advance();
            if (la.tok == '[') {
                advance();
                Node *idx = expression();
                expect(']');
                return nodeNew(ARR_IDENTIFIER, name, 0, 0, NULL, idx);
            }
            return nodeNew(IDENTIFIER, name, 0, 0, NULL, NULL);
        }
        if (la.tok == CONSTANT) { int v = la.num; advance(); return nodeNew(CONSTANT, NULL, v, 0, NULL, NULL); }
        if (la.tok == '(') { advance(); Node *n = expression(); expect(')'); return n; }
        syntaxError();
        return NULL;
    }
};

}  // namespace

int main(int argc, char *argv[]) {                          /* stcsp.y:180-219 */
    struct rlimit x;
    if (getrlimit(RLIMIT_STACK, &x) == 0) { x.rlim_cur = x.rlim_max; setrlimit(RLIMIT_STACK, &x); }

    my_argc = argc;
    my_argv = argv;
    const char *filename = NULL;
    for (int i = 1; filename == NULL && i < argc; i++)
        if (argv[i][0] != '-') filename = argv[i];

    FILE *in = filename ? fopen(filename, "r") : stdin;
    if (in == NULL) { perror(filename); return 1; }
    Parser p;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, in)) > 0) p.sc.src.append(buf, n);
    if (filename) fclose(in);

    p.advance();
    Node *ast = p.program();
    solve(ast);
    return 0;
}
