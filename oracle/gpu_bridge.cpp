/* TEST INFRASTRUCTURE (oracle/_ref build only) -- the reference-side binding of the C ABI, COMPILED.
 *
 * This is the file INTEGRATION.md tells a reference maintainer to add (src/gpu_bridge.cpp): a
 * replacement for `double solverSolve(Solver *, bool)` (reference src/solveralgorithm.h:11, body
 * src/solveralgorithm.cpp:945-1005, callers src/solver.cpp:289 and :323) that
 *
 *   1. flattens what solverSolve reads from `Solver` (src/solver.h:22-49: varQueue, arrayQueue,
 *      constrQueue with its ConstraintNode trees src/constraint.h:24-31, prefixK) into
 *      stcsp_problem_t (include/stcsp_b200.h),
 *   2. calls stcsp_gpu_solve (the B200 path; no CPU fallback),
 *   3. rebuilds the reference's own `Graph` from stcsp_automaton_t with the reference's own
 *      constructors -- vertexNew src/graph.cpp:14-29, vertexTableAddVertex :103-105, edgeNew :78-89,
 *      vertexAddEdge :33-38 --
 *   4. and then runs the UNMODIFIED tail of solverSolve: graphTraverse (src/graph.cpp:357-418),
 *      adversarialTraverse / adversarialTraverse2 (:304-355 / :247-302), renumberVertex (:420-442),
 *      solverOut (src/solveralgorithm.cpp:709-730) and the stat line (:1000-1001).
 *
 * oracle/build_ref.sh links it with the reference's eight .cpp files (compiled where they lie) and
 * `-Wl,--wrap=_Z11solverSolveP6Solverb`, so that the calls in src/solver.cpp land here without
 * touching a reference source file: the result, oracle/_ref/stcsp_ref_gpu, is the reference's front
 * end, normaliser, post-processing and DOT writer around the GPU search.  tests/test_gpu_bridge.py
 * canonicalises its solutions.dot against the goldens -- which also cross-checks the product's own
 * post-processing (csrc/host/postprocess.cpp) against the reference's src/graph.cpp:357-442.
 *
 * C++98, like the reference (src/graph.h needs it, see build_ref.sh).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "util.h"
#include "solver.h"
#include "solveralgorithm.h"
#include "constraint.h"
#include "variable.h"
#include "graph.h"
#include "y.tab.h"

#include "stcsp_b200.h"

void solverOut(Solver *solver);                 /* src/solveralgorithm.cpp:709 (not in its header) */

namespace {

int opOf(int token) {                           /* reference token -> enum stcsp_op */
    switch (token) {
        case CONSTANT: return STCSP_OP_CONST;
        case IDENTIFIER: return STCSP_OP_VAR;
        case ARR_IDENTIFIER: return STCSP_OP_ARR;
        case ABS: return STCSP_OP_ABS;
        case NOT_OP: return STCSP_OP_NOT;
        case FIRST: return STCSP_OP_FIRST;
        case NEXT: return STCSP_OP_NEXT;
        case AT: return STCSP_OP_AT;
        case IF: return STCSP_OP_IF;
        case LT_OP: return STCSP_OP_LT;
        case GT_OP: return STCSP_OP_GT;
        case LE_OP: return STCSP_OP_LE;
        case GE_OP: return STCSP_OP_GE;
        case EQ_OP: return STCSP_OP_EQ;
        case NE_OP: return STCSP_OP_NE;
        case AND_OP: return STCSP_OP_AND;
        case OR_OP: return STCSP_OP_OR;
        case '+': return STCSP_OP_ADD;
        case '-': return STCSP_OP_SUB;
        case '*': return STCSP_OP_MUL;
        case '/': return STCSP_OP_DIV;
        case '%': return STCSP_OP_MOD;
        case '<': return STCSP_CON_LT;
        case '>': return STCSP_CON_GT;
        case LE_CON: return STCSP_CON_LE;
        case GE_CON: return STCSP_CON_GE;
        case EQ_CON: return STCSP_CON_EQ;
        case NE_CON: return STCSP_CON_NE;
        case IMPLY_CON: return STCSP_CON_IMPLY;
        case UNTIL_CON: return STCSP_CON_UNTIL;
    }
    myLog(LOG_ERROR, "gpu_bridge: token %d has no C-ABI operator\n", token);
    exit(1);
}

int indexOfVar(Solver *s, Variable *v) {
    for (size_t i = 0; i < s->varQueue->size(); i++)
        if ((*s->varQueue)[i] == v) return (int)i;
    myLog(LOG_ERROR, "gpu_bridge: variable not in varQueue\n");
    exit(1);
}

int indexOfArray(Solver *s, Array *a) {
    for (size_t i = 0; i < s->arrayQueue->size(); i++)
        if ((*s->arrayQueue)[i] == a) return (int)i;
    myLog(LOG_ERROR, "gpu_bridge: array not in arrayQueue\n");
    exit(1);
}

/* ConstraintNode tree -> postfix tokens, children first.  Child layout per token as constraintNodeParse builds it
 * (src/constraint.cpp:56-90) and solverValidateRe reads it (src/solveralgorithm.cpp:336-424): unary operators keep their
 * operand in ->right, `T[i]` its index in ->right, `y@n` = (left: y, right: CONSTANT n), IF = (left: cond, right: THEN
 * node (left: then, right: else)). */
void flatten(Solver *s, ConstraintNode *n, std::vector<stcsp_tok_t> &out) {
    if (n == NULL) return;
    stcsp_tok_t t;
    t.op = opOf(n->token);
    t.arg = 0;
    switch (n->token) {
        case CONSTANT: t.arg = n->num; break;
        case IDENTIFIER: t.arg = indexOfVar(s, n->var); break;
        case ARR_IDENTIFIER: flatten(s, n->right, out); t.arg = indexOfArray(s, n->array); break;
        case ABS: case NOT_OP: case FIRST: case NEXT: flatten(s, n->right, out); break;
        case AT: flatten(s, n->left, out); t.arg = n->right->num; break;
        case IF:
            flatten(s, n->left, out);
            flatten(s, n->right->left, out);
            flatten(s, n->right->right, out);
            break;
        default: flatten(s, n->left, out); flatten(s, n->right, out); break;
    }
    out.push_back(t);
}

}  // namespace

/* The linker redirects every call of solverSolve(Solver*, bool) here (--wrap on the mangled name). */
extern "C" double __wrap__Z11solverSolveP6Solverb(Solver *solver, bool testing) {
    solver->solveTime = cpuTime();
    const int numVar = (int)solver->varQueue->size();

    /* ---- 1. Solver -> stcsp_problem_t */
    std::vector<int32_t> lb, ub, conOff(1, 0), arrOff(1, 0), arrVal;
    std::vector<const char *> names;
    std::vector<stcsp_tok_t> toks;
    for (int v = 0; v < numVar; v++) {
        Variable *var = (*solver->varQueue)[v];
        lb.push_back(var->lb);
        ub.push_back(var->ub);
        names.push_back(var->name);
    }
    for (size_t a = 0; a < solver->arrayQueue->size(); a++) {
        Array *arr = (*solver->arrayQueue)[a];
        for (int i = 0; i < arr->size; i++) arrVal.push_back(arr->elements[i]);
        arrOff.push_back((int32_t)arrVal.size());
    }
    for (size_t c = 0; c < solver->constrQueue->size(); c++) {
        flatten(solver, (*solver->constrQueue)[c]->node, toks);
        conOff.push_back((int32_t)toks.size());
    }
    int32_t none = 0;
    stcsp_tok_t no_tok;
    no_tok.op = no_tok.arg = 0;
    stcsp_problem_t p;
    memset(&p, 0, sizeof p);
    p.abi_version = STCSP_ABI_VERSION;
    p.prefix_k = solver->prefixK;
    p.n_vars = numVar;
    p.var_lb = lb.empty() ? &none : &lb[0];
    p.var_ub = ub.empty() ? &none : &ub[0];
    p.var_names = names.empty() ? NULL : &names[0];
    p.n_arrays = (int32_t)arrOff.size() - 1;
    p.arr_offsets = &arrOff[0];
    p.arr_values = arrVal.empty() ? &none : &arrVal[0];
    p.n_constraints = (int32_t)conOff.size() - 1;
    p.con_offsets = &conOff[0];
    p.con_tokens = toks.empty() ? &no_tok : &toks[0];

    /* ---- 2. the search, on the GPU */
    stcsp_options_t o;
    memset(&o, 0, sizeof o);
    o.device = -1;
    stcsp_automaton_t a;
    if (stcsp_gpu_solve(&p, &o, &a) != STCSP_OK) {
        fprintf(stderr, "gpu_bridge: %s\n", stcsp_last_error());
        exit(1);
    }

    /* ---- 3. stcsp_automaton_t -> the reference's Graph (what solverSolveRe leaves behind, src/solveralgorithm.cpp:842-906) */
    solver->seenConstraints->push_back(solver->constrQueue);
    std::vector<Vertex *> vertex((size_t)a.n_states);
    for (int64_t s = 0; s < a.n_states; s++) {
        vector<int> sig;
        if (s != 0)
            for (int j = 0; j < a.sig_len; j++) sig.push_back(a.state_sig[s * a.sig_len + j]);
        Signature *signature = new Signature(sig, s == 0 ? 0 : a.state_cset[s]);
        vertex[s] = vertexNew(solver->graph, signature, 0);
        vertex[s]->fail = a.state_failed[s] != 0;
        vertexTableAddVertex(solver->graph->vertexTable, vertex[s]);
    }
    solver->graph->root = vertex[0];
    solver->graph->root->final = true;                       /* src/solveralgorithm.cpp:956-964 */
    for (size_t c = 0; c < solver->constrQueue->size(); c++)
        if ((*solver->constrQueue)[c]->type == CONSTR_UNTIL) { solver->graph->root->final = false; break; }
    for (int64_t e = 0; e < a.n_edges; e++) {
        /* edgeNew copies variableGetValue() of every variable (src/graph.cpp:84-87): bind the window to the label first */
        for (int v = 0; v < numVar; v++) {
            Variable *var = (*solver->varQueue)[v];
            var->currLB[0] = var->currUB[0] = a.edge_label[e * numVar + v];
        }
        Edge *edge = edgeNew(vertex[a.edge_src[e]], vertex[a.edge_dst[e]], solver->varQueue, numVar);
        vertexAddEdge(vertex[a.edge_src[e]], edge);
    }
    solver->numDominance = (int)a.n_dominance;
    solver->numFails = (int)a.n_fails;
    stcsp_automaton_free(&a);
    solver->solveTime = cpuTime() - solver->solveTime;

    /* ---- 4. the tail of solverSolve, unchanged (src/solveralgorithm.cpp:972-1004) */
    solver->processTime = cpuTime();
    graphTraverse(solver->graph, solver->numSignVar, solver->numUntil);
    if (solver->adversarial1) {
        adversarialTraverse(solver->graph, solver->varQueue);
        printf("adver1: %d; ", solver->graph->root->valid);
    }
    if (solver->adversarial2) {
        adversarialTraverse2(solver->graph, solver->varQueue);
        printf("adver2: %d\n", solver->graph->root->valid);
    }
    renumberVertex(solver->graph);
    solver->processTime = cpuTime() - solver->processTime;
    solver->numNodes = (int)solver->graph->vertexTable->size();
    if (solver->printSolution) solverOut(solver);
    if (!testing) {
        printf("%.2f\t%d\t%d\t%d\t%d\t%d\t%.2f\t%.5f\n", solver->initTime, (int)solver->varQueue->size(),
               (int)solver->constrQueue->size(), solver->numDominance, solver->numNodes, solver->numFails, solver->solveTime,
               solver->processTime);
        fflush(stdout);
    }
    return solver->solveTime + solver->processTime;
}
