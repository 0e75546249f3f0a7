/* stcsp_host.h -- C ABI of the host side that surrounds the GPU path: the .csp front end
 * (language of reference src/stcsp.l + src/stcsp.y), the normaliser that produces the
 * constraint queue (reference src/solveralgorithm.cpp:16-332, src/constraint.cpp:254-318), and
 * the automaton post-processing / DOT output that follows the search (reference
 * src/graph.cpp:357-442, :167-355, :41-101; src/solveralgorithm.cpp:709-730).
 *
 * None of this touches the GPU; it exists so that tests and bindings can go
 *   .csp text -> stcsp_problem_t -> stcsp_gpu_solve -> stcsp_automaton_t -> solutions.dot
 * through C calls only.
 */
#ifndef STCSP_HOST_H
#define STCSP_HOST_H

#include "stcsp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stcsp_model stcsp_model_t;

/* Parse + normalise a model (replaces yyparse -> solverNew -> solverParse, reference
 * src/stcsp.y:57, src/solver.cpp:271-273).  prefix_k <= 0 means the default 2.
 * Errors: STCSP_ERR_PARSE with the reference's message format "Line %d: syntax error"
 * (src/stcsp.y:221-224) or "Variable '%s' has not been defined." (src/solver.cpp:33-36). */
int stcsp_model_parse_text(const char *text, int32_t prefix_k, stcsp_model_t **out);
int stcsp_model_parse_file(const char *path, int32_t prefix_k, stcsp_model_t **out);
void stcsp_model_free(stcsp_model_t *m);
/* Flat view of the model; valid until stcsp_model_free. */
const stcsp_problem_t *stcsp_model_problem(const stcsp_model_t *m);
/* Human-readable dump of variables and constraints (one per line, fully parenthesised); used by
 * tests to pin the normaliser.  Release with stcsp_string_free. */
char *stcsp_model_dump(const stcsp_model_t *m);

/* Post-processed automaton = what the reference prints: liveness w.r.t. `until`
 * (graphTraverse, src/graph.cpp:357-418), optional adversarial fixpoints (flags -a / -z,
 * src/graph.cpp:167-355), restricted to what is reachable from the root. */
typedef struct stcsp_solution {
    int32_t root_valid;              /* graph->root->valid after all passes */
    int32_t adver1;                  /* root->valid right after adversarialTraverse  (-1 if not run) */
    int32_t adver2;                  /* root->valid right after adversarialTraverse2 (-1 if not run) */
    int64_t n_states;                /* reachable, canonical (BFS) numbering, root = 0 */
    int64_t n_edges;
    int64_t n_table_states;          /* "# Number of nodes" line: all states seen incl. failed ones */
    int32_t n_vars, sig_len;
    int32_t *state_cset;             /* [n_states] canonical constraint-set numbering */
    uint8_t *state_final;            /* [n_states] doublecircle */
    int32_t *state_sig;              /* [n_states * sig_len] (root row unused) */
    int32_t *edge_src, *edge_dst;    /* [n_edges] sorted by (src, label) */
    int32_t *edge_label;             /* [n_edges * n_vars] */
    void *impl;
} stcsp_solution_t;

int stcsp_postprocess(const stcsp_problem_t *problem, const stcsp_automaton_t *automaton,
                      int32_t adversarial1, int32_t adversarial2, stcsp_solution_t *out);
void stcsp_solution_free(stcsp_solution_t *s);
/* solutions.dot text in the reference's format (src/solveralgorithm.cpp:709-730, src/graph.cpp:41-101),
 * vertices in canonical order.  Release with stcsp_string_free. */
char *stcsp_solution_dot(const stcsp_problem_t *problem, const stcsp_solution_t *s);
/* Canonical text; "EMPTY" when the root is not valid. */
char *stcsp_solution_canonical(const stcsp_problem_t *problem, const stcsp_solution_t *s);
void stcsp_string_free(char *s);
/* The same two texts streamed to a file (the reference writes ./solutions.dot, src/solveralgorithm.cpp:709-730; at
 * partialorder_20 size the text is ~10 GB and must not exist as one string), and the SHA-256 of the canonical text
 * computed line by line (65 bytes incl. the terminating NUL). */
int stcsp_solution_write_dot(const stcsp_problem_t *problem, const stcsp_solution_t *s, const char *path);
int stcsp_solution_write_canonical(const stcsp_problem_t *problem, const stcsp_solution_t *s, const char *path);
int stcsp_solution_canonical_sha256(const stcsp_problem_t *problem, const stcsp_solution_t *s, char out_hex[65]);

#ifdef __cplusplus
}
#endif
#endif /* STCSP_HOST_H */
