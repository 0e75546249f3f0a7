/* TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * Lets oracle/gpu_bridge.cpp -- the compiled reference-side binding of the C ABI -- be exercised on a machine WITHOUT a
 * GPU: the three C-ABI symbols the bridge calls are served by the CPU oracle (oracle/stcsp_oracle.cpp, same contract and
 * same flat structs as stcsp_gpu_solve).  The result, oracle/_ref/stcsp_ref_bridge_cpu, checks the bridge itself (tree
 * flattening, array flattening, Graph rebuild, the reference's own post-processing on the rebuilt Graph) in the CPU test
 * suite; the product is checked through oracle/_ref/stcsp_ref_gpu, which links the real library.  Never shipped.
 */
#include "stcsp_b200.h"

extern "C" {
int stcsp_oracle_solve(const stcsp_problem_t *problem, double time_limit_s, stcsp_automaton_t *out, void *stats);
void stcsp_oracle_automaton_free(stcsp_automaton_t *a);
const char *stcsp_oracle_last_error(void);

int stcsp_gpu_solve(const stcsp_problem_t *problem, const stcsp_options_t *options, stcsp_automaton_t *out) {
    (void)options;
    return stcsp_oracle_solve(problem, 0.0, out, 0);
}
void stcsp_automaton_free(stcsp_automaton_t *a) { stcsp_oracle_automaton_free(a); }
const char *stcsp_last_error(void) { return stcsp_oracle_last_error(); }
}
