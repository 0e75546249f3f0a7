// stcsp -- command-line front of the B200 solver.
//
// Keeps the reference's command line (reference src/stcsp.y:180-219 `main`, src/solver.cpp:195-359
// `solve`): `stcsp [-s] [-m<sec>] [-t] [-a] [-z] [-k<K>] [-l<level>] file.csp`; the input file is the
// first argument that does not start with '-' (so `-m 3600` with a space is the reference's
// pitfall too).  Output is the reference's: one tab-separated statistics line
// `initTime vars constraints numDominance numNodes numFails solveTime processTime`
// (src/solveralgorithm.cpp:1000-1001), `adver1: %d; ` / `adver2: %d\n` prefixes for -a / -z
// (:975-983), and with -s the automaton in `solutions.dot` of the working directory (:709-730).
// The search itself runs on the GPU through the C ABI (include/stcsp_b200.h); there is no CPU solver.
// Extensions: --canonical (print the canonical automaton text instead of writing DOT),
// --stats (print the GPU path's own counters on stderr), --sha256 (SHA-256 of the canonical text on stderr),
// --gpus N (shard the search over N GPUs of this box, one host thread per GPU; instances whose waves fit one GPU stay on
// one unless --shard is given).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "stcsp_host.h"

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Cli {
    bool print_solution = false, testing = false, adv1 = false, adv2 = false, canonical = false, stats = false, sha256 = false;
    int prefix_k = 2, time_limit = 0, log_level = 0, gpus = 1, shard = 0;
    const char *file = nullptr;
};

struct Run {
    double init_s = 0, solve_s = 0, process_s = 0;
};

// One parse + solve + post-process, like one solverNew/solverParse/solverSolve/solverFree round.
int run_once(const Cli &cli, bool print_stat, Run &r) {
    double t = now_s();
    stcsp_model_t *model = nullptr;
    int rc = stcsp_model_parse_file(cli.file, cli.prefix_k, &model);
    if (rc != STCSP_OK) {
        // syntax errors go to stdout like yyerror (src/stcsp.y:221-224); the rest is the reference's error.txt text
        if (rc == STCSP_ERR_PARSE && !strncmp(stcsp_last_error(), "Line ", 5)) printf("%s\n", stcsp_last_error());
        else fprintf(stderr, "stcsp: %s\n", stcsp_last_error());
        return 1;
    }
    const stcsp_problem_t *problem = stcsp_model_problem(model);
    r.init_s = now_s() - t;

    stcsp_options_t opt;
    memset(&opt, 0, sizeof opt);
    opt.device = -1;
    opt.time_limit_s = cli.time_limit;
    opt.verbosity = cli.log_level;
    stcsp_automaton_t automaton;
    t = now_s();
    opt.shard_mode = cli.shard;
    opt.adversarial = (cli.adv1 ? 1 : 0) | (cli.adv2 ? 2 : 0);      // -a / -z fixpoints run on the device, before the download
    stcsp_exchange_stats_t xs;
    memset(&xs, 0, sizeof xs);
    rc = cli.gpus > 1 ? stcsp_gpu_solve_multi(problem, &opt, cli.gpus, nullptr, &automaton, &xs)     // one host thread per GPU
                      : stcsp_gpu_solve(problem, &opt, &automaton);
    r.solve_s = now_s() - t;
    if (rc == STCSP_ERR_TIMEOUT) {          // reference: SIGALRM handler exits 0 silently (src/solver.cpp:190-193)
        fprintf(stderr, "stcsp: time limit reached\n");
        stcsp_model_free(model);
        exit(0);
    }
    if (rc != STCSP_OK) {
        fprintf(stderr, "stcsp: %s\n", stcsp_last_error());
        stcsp_model_free(model);
        return 1;
    }
    t = now_s();
    stcsp_solution_t sol;
    rc = stcsp_postprocess(problem, &automaton, cli.adv1, cli.adv2, &sol);
    r.process_s = now_s() - t;
    if (rc != STCSP_OK) {
        fprintf(stderr, "stcsp: %s\n", stcsp_last_error());
        stcsp_automaton_free(&automaton);
        stcsp_model_free(model);
        return 1;
    }
    if (cli.adv1) printf("adver1: %d; ", sol.adver1);
    if (cli.adv2) printf("adver2: %d\n", sol.adver2);
    if (cli.print_solution && stcsp_solution_write_dot(problem, &sol, "solutions.dot") != STCSP_OK)      // streamed line by line
        fprintf(stderr, "stcsp: %s\n", stcsp_last_error());
    if (cli.sha256) {
        char hex[65];
        if (stcsp_solution_canonical_sha256(problem, &sol, hex) == STCSP_OK)
            fprintf(stderr, "canonical sha256 %s states %lld edges %lld\n", hex, (long long)sol.n_states, (long long)sol.n_edges);
    }
    if (cli.canonical) {
        char *txt = stcsp_solution_canonical(problem, &sol);
        fputs(txt, stdout);
        if (!strcmp(txt, "EMPTY")) fputc('\n', stdout);
        stcsp_string_free(txt);
    }
    if (print_stat) {
        printf("%.2f\t%d\t%d\t%d\t%d\t%d\t%.2f\t%.5f\n", r.init_s, problem->n_vars, problem->n_constraints,
               (int)automaton.n_dominance, (int)automaton.n_states, (int)automaton.n_fails, r.solve_s, r.process_s);
        fflush(stdout);
    }
    if (cli.stats)
        fprintf(stderr,
                "gpu: states %lld edges %lld sets %d search_nodes %lld fails %lld leaves %lld waves %lld tuples %lld "
                "launches %lld device_ms %.3f wall_ms %.3f\n",
                (long long)automaton.n_states, (long long)automaton.n_edges, automaton.n_constraint_sets,
                (long long)automaton.n_search_nodes, (long long)automaton.n_fails, (long long)automaton.n_leaves,
                (long long)automaton.n_waves, (long long)automaton.n_tuples, (long long)automaton.n_kernel_launches,
                automaton.solve_ms, automaton.wall_ms);
    if (cli.stats && cli.gpus > 1)
        fprintf(stderr, "gpus %d: sharded %d waves %lld exchanges %lld records %lld bytes_over_nvlink %lld exchange_ms %.3f\n", cli.gpus,
                xs.sharded, (long long)xs.waves, (long long)xs.exchanges, (long long)xs.records, (long long)xs.bytes_pulled,
                xs.exchange_ms);
    stcsp_solution_free(&sol);
    stcsp_automaton_free(&automaton);
    stcsp_model_free(model);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    Cli cli;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (a[0] != '-') {
            if (!cli.file) cli.file = a;
            continue;
        }
        if (!strcmp(a, "--canonical")) { cli.canonical = true; continue; }
        if (!strcmp(a, "--stats")) { cli.stats = true; continue; }
        if (!strcmp(a, "--sha256")) { cli.sha256 = true; continue; }
        if (!strcmp(a, "--shard")) { cli.shard = 1; continue; }
        if (!strncmp(a, "--gpus", 6)) {          // --gpus N or --gpus=N (the one new flag)
            const char *val = a[6] == '=' ? a + 7 : (i + 1 < argc ? argv[++i] : "");
            if (sscanf(val, "%d", &cli.gpus) != 1 || cli.gpus < 1 || cli.gpus > 16) {
                fprintf(stderr, "Invalid argument: %s\n", val);
                return 1;
            }
            continue;
        }
        for (const char *p = a + 1; *p; p++) {
            const char c = *p;
            if (c == 's') cli.print_solution = true;
            else if (c == 't') cli.testing = true;
            else if (c == 'a') cli.adv1 = true;
            else if (c == 'z') cli.adv2 = true;
            else if (c == 'k' || c == 'm' || c == 'l' || c == 'b' || c == 'e' || c == 'v') {
                const char *val = p[1] ? p + 1 : (i + 1 < argc ? argv[++i] : "");
                int n = 0;
                if (c == 'k' || c == 'm' || c == 'l') {
                    if (sscanf(val, "%d", &n) != 1) {
                        fprintf(stderr, "Invalid argument: %s\n", val);
                        return 1;
                    }
                    if (c == 'k') cli.prefix_k = n;
                    else if (c == 'm') cli.time_limit = n;
                    else cli.log_level = n;
                }                                   // -b -e -v are accepted and ignored, like the reference
                break;
            } else {
                fprintf(stderr, "Unknown argument: %c\n", c);
                return 1;
            }
        }
    }
    if (!cli.file) {
        fprintf(stderr, "usage: stcsp [-s] [-m<sec>] [-t] [-a] [-z] [-k<K>] [-l<level>] [--canonical] [--stats] [--sha256] [--gpus N] [--shard] file.csp\n");
        return 1;
    }
    Run r;
    int rc = run_once(cli, !cli.testing, r);
    if (rc) return rc;
    if (cli.testing) {          // timing loop of the reference (src/solver.cpp:295-349): until the 95% CI is < 5% of the mean
        std::vector<double> times;
        for (int n = 0;; n++) {
            printf("%d ", n);
            fflush(stdout);
            if ((rc = run_once(cli, true, r))) return rc;
            times.push_back(r.solve_s + r.process_s);
            const size_t cnt = times.size();
            if (cnt < 10) continue;
            double mean = 0, var = 0;
            for (double x : times) mean += x;
            mean /= (double)cnt;
            for (double x : times) var += (x - mean) * (x - mean);
            var /= (double)cnt - 1;
            if (2 * 1.96 * std::sqrt(var) / std::sqrt((double)cnt) < 0.05 * mean) {
                printf("\nMean execution time is %f pm %f\n", mean, 1.96 * std::sqrt(var) / std::sqrt((double)cnt));
                break;
            }
        }
    }
    return 0;
}
