#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- regenerate tests/golden/*.json by running the reference itself.

Runs ``oracle/_ref/stcsp_ref`` (the unmodified reference solver, see oracle/build_ref.sh) on
the generated benchmark instances (stcsp_solver_b200/instances.py) and on the hand-made
feature probes below, canonicalises the ``solutions.dot`` it writes
(stcsp_solver_b200/canonical.py) and stores, per case: the model text
(probes only), flags, the reference's stat line, reachable states/edges, the canonical
SHA-256 and -- for small automata -- the canonical text itself.

Usage:  python oracle/make_goldens.py [--jobs N] [--only REGEX] [--max-seconds S]
(REGEX is searched in the case key, e.g. "digitinvader[6-9]_[az]".)
The big instances take long on one core (digitinvader9 ~35 min, digitinvader8 ~13 min).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import re
import resource
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stcsp_solver_b200 import canonical, instances  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "stcsp_ref")
OUT = os.path.join(ROOT, "tests", "golden")

# Hand-made feature probes: every language feature and leaf rule the shipped examples do not
# exercise (SURVEY.md Appendix G "feature probes").  name -> (model text, [flag sets]).
PROBES = {
    "probe_stateless": ("var X : [0, 3];\nX < 2;\n", [""]),
    "probe_unsat_next": ("var X : [0, 3];\nnext X > X;\n", [""]),
    "probe_unsat_root": ("var X : [0, 3];\nvar Y : [0, 3];\nX > Y;\nY > X;\n", [""]),
    "probe_counter": ("var X : [0, 5];\nfirst X == 0;\nnext X == if X eq 5 then 0 else (X + 1);\n", [""]),
    "probe_until": ("var X : [0, 1];\nvar Y : [0, 1];\nvar C : [0, 3];\nfirst C == 0;\n"
                    "next C == if C eq 3 then 3 else (C + 1);\nY == (C eq 2);\nX until Y;\n", [""]),
    "probe_until_expr": ("var C : [0, 4];\nfirst C == 0;\nnext C == if C eq 4 then 0 else (C + 1);\n"
                         "(C lt 3) until (C eq 3);\n", [""]),
    "probe_until_never": ("var X : [0, 1];\nvar Y : [0, 1];\nY == 0;\nX until Y;\n", [""]),
    "probe_until_two": ("var X : [0, 1];\nvar Y : [0, 1];\nvar Z : [0, 1];\nvar C : [0, 3];\nfirst C == 0;\n"
                        "next C == if C eq 3 then 3 else (C + 1);\nY == (C ge 1);\nZ == (C ge 2);\n"
                        "X until Y;\nX until Z;\n", [""]),
    "probe_at2": ("var X : [0, 2];\nvar Y : [0, 2];\nY == X@2;\n", [""]),
    "probe_at1_next": ("var X : [0, 3];\nvar Y : [0, 3];\nfirst X == 0;\n"
                       "next X == if X eq 3 then 0 else (X + 1);\nY == X@1;\n", [""]),
    "probe_first_capture": ("var X : [0, 2];\nvar Y : [0, 2];\nY == first X;\nnext X != X;\n", [""]),
    "probe_first_expr": ("var X : [0, 2];\nvar Y : [0, 4];\nvar Z : [0, 1];\nY == first (X + Z) ;\n"
                         "next Z == 1 - Z;\n", [""]),
    "probe_first_gt_quirk": ("var X : [0, 3];\nvar Y : [0, 1];\nY == first (X gt 1);\nnext X == X;\n", [""]),
    "probe_first_or_quirk": ("var X : [0, 2];\nvar Y : [0, 2];\nvar Z : [0, 2];\nZ == first (X or Y);\n", [""]),
    "probe_array": ("arr T : {3, 1, 2};\nvar I : [0, 3];\nvar X : [0, 3];\nX == T[I];\nnext I != I;\n", [""]),
    "probe_array_short": ("arr T : {2, 0, 1};\nvar I : [0, 2];\nvar X : [0, 2];\nfirst I == 0;\n"
                          "X == T[I];\nnext I == T[I];\n", [""]),
    "probe_imply": ("var X : [0, 2];\nvar Y : [0, 2];\nX -> Y;\nnext X == Y;\n", [""]),
    "probe_not": ("var X : [0, 1];\nvar Y : [0, 1];\nY == not X;\nnext X == Y;\n", [""]),
    "probe_abs_neg": ("var X : [-2, 2];\nvar Y : [0, 2];\nY == abs X;\nnext X == 0 - X;\n", [""]),
    "probe_mul": ("var X : [0, 3];\nvar Y : [0, 3];\nvar Z : [0, 9];\nZ == X * Y;\nnext X == Y;\nZ >= 2;\n", [""]),
    "probe_fby": ("var X : [0, 3];\nX == 1 fby 2 fby X;\n", [""]),
    "probe_fby_expr": ("var X : [0, 3];\nvar Y : [0, 3];\nX == 0 fby (Y + 1);\nY == first 2;\nY < 3;\n", [""]),
    "probe_next_expr": ("var X : [0, 3];\nvar Y : [0, 3];\nnext (X + Y) <= 3;\nX <= Y;\nnext X >= X;\n", [""]),
    "probe_next_next": ("var X : [0, 2];\nfirst X == 0;\nnext next X == X;\n", [""]),
    "probe_two_next_same": ("var X : [0, 2];\nvar Y : [0, 2];\nvar Z : [0, 2];\nY == next X;\nZ == next X;\n"
                            "X != Y;\n", [""]),
    "probe_next_out_of_range": ("var X : [0, 3];\nvar Y : [0, 1];\nnext Y == X;\nnext X == X;\n", [""]),
    "probe_dead_branch": ("var X : [0, 3];\nfirst X <= 1;\nnext X == if X eq 0 then 0 else (X + 1);\n"
                          "X < 3;\n", [""]),
    "probe_k1": ("var X : [0, 3];\nfirst X <= 1;\nnext X == if X eq 0 then 0 else (X + 1);\nX < 3;\n",
                 ["-k1", "-k3"]),
    "probe_adversarial": ("var P0 : [0, 0];\nvar P1 : [0, 0];\nvar P2 : [0, 0];\nvar P3 : [0, 0];\n"
                          "var S : [0, 2];\nvar ADV : [0, 1];\nvar AVA : [0, 1];\nfirst S == 0;\n"
                          "next S == if ADV eq AVA then S else (S + 1);\nS < 2;\n", ["", "-a", "-z"]),
    "probe_adversarial_win": ("var P0 : [0, 0];\nvar P1 : [0, 0];\nvar P2 : [0, 0];\nvar P3 : [0, 0];\n"
                              "var S : [0, 2];\nvar ADV : [0, 1];\nvar AVA : [0, 2];\nfirst S == 0;\n"
                              "next S == if (AVA eq 2) or (ADV eq AVA) then S else (S + 1);\nS < 2;\n",
                              ["", "-a", "-z"]),
    "probe_comments": ("// leading comment\nvar X : [0, 2]; ' trailing quote comment\n/* block */\n"
                       "var Y : [0, 2];\nX + Y == 2;\nnext X == Y;\n", [""]),
    "probe_neg_const": ("var X : [-2, 2];\nfirst X == -2;\nnext X == if X eq 2 then -2 else (X + 1);\n", [""]),
}

DI_FLAGS = ["", "-a", "-z"]          # BASELINE.json config 2: digitinvader sweep incl. adversarial modes
TEXT_LIMIT = 400                     # keep the canonical text for automata up to this many lines


def cases():
    for name in instances.SHIPPED:
        text = instances.by_name(name)
        flagsets = DI_FLAGS if name.startswith("digitinvader") else [""]   # digitinvader1-9, BASELINE config 2
        for fl in flagsets:
            yield name, text, fl, False
    for name, (text, flagsets) in PROBES.items():
        for fl in flagsets:
            yield name, text, fl, True


def run_case(name, text, flags, is_probe, max_seconds):
    key = name + ("" if not flags else "_" + flags.strip("-"))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, name + ".csp")
        with open(path, "w") as f:
            f.write(text)
        argv = [REF, "-s"] + ([flags] if flags else []) + ["-m%d" % max_seconds, path]

        def unlimited_stack():
            resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
        t0 = time.time()
        p = subprocess.run(argv, cwd=d, capture_output=True, text=True, preexec_fn=unlimited_stack)
        wall = time.time() - t0
        rec = {"name": name, "flags": flags, "rc": p.returncode, "stdout": p.stdout, "wall_s": round(wall, 2)}
        if is_probe:
            rec["model"] = text
        dot = os.path.join(d, "solutions.dot")
        if p.returncode == 0 and os.path.exists(dot) and p.stdout.strip():
            a = canonical.parse_dot(open(dot).read())
            txt = canonical.canonical_text(a)
            st, ed = canonical.counts(a)
            rec.update(states=st, edges=ed, sha256=canonical.canonical_sha256(a),
                       header_vars=a.header_vars, header_sig=a.header_sig)
            if txt.count("\n") <= TEXT_LIMIT:
                rec["canonical"] = txt
            stat = [ln for ln in p.stdout.split("\n") if "\t" in ln]
            if stat:
                fld = stat[-1].split("\t")
                rec["stat"] = {"vars": int(fld[1]), "cons": int(fld[2]), "dominance": int(fld[3]),
                               "nodes": int(fld[4]), "fails": int(fld[5]), "solve_s": float(fld[6])}
        else:
            rec["error"] = "no automaton (rc=%d, timeout or fatal)" % p.returncode
            err = os.path.join(d, "error.txt")
            if os.path.exists(err):
                rec["error_txt"] = open(err).read()[-500:]
    with open(os.path.join(OUT, key + ".json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
        f.write("\n")
    return key, rec.get("states"), rec.get("edges"), rec.get("sha256", rec.get("error")), wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--only", default=".*")
    ap.add_argument("--max-seconds", type=int, default=5400)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    todo = [c for c in cases() if re.search(args.only, c[0] + ("" if not c[2] else "_" + c[2].strip("-")))]
    with cf.ThreadPoolExecutor(args.jobs) as ex:
        futs = [ex.submit(run_case, *c, args.max_seconds) for c in todo]
        for fu in cf.as_completed(futs):
            print(*fu.result(), flush=True)


if __name__ == "__main__":
    main()
