#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- goldens for instances the reference cannot finish.

The reference needs weeks for juggling_b8_f8_nosym and terabytes for partialorder_16+ (SURVEY.md
section 0), so these goldens come from oracle/semantic_oracle.cpp -- the automaton computed from
its definition -- AFTER that oracle has been pinned against every golden the reference itself
produced (tests/test_semantic_oracle.py).  Files are written as tests/golden/semantic_<name>.json
with ``"source": "semantic_oracle"`` so that nobody mistakes them for reference output.

Usage:  python oracle/make_semantic_goldens.py [--only REGEX] [--jobs N]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json config 5 (synthetic scaled juggling b8/f8 and partialorder_20) and the steps towards it,
# SURVEY.md section 8(d) "Concrete inputs" (5).
NAMES = (["juggling_b%d_f%d_nosym" % bf for bf in [(6, 7), (6, 8), (7, 7), (7, 8), (8, 8)]]
         + ["partialorder_%d" % n for n in range(15, 21)]
         + ["digitinvader%d" % n for n in (10, 11, 12)])
TEXT_LIMIT = 400


def run(name):
    import _oracle
    from stcsp_solver_b200 import binding, instances
    model = binding.Model(instances.by_name(name))
    r = _oracle.semantic(model, want_text=False)
    rec = {"name": name, "flags": "", "source": "semantic_oracle", "states": r["n_states"], "edges": r["n_edges"],
           "sha256": r["sha256"], "oracle_seconds": round(r["seconds"], 2), "oracle_table_states": r["n_table_states"],
           "constraint_sets": r["n_constraint_sets"]}
    if r["n_states"] + r["n_edges"] + 2 <= TEXT_LIMIT:
        rec["canonical"] = _oracle.semantic(model, want_text=True)["text"]
    with open(os.path.join(ROOT, "tests", "golden", "semantic_" + name + ".json"), "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
        f.write("\n")
    return name, rec["states"], rec["edges"], rec["sha256"], rec["oracle_seconds"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=".*")
    ap.add_argument("--jobs", type=int, default=2)
    args = ap.parse_args()
    todo = [n for n in NAMES if re.search(args.only, n)]
    with cf.ProcessPoolExecutor(args.jobs) as ex:
        for res in ex.map(run, todo):
            print(*res, flush=True)


if __name__ == "__main__":
    main()
