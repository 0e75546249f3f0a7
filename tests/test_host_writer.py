"""The host side of a result (csrc/host/postprocess.cpp): canonical numbering, solutions.dot, canonical text and its SHA-256
are produced by several host threads and, where the CPU has them, the x86 SHA extensions.  One thread / many threads and
portable / accelerated SHA-256 must give the same bytes, and those must be the reference's (golden SHA, hashlib)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

from conftest import GOLDENS, ROOT, golden_text

SCRIPT = r"""
import hashlib, json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from stcsp_solver_b200 import binding, instances
import _oracle
out = {}
for name, text in %(models)r:
    m = binding.Model(text)
    a, _ = _oracle.solve(m)
    sol = binding.Solution(m, a)
    text = sol.canonical_text()
    out[name] = {"streamed": sol.canonical_sha256_streamed(), "hashlib": hashlib.sha256(text.encode()).hexdigest(),
                 "dot": hashlib.sha256(sol.dot().encode()).hexdigest(), "states": int(sol.n_states), "edges": int(sol.n_edges)}
    path = %(tmp)r + "/" + name + ".dot"
    sol.write_dot(path)
    out[name]["dot_file"] = hashlib.sha256(open(path, "rb").read()).hexdigest()
print(json.dumps(out))
"""

NAMES = ["juggling_b4_f6_nosym", "partialorder_10", "digitinvader3", "probe_until_two"]     # (partialorder_10: 28 778 edges = two text chunks)


def run(tmp_path, **env):
    e = dict(os.environ)
    e.update(env)
    models = [(n, golden_text(GOLDENS[n])) for n in NAMES]
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "models": models, "tmp": str(tmp_path)}], env=e,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().split("\n")[-1])


def test_threads_and_sha_paths_give_the_same_bytes(tmp_path):
    import concurrent.futures as cf
    envs = [dict(STCSP_HOST_THREADS="1", STCSP_NO_SHA_NI="1"), dict(STCSP_HOST_THREADS="8"), dict(STCSP_HOST_THREADS="3", STCSP_NO_SHA_NI="1")]
    dirs = []
    for i in range(len(envs)):
        d = tmp_path / ("run%d" % i)
        d.mkdir()
        dirs.append(d)
    with cf.ThreadPoolExecutor(len(envs)) as ex:
        results = list(ex.map(lambda a: run(a[0], **a[1]), zip(dirs, envs)))
    base = results[0]
    for env, r in zip(envs[1:], results[1:]):
        assert r == base, env
    for name in NAMES:
        r = base[name]
        g = GOLDENS[name]
        assert r["streamed"] == r["hashlib"] == g["sha256"], name
        assert (r["states"], r["edges"]) == (g["states"], g["edges"]), name
        assert r["dot"] == r["dot_file"], name
