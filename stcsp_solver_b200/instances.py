"""Deterministic text templates for the three benchmark families.

The reference ships its benchmark models as ``examples/*.csp`` (reference
examples/juggling_b6_f6.csp:1-45, examples/partialorder_14.csp:1-51,
examples/digitinvader9.csp:1-36).  ``/root/reference`` does not exist on the GPU box, so the
same models are regenerated here from templates: statement order is the shipped files' order
(it defines the auxiliary-variable numbering and therefore the edge-label columns), and the
generated text is checked against the reference's golden automaton hashes in
``tests/test_golden.py``.  The same templates give the scaled synthetic instances
(``juggling_b8_f8_nosym``, ``partialorder_20``) named in BASELINE.json.
"""
from __future__ import annotations

import re
from typing import List

# The one shipped juggling file that puts the symmetry-breaking ``first`` block LAST
# (reference examples/juggling_b4_f6.csp:22-25); the others put it right after the declarations.
_FIRST_BLOCK_LAST = {(4, 6)}


def juggling(balls: int, height: int, sym: bool = True, first_block_last: bool | None = None) -> str:
    """``juggling_b{balls}_f{height}[_nosym]``: A is the throw, B_i the landing times."""
    if first_block_last is None:
        first_block_last = (balls, height) in _FIRST_BLOCK_LAST
    out: List[str] = ["var A : [0, %d];" % height]
    out += ["var B%d : [0, %d];" % (i, height) for i in range(balls)]
    out.append("")
    first = ["first B0 == 1;"] + ["first B%d < first B%d;" % (i, i + 1) for i in range(balls - 1)]
    if sym and not first_block_last:
        out += first + [""]
    out += ["next B%d == if B%d eq 1 then A else (B%d - 1);" % (i, i, i) for i in range(balls)]
    out.append("")
    out += ["B%d != B%d;" % (i, j) for i in range(balls) for j in range(i + 1, balls)]
    out.append("")
    chain = ["A == if B0 eq 1 then next B0"]
    chain += ["else if B%d eq 1 then next B%d" % (i, i) for i in range(1, balls)]
    chain.append("else 0;")
    out += chain
    out.append("")
    if sym and first_block_last:
        out += first
    return "\n".join(out) + "\n"


def partialorder(n: int) -> str:
    """``partialorder_{n}``: give n items in some order; ``succ`` once all have been seen."""
    out = ["var succ : [0, 1];", "var giveTo : [0, %d];" % (n - 1)]
    out += ["var seen%d : [0, 1];" % i for i in range(n)]
    out.append("")
    out.append("first giveTo < %d;" % ((n - 1) // 2))
    for i in range(n):
        out.append("first seen%d == 0;" % i)
        out.append("next seen%d == seen%d or (giveTo eq %d);" % (i, i, i))
    out.append("")
    out.append("first succ == 0;")
    out.append("")
    out.append("succ >= (" + " and ".join("seen%d" % i for i in range(n)) + ");")
    out.append("next succ >= succ;")
    return "\n".join(out) + "\n"


def digitinvader(n: int) -> str:
    """``digitinvader{n}``: digits 0..n march in through D5; the player shoots digit I."""
    out = ["var I : [0, %d];" % n]
    out += ["var D%d : [-1, %d];" % (i, n) for i in range(6)]
    out += ["var A%d : [0, 1];" % i for i in range(6)]
    out += ["var MISS : [0, 1];", "var GAMEOVER : [0, 1];", ""]
    out += ["first D%d == -1;" % i for i in range(5)]
    out.append("D5 == " + " fby ".join(str(i) for i in range(n + 1)) + " fby D5;")
    for i in range(6):
        terms = ["I ne D%d" % j for j in range(i)] + ["I eq D%d" % i]
        out.append("A%d == %s;" % (i, " and ".join(terms)))
    out.append("")
    out.append("MISS == (" + " + ".join("A%d" % i for i in range(6)) + ") eq 0;")
    out.append("GAMEOVER == D0 ne -1 and MISS;")
    out.append("")
    for i in range(5):
        hit = " or ".join(["MISS"] + ["A%d" % j for j in range(i + 1)])
        out.append("next D%d == if GAMEOVER then -1 else if %s then D%d else D%d;" % (i, hit, i + 1, i))
    return "\n".join(out)


_NAME = re.compile(r"^(juggling_b(\d+)_f(\d+)(_nosym)?|partialorder_(\d+)|digitinvader(\d+))$")


def by_name(name: str) -> str:
    """Instance text for a benchmark name such as ``juggling_b6_f6_nosym`` or ``partialorder_14``."""
    m = _NAME.match(name)
    if not m:
        raise KeyError("unknown instance name %r" % name)
    if m.group(2):
        return juggling(int(m.group(2)), int(m.group(3)), sym=m.group(4) is None)
    if m.group(5):
        return partialorder(int(m.group(5)))
    return digitinvader(int(m.group(6)))


SHIPPED = (["digitinvader%d" % i for i in range(1, 10)]
           + ["juggling_b%d_f%d%s" % (b, f, s) for (b, f) in [(4, 4), (4, 5), (4, 6), (5, 5), (5, 6), (6, 6)]
              for s in ("", "_nosym")]
           + ["partialorder_%d" % i for i in range(10, 15)])
