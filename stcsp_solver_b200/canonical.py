"""Canonical form of a solution automaton.

The reference numbers vertices by a DFS driven by ``__gnu_cxx::hash_map`` iteration
(reference src/graph.cpp:41-76, 420-442) and numbers constraint sets in DFS discovery
order (src/solveralgorithm.cpp:792-803); both are artefacts of search order.  Two automata
are the same iff their canonical texts are equal:

* BFS from the root with a FIFO queue, out-edges visited in ascending ``(label, dst)``
  order, vertices numbered in first-visit order;
* constraint-set ids renumbered by first appearance along the canonical vertex order;
* text = the two DOT header lines (all variable names / signature variable names), one
  ``V`` line per vertex, then the edges of each vertex in canonical order, sorted by label.

Only vertices reachable from the root through printed edges are part of the automaton.
"""
from __future__ import annotations

import hashlib
import re
from collections import deque
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass
class Automaton:
    """Plain automaton: vertex ids are arbitrary hashables, ``root`` may be None (empty)."""
    header_vars: str = "#"            # DOT header line 2, e.g. "# A B0 B1"
    header_sig: str = "#"             # DOT header line 3
    root: Optional[object] = None
    final: Dict[object, bool] = field(default_factory=dict)
    cset: Dict[object, int] = field(default_factory=dict)
    sig: Dict[object, Optional[Tuple[int, ...]]] = field(default_factory=dict)   # None == root "S"
    edges: Dict[object, List[Tuple[Tuple[int, ...], object]]] = field(default_factory=dict)

    @property
    def n_states(self) -> int:
        return len(self.final)

    @property
    def n_edges(self) -> int:
        return sum(len(v) for v in self.edges.values())


_VERTEX = re.compile(r'^(\d+) \[shape=(doublecircle|circle), label="(-?\d+): ([^"]*)"\];$')
_EDGE = re.compile(r'^(\d+) -> (\d+) \[label="([^"]*)"\];$')


def parse_dot(text: str) -> Automaton:
    """Parse a ``solutions.dot`` written by the reference (src/solveralgorithm.cpp:709-730)."""
    lines = text.split("\n")
    if len(lines) < 4 or not lines[3].startswith("digraph"):
        raise ValueError("not a solutions.dot file")
    a = Automaton(header_vars=lines[1], header_sig=lines[2])
    for ln in lines[4:]:
        if ln == "}" or ln == "":
            continue
        m = _VERTEX.match(ln)
        if m:
            vid = int(m.group(1))
            a.final[vid] = m.group(2) == "doublecircle"
            a.cset[vid] = int(m.group(3))
            body = m.group(4)
            if body == "S":
                a.sig[vid] = None
                a.root = vid
            else:
                a.sig[vid] = tuple(int(x) for x in body.split(", ")) if body else tuple()
            a.edges.setdefault(vid, [])
            continue
        m = _EDGE.match(ln)
        if m:
            src, dst = int(m.group(1)), int(m.group(2))
            label = tuple(int(x) for x in m.group(3).split(", ")) if m.group(3) else tuple()
            a.edges.setdefault(src, []).append((label, dst))
            continue
        raise ValueError("unrecognised DOT line: %r" % ln)
    return a


def canonical_text(a: Automaton) -> str:
    """Canonical text rendering; the literal token ``EMPTY`` for an automaton without root."""
    if a.root is None:
        return "EMPTY"
    order: List[object] = []
    number: Dict[object, int] = {a.root: 0}
    queue = deque([a.root])
    while queue:
        v = queue.popleft()
        order.append(v)
        for label, dst in sorted(a.edges.get(v, []), key=lambda e: (e[0], number.get(e[1], 1 << 62))):
            if dst not in number:
                number[dst] = len(number)
                queue.append(dst)
    cmap: Dict[int, int] = {}
    out = [a.header_vars, a.header_sig]
    for v in order:
        c = cmap.setdefault(a.cset[v], len(cmap))
        s = a.sig[v]
        body = "S" if s is None else " ".join(str(x) for x in s)
        out.append("V %d %s %d %s" % (number[v], "F" if a.final[v] else "N", c, body))
    for v in order:
        es = sorted(a.edges.get(v, []), key=lambda e: e[0])
        labels = [e[0] for e in es]
        if len(set(labels)) != len(labels):
            raise ValueError("duplicate edge labels out of one vertex; canonical order is not total")
        for label, dst in es:
            out.append("E %d %d %s" % (number[v], number[dst], " ".join(str(x) for x in label)))
    return "\n".join(out) + "\n"


def canonical_sha256(a: Automaton) -> str:
    return hashlib.sha256(canonical_text(a).encode("utf-8")).hexdigest()


def counts(a: Automaton) -> Tuple[int, int]:
    """(states, edges) of the part reachable from the root."""
    if a.root is None:
        return 0, 0
    seen = {a.root}
    queue = deque([a.root])
    n_edges = 0
    while queue:
        v = queue.popleft()
        for _, dst in a.edges.get(v, []):
            n_edges += 1
            if dst not in seen:
                seen.add(dst)
                queue.append(dst)
    return len(seen), n_edges


def from_arrays(var_names: Sequence[str], sig_names: Sequence[str], n_states: int, root: int,
                final: Sequence[int], cset: Sequence[int], sig_len: int, sig: Sequence[int],
                edge_src: Sequence[int], edge_dst: Sequence[int], n_vars: int,
                edge_label: Sequence[int], root_valid: bool = True) -> Automaton:
    """Build an :class:`Automaton` from the flat arrays of ``stcsp_automaton_t`` (include/stcsp_b200.h)."""
    a = Automaton(header_vars="#" + "".join(" " + n for n in var_names),
                  header_sig="#" + "".join(" " + n for n in sig_names))
    if not root_valid:
        return a
    a.root = root
    for s in range(n_states):
        a.final[s] = bool(final[s])
        a.cset[s] = int(cset[s])
        a.sig[s] = None if s == root else tuple(int(x) for x in sig[s * sig_len:(s + 1) * sig_len])
        a.edges[s] = []
    for e in range(len(edge_src)):
        a.edges[int(edge_src[e])].append(
            (tuple(int(x) for x in edge_label[e * n_vars:(e + 1) * n_vars]), int(edge_dst[e])))
    return a
