"""Seeded random .csp models for differential testing (CUDA path vs the oracle).

Small variable counts and domains keep the oracle fast; the grammar coverage is what matters: every pointwise
operator, arrays, if-then-else, `->`, `not`, and the temporal forms next / first / fby / @ / until."""
import random


def _expr(rng, names, arrays, depth, boolean=False, div=True):
    """div=False: no `/` or `%` -- where the normaliser names the expression by an auxiliary variable, a division
    gives that variable the domain [INT_MIN, INT_MAX] (reference src/solveralgorithm.cpp:316-322) and the reference's
    support search never ends."""
    if depth <= 0 or rng.random() < 0.3:
        if names and rng.random() < 0.75:
            return rng.choice(names)
        return str(rng.randint(-2, 3))
    kind = rng.random()
    a = lambda b=False: _expr(rng, names, arrays, depth - 1, b, div)
    if boolean or kind < 0.30:
        op = rng.choice(["lt", "gt", "le", "ge", "eq", "ne", "and", "or"])
        if op in ("and", "or"):
            return "(%s %s %s)" % (a(True), op, a(True))
        return "(%s %s %s)" % (a(), op, a())
    if kind < 0.60:
        return "(%s %s %s)" % (a(), rng.choice(["+", "-", "+", "-", "*"]), a())
    if kind < 0.68:
        return "(%s %s %s)" % (a(), rng.choice(["/", "%"]) if div else "-", a())
    if kind < 0.76:
        return "(abs %s)" % a()
    if kind < 0.86:
        return "(if %s then %s else %s)" % (a(True), a(), a())
    if kind < 0.92 and arrays:
        return "%s[%s]" % (rng.choice(arrays), a())
    return "(not %s)" % a(True) if rng.random() < 0.5 else a()


def dynamic_model(seed):
    """Models that are satisfiable by construction and have real dynamics: every variable gets an in-range successor
    rule (modular arithmetic over non-negative domains) or a nondeterministic one, plus a few side constraints."""
    rng = random.Random(1000003 * seed + 17)
    n = rng.randint(2, 4)
    ub = [rng.randint(1, 4) for _ in range(n)]
    names = ["X%d" % i for i in range(n)]
    lines = ["var X%d : [0, %d];" % (i, ub[i]) for i in range(n)]
    use_until = rng.random() < 0.35
    if use_until:
        lines += ["var F0 : [0, 1];", "var F1 : [0, 1];"]
    for i in range(n):
        j, w = rng.randrange(n), ub[i] + 1
        form = rng.random()
        if form < 0.45:
            lines.append("next X%d == (X%d + X%d * %d + %d) %% %d;" % (i, i, j, rng.randint(0, 2), rng.randint(0, 2), w))
        elif form < 0.6:
            lines.append("next X%d == if X%d eq %d then 0 else (X%d + 1);" % (i, i, ub[i], i))
        elif form < 0.75:
            lines.append("next X%d %s X%d;" % (i, rng.choice(["!=", ">=", "<="]), j))
        elif form < 0.85:
            lines.append("X%d == %d fby (if X%d eq %d then 0 else (X%d + 1));" % (i, rng.randint(0, ub[i]), i, ub[i], i))
        # else: unconstrained successor
    for _ in range(rng.randint(0, 2)):
        a, b = rng.sample(names, 2)
        lines.append(rng.choice(["%s != %s;" % (a, b), "%s + %s <= %d;" % (a, b, rng.randint(2, 6)),
                                 "(%s lt %s) -> (%s le 1);" % (a, b, a),
                                 "(if %s gt %s then %s else %s) >= 1;" % (a, b, a, b)]))
    if rng.random() < 0.5:
        i = rng.randrange(n)
        lines.append("first X%d %s %d;" % (i, rng.choice(["==", "<=", ">="]), rng.randint(0, ub[i])))
    if rng.random() < 0.25:
        a, b = rng.sample(names, 2)
        if ub[int(a[1:])] >= ub[int(b[1:])]:
            lines.append("%s == %s@%d;" % (a, b, rng.randint(1, 2)))
    if use_until:
        lines.append("F1 == (X0 eq %d);" % rng.randint(0, ub[0]))
        lines.append("F0 == (X1 le %d);" % rng.randint(0, ub[1]))
        lines.append("F0 until F1;")
    return "\n".join(lines) + "\n"


def random_model(seed):
    if seed >= 300:
        return dynamic_model(seed)
    rng = random.Random(seed)
    lines, names, bounds = [], [], {}
    for i in range(rng.randint(2, 4)):
        lb = rng.randint(-2, 1)
        ub = lb + rng.randint(0, 3)
        names.append("X%d" % i)
        bounds[names[-1]] = (lb, ub)
        lines.append("var X%d : [%d, %d];" % (i, lb, ub))
    flags = []
    if rng.random() < 0.3:
        for i in range(2):
            flags.append("F%d" % i)
            lines.append("var F%d : [0, 1];" % i)
    arrays = []
    if rng.random() < 0.4:
        arrays.append("T")
        lines.append("arr T : {%s};" % ", ".join(str(rng.randint(-1, 3)) for _ in range(rng.randint(1, 4))))
    allv = names + flags
    ops = ["<", ">", "<=", ">=", "==", "!=", "->"]
    for _ in range(rng.randint(1, 3)):                          # pointwise constraints
        lines.append("%s %s %s;" % (_expr(rng, allv, arrays, 2), rng.choice(ops), _expr(rng, allv, arrays, 2)))
    for _ in range(rng.randint(0, 2)):                          # transitions
        x = rng.choice(names)
        form = rng.random()
        if form < 0.6:
            lines.append("next %s %s %s;" % (x, rng.choice(["==", "==", ">=", "!=", "<="]), _expr(rng, allv, arrays, 2)))
        elif form < 0.8:
            lines.append("%s == %s fby %s;" % (x, rng.randint(*bounds[x]), _expr(rng, allv, arrays, 1, div=False)))
        else:
            lines.append("next next %s == %s;" % (x, rng.choice(names)))
    if rng.random() < 0.5:                                      # initial conditions
        x = rng.choice(names)
        lines.append("first %s %s %d;" % (x, rng.choice(["==", "<=", ">="]), rng.randint(*bounds[x])))
    if rng.random() < 0.2:
        y, x = rng.choice(names), rng.choice(names)
        lines.append("%s == %s@%d;" % (y, x, rng.randint(1, 2)))
    if rng.random() < 0.2:
        lines.append("%s == first (%s);" % (rng.choice(names), _expr(rng, names, arrays, 1)))
    if flags and rng.random() < 0.7:
        lines.append("F0 until F1;")
        if rng.random() < 0.5:
            lines.append("F1 == (%s);" % _expr(rng, names, arrays, 1, True))
    return "\n".join(lines) + "\n"
