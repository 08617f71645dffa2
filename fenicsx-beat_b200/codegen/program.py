"""Time-stepping schemes over an :class:`OdeModel` -> a scheduled straight-line program.

Schemes
-------
``forward_explicit_euler``   y_new = y + dt*f                       (README.md:58-89 of the reference
                             writes the same update by hand for FitzHugh-Nagumo)
``generalized_rush_larsen``  first-order generalized Rush-Larsen (GRL1) as gotranx emits it for the
                             reference's demos (demos/niederer_benchmark.py:82-98): for each state y
                             with right-hand side f,
                                 lin = d f / d y   TOTAL derivative: the chain rule runs through every
                                                   intermediate that depends on y (other states held fixed)
                                 lin == 0          -> y + dt*f
                                 otherwise         -> y + f*(exp(lin*dt) - 1)/lin,
                                                      guarded by |lin| > 1e-8 (else dt*f) unless lin is a
                                                      fraction with a non-zero constant numerator.
                             gotranx itself is not installable here; which derivative it
                             takes was settled empirically: with the total derivative the oracle reproduces the
                             published Niederer activation times (demos/niederer_benchmark.py:315-325) within one
                             dt at dt = 0.05 (dx = 0.5 and 0.2), with intermediates held fixed (V then falls back
                             to forward Euler) it is 0.15-0.7 ms off.
"""

from __future__ import annotations

from dataclasses import dataclass, field

from . import ir
from .odefile import OdeModel

SCHEMES = ("forward_explicit_euler", "generalized_rush_larsen")
RL_DELTA = 1e-8


@dataclass
class Program:
    model: OdeModel
    scheme: str
    uniform: list[tuple[str, ir.Node]] = field(default_factory=list)  # parameter-only, topo order
    body: list[tuple[str, ir.Node]] = field(default_factory=list)  # node-dependent, topo order
    outputs: list[ir.Node] = field(default_factory=list)  # one per state, index order
    rl_states: list[str] = field(default_factory=list)
    fe_states: list[str] = field(default_factory=list)

    @property
    def used_parameters(self) -> list[str]:
        used: set[str] = set()
        for _, e in self.uniform + self.body:
            used |= ir.free_symbols(e)
        for e in self.outputs:
            used |= ir.free_symbols(e)
        return [p for p in self.model.parameters if p in used]


def build_program(model: OdeModel, scheme: str, reciprocal_constants: bool = True) -> Program:
    if scheme not in SCHEMES:
        raise ValueError(f"unknown scheme {scheme!r}; have {SCHEMES}")
    prog = Program(model=model, scheme=scheme)
    dt = ir.sym("dt")
    defs: dict[str, ir.Node] = dict(model.intermediates)
    roots: list[str] = []
    base_names = set(model.intermediates)
    dmemo: dict[tuple[str, str], ir.Node] = {}

    def chain(s: str):
        """d(intermediate)/d(state s) as a named intermediate d<name>_d<s> (defined once, shared by CSE)."""

        def sd(name: str) -> ir.Node:
            if name not in base_names:
                return ir.ZERO  # another state, a parameter, time
            key = (name, s)
            r = dmemo.get(key)
            if r is not None:
                return r
            dmemo[key] = ir.ZERO  # definitions are acyclic; this only guards against a malformed model
            dexpr = ir.diff(model.intermediates[name], s, sym_diff=sd)
            if ir.is_num(dexpr):
                r = dexpr
            else:
                dn = f"d{name}_d{s}"
                assert dn not in defs, dn
                defs[dn] = dexpr
                r = ir.sym(dn)
            dmemo[key] = r
            return r

        return sd

    for s in model.states:
        dname = f"d{s}_dt"
        f = model.derivatives[s]
        defs[dname] = f
        fsym = ir.sym(dname)
        y = ir.sym(s)
        if scheme == "forward_explicit_euler":
            prog.outputs.append(ir.add(y, ir.mul(dt, fsym)))
            prog.fe_states.append(s)
            roots.append(dname)
            continue
        lin = ir.diff(f, s, sym_diff=chain(s))
        if ir.is_num(lin, 0.0):
            prog.outputs.append(ir.add(y, ir.mul(dt, fsym)))
            prog.fe_states.append(s)
            roots.append(dname)
            continue
        lname = f"{dname}_linearized"
        defs[lname] = lin
        lsym = ir.sym(lname)
        rl = ir.div(ir.mul(fsym, ir.sub(ir.call("exp", ir.mul(lsym, dt)), ir.ONE)), lsym)
        if not ir.numerator_is_nonzero_constant(lin):
            rl = ir.cond(ir.cmp("gt", ir.call("abs", lsym), ir.num(RL_DELTA)), rl, ir.mul(dt, fsym))
        prog.outputs.append(ir.add(y, rl))
        prog.rl_states.append(s)
        roots += [dname, lname]

    # ---- dependency-ordered schedule, dead code dropped -------------------------
    params = set(model.parameters)
    order: list[str] = []
    state = {}  # name -> 0 visiting / 1 done

    def visit(name: str) -> None:
        st = state.get(name)
        if st == 1:
            return
        if st == 0:
            raise ValueError(f"cyclic definition through {name}")
        state[name] = 0
        for dep in sorted(ir.free_symbols(defs[name])):
            if dep in defs:
                visit(dep)
        state[name] = 1
        order.append(name)

    # visit in source order so the schedule is stable and readable
    needed: set[str] = set()
    stack = list(roots)
    while stack:
        n = stack.pop()
        if n in needed:
            continue
        needed.add(n)
        stack += [d for d in ir.free_symbols(defs[n]) if d in defs]
    for name in defs:
        if name in needed:
            visit(name)
    # derivative intermediates d<u>_d<s> right after their base u: they share most operands with it, so the values are
    # still in registers (emitted after ALL base intermediates they cost ToR-ORd kilobytes of spills)
    derived_of: dict[str, list[str]] = {}
    for (base, s_), node in dmemo.items():
        if node.kind == "sym" and node.value == f"d{base}_d{s_}" and node.value in needed:
            derived_of.setdefault(base, []).append(node.value)
    all_derived = {d for lst in derived_of.values() for d in lst}
    reordered: list[str] = []
    for name in order:
        if name in all_derived:
            continue
        reordered.append(name)
        reordered += derived_of.get(name, [])
    assert sorted(reordered) == sorted(order)
    order = reordered

    uniform_names: set[str] = set()
    for name in order:
        deps = ir.free_symbols(defs[name])
        if all((d in params) or (d in uniform_names) for d in deps):
            uniform_names.add(name)
            prog.uniform.append((name, defs[name]))
        else:
            prog.body.append((name, defs[name]))

    # ---- hoist maximal parameter-only sub-expressions out of the node-dependent part ----------
    # (e.g. R*T/F, sqrt(K_o/5.4) in TP06): evaluated once on the host per parameter set, or once per
    # thread when parameters are per-node.  Evaluation order inside the hoisted tree is unchanged.
    fs_memo: dict = {}
    hoisted: dict[int, ir.Node] = {}

    def leaf_map(x: ir.Node):
        if x.kind in ("num", "sym"):
            return None
        if x.kind in ("lt", "gt", "le", "ge", "eq", "ne", "and", "or"):
            return None  # booleans are not stored as doubles
        syms = ir.free_symbols(x, fs_memo)
        if syms and all((s in params) or (s in uniform_names) for s in syms):
            h = hoisted.get(id(x))
            if h is None:
                name = f"_u{len(hoisted)}"
                h = ir.sym(name)
                hoisted[id(x)] = h
                prog.uniform.append((name, x))
                uniform_names.add(name)
            return h
        return None

    memo: dict = {}
    prog.body = [(n, ir.rebuild(e, leaf_map, memo)) for n, e in prog.body]
    prog.outputs = [ir.rebuild(e, leaf_map, memo) for e in prog.outputs]
    if reciprocal_constants:
        _divisions_by_constants_to_multiplications(prog, params, uniform_names)
        _shared_denominators_to_reciprocals(prog)
    return prog


def _shared_denominators_to_reciprocals(prog: Program) -> None:
    """a/b, c/b, ... with the SAME node-dependent denominator b  ->  r = 1/b once, then a*r, c*r, ...  The total-
    derivative Rush-Larsen rule produces many of these (f = (y_inf - y)/tau with lin = -1/tau; the quotient rule divides
    by the denominator of the quotient it differentiates): 78 of TP06's 136 divisions share 35 denominators.  On the
    device a reciprocal is a seed + two Newton steps (RCP, csrc/ode_math.cuh), so a group of n divisions costs 6 + n
    instructions instead of 8 n.  Each product is within 1.5 ulp of the quotient.  DEVICE code only."""
    seen: set[int] = set()
    uses: dict[int, int] = {}

    def count(x: ir.Node) -> None:
        if id(x) in seen:
            return
        seen.add(id(x))
        for a in x.args:
            count(a)
        if x.kind == "div" and not ir.is_num(x.args[1]) and not ir.is_num(x.args[0], 1.0):
            uses[id(x.args[1])] = uses.get(id(x.args[1]), 0) + 1

    for _, e in prog.body:
        count(e)
    for e in prog.outputs:
        count(e)
    shared = {k for k, n in uses.items() if n >= 2}
    memo: dict[int, ir.Node] = {}

    def rec(x: ir.Node) -> ir.Node:
        r = memo.get(id(x))
        if r is not None:
            return r
        if not x.args:
            r = x
        else:
            old_den = x.args[1] if x.kind == "div" else None
            a = tuple(rec(c) for c in x.args)
            r = x if all(p is q for p, q in zip(a, x.args)) else ir._mk(x.kind, a, x.value)
            if old_den is not None and id(old_den) in shared and not ir.is_num(a[0], 1.0):
                r = ir.mul(a[0], ir.div(ir.ONE, a[1]))  # div(ONE, b) is hash-consed: one reciprocal per denominator
        memo[id(x)] = r
        return r

    prog.body = [(n, rec(e)) for n, e in prog.body]
    prog.outputs = [rec(e) for e in prog.outputs]


def _divisions_by_constants_to_multiplications(prog: Program, params: set, uniform_names: set) -> None:
    """x / c  ->  x * (1/c) when c is a literal or a per-launch constant (a parameter or a hoisted parameter-only
    value): the reciprocal is folded at generation time (literal) or becomes one more hoisted constant.  A fp64
    divide is ~8 fp64-pipe instructions on the device (csrc/ode_math.cuh), a multiply one; about a third of the
    divisions of TP06 / ToR-ORd have such a divisor.  The product differs from the correctly rounded quotient by at
    most 1 ulp - inside the budget the device division already has, and far inside the 1e-12 parity bar.  DEVICE
    code only: the oracle keeps true divisions (it is generated independently, oracle/gen_models.py)."""
    memo: dict = {}
    rcp_of: dict[str, ir.Node] = {}

    def rec(x: ir.Node) -> ir.Node:
        r = memo.get(id(x))
        if r is not None:
            return r
        if not x.args:
            r = x
        else:
            a = tuple(rec(c) for c in x.args)
            r = x if all(p is q for p, q in zip(a, x.args)) else ir._mk(x.kind, a, x.value)
            if r.kind == "div":
                top, den = r.args
                if ir.is_num(den) and den.value != 0.0:
                    r = ir.mul(top, ir.num(1.0 / den.value))
                elif den.kind == "sym" and (den.value in params or den.value in uniform_names) and not ir.is_num(top):
                    h = rcp_of.get(den.value)
                    if h is None:
                        name = f"_r{len(rcp_of)}"
                        prog.uniform.append((name, ir.div(ir.ONE, den)))
                        uniform_names.add(name)
                        h = rcp_of[den.value] = ir.sym(name)
                    r = ir.mul(top, h)
        memo[id(x)] = r
        return r

    prog.body = [(n, rec(e)) for n, e in prog.body]
    prog.outputs = [rec(e) for e in prog.outputs]


def op_counts(prog: Program, include_uniform: bool = False) -> dict[str, int]:
    """Arithmetic per node-step after common-subexpression sharing (the DAG is hash-consed, so each
    distinct node is counted once).  Used for the fp64 roofline in DESIGN.md / bench.py."""
    seen: set[int] = set()
    counts = {k: 0 for k in ("add", "mul", "div", "exp", "log", "sqrt", "pow", "floor", "abs", "cmp", "select", "neg")}

    def rec(x: ir.Node) -> None:
        if id(x) in seen:
            return
        seen.add(id(x))
        for a in x.args:
            rec(a)
        k = x.kind
        if k in ("add", "sub"):
            counts["add"] += 1
        elif k == "mul":
            counts["mul"] += 1
        elif k == "div":
            counts["div"] += 1
        elif k == "neg":
            counts["neg"] += 1
        elif k == "pow":
            b = x.args[1]
            if ir.is_num(b) and float(b.value).is_integer() and 2 <= abs(b.value) <= 16:
                n = int(abs(b.value))
                counts["mul"] += n.bit_length() - 1 + bin(n).count("1") - 1
                if b.value < 0:
                    counts["div"] += 1
            else:
                counts["pow"] += 1
        elif k == "call":
            counts[x.value] += 1
        elif k == "cond":
            counts["select"] += 1
        elif k in ("lt", "gt", "le", "ge", "eq", "ne", "and", "or"):
            counts["cmp"] += 1

    exprs = [e for _, e in prog.body] + list(prog.outputs)
    if include_uniform:
        exprs += [e for _, e in prog.uniform]
    for e in exprs:
        rec(e)
    return counts
