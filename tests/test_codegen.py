"""The device code generator (fenicsx-beat_b200/codegen) against the independent SymPy-generated oracle models, on the
CPU: the scheduled program the CUDA emitter prints is interpreted with NumPy (tests/_ir_eval.py) and must reproduce the
oracle step to rounding, for every model and scheme, shared and per-node parameters.  Also: the committed generated
files are what the generator produces from the current sources (needs /root/reference for the .ode inputs)."""
import os
import sys

import numpy as np
import pytest

import _ir_eval
import _problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fenicsx-beat_b200"))
from codegen import generate, program  # noqa: E402

ODES = "/root/reference/odes"
needs_reference = pytest.mark.skipif(not os.path.isdir(ODES), reason="the .ode sources live under /root/reference (not on the GPU box)")


def _model(tag):
    if generate.MODELS[tag][1] is None or os.path.isdir(ODES):
        return generate.load_model(tag, ODES)[0]
    pytest.skip("needs /root/reference/odes")


@pytest.mark.parametrize("tag", ["fhn", "tp06", "torord"])
@pytest.mark.parametrize("scheme", ["forward_explicit_euler", "generalized_rush_larsen"])
def test_program_matches_oracle_model(tag, scheme):
    model = _model(tag)
    om = P.oracle_model(tag)
    assert list(model.states) == list(om.state) and list(model.parameters) == list(om.parameter)  # same index order
    prog = program.build_program(model, scheme)
    rng = np.random.default_rng(42)
    states = P.perturbed_states(om, 2000, rng, P.V_NAME[tag])
    p0 = om.init_parameter_values()
    for params in (p0, np.repeat(p0[:, None], 2000, axis=1) * (1 + 0.02 * rng.uniform(-1, 1, (len(p0), 2000)))):
        for t in (0.0, 10.5):
            got = _ir_eval.evaluate(prog, states, t, 0.01, params)
            with np.errstate(all="ignore"):
                want = getattr(om, scheme)(states, t, 0.01, params)
            scale = np.maximum(np.maximum(np.abs(want), np.abs(states)), 1e-6 * np.abs(want).max(axis=1, keepdims=True) + 1e-300)
            err = (np.abs(got - want) / scale).max()
            # forward Euler: a few ulp (reciprocal pass); Rush-Larsen: exp(x) - 1 cancellation, see tests/test_gpu_parity.py
            assert err <= (1e-12 if scheme == "forward_explicit_euler" else 2e-11), (tag, scheme, t, err)


def test_rush_larsen_uses_the_total_derivative():
    """TP06: V and the concentrations have no explicit self-dependence in their written right-hand sides; the chain rule
    through the intermediates makes all 19 states Rush-Larsen (what reproduces the published Niederer table)."""
    prog = program.build_program(_model("tp06"), "generalized_rush_larsen")
    assert prog.fe_states == [] and len(prog.rl_states) == 19
    names = {n for n, _ in prog.body}
    assert "dV_dt_linearized" in names and any(n.startswith("di_") and n.endswith("_dV") for n in names)
    opaque = program.build_program(_model("fhn"), "forward_explicit_euler")
    assert opaque.rl_states == []


@needs_reference
def test_committed_generated_files_are_current():
    assert generate.main(["--check", "--odes", ODES]) == 0


def _eval_node(x, env, memo=None):
    """Scalar evaluation of one IR node (the same operations tests/_ir_eval.py interprets, on Python floats)."""
    import math

    memo = {} if memo is None else memo
    r = memo.get(id(x))
    if r is not None:
        return r
    k, a = x.kind, x.args
    ev = lambda n: _eval_node(n, env, memo)  # noqa: E731
    if k == "num":
        r = x.value
    elif k == "sym":
        v = env[x.value]
        r = ev(v) if hasattr(v, "kind") else v
    elif k == "add":
        r = ev(a[0]) + ev(a[1])
    elif k == "sub":
        r = ev(a[0]) - ev(a[1])
    elif k == "mul":
        r = ev(a[0]) * ev(a[1])
    elif k == "div":
        r = ev(a[0]) / ev(a[1])
    elif k == "neg":
        r = -ev(a[0])
    elif k == "pow":
        r = ev(a[0]) ** ev(a[1])
    elif k == "call":
        r = {"exp": math.exp, "log": math.log, "sqrt": math.sqrt, "floor": math.floor, "abs": abs}[x.value](ev(a[0]))
    elif k == "cond":
        r = ev(a[1]) if ev(a[0]) else ev(a[2])
    elif k in ("lt", "gt", "le", "ge"):
        p, q = ev(a[0]), ev(a[1])
        r = {"lt": p < q, "gt": p > q, "le": p <= q, "ge": p >= q}[k]
    else:
        raise NotImplementedError(k)
    memo[id(x)] = r
    return r


def test_differentiation_against_finite_differences_on_random_expressions():
    """ir.diff - product / both quotient forms / power / exp / log / sqrt / abs / Conditional, and the chain-rule hook
    through named intermediates that the total-derivative Rush-Larsen step relies on - against a 4th-order central
    difference on 300 random expression DAGs."""
    import random

    from codegen import ir

    rnd = random.Random(7)
    y, p, u = ir.sym("y"), ir.sym("p"), ir.sym("u")  # state, parameter, named intermediate u = u(y)
    u_def = ir.call("exp", ir.mul(ir.num(0.3), y)) + ir.mul(p, y)

    def leaf():
        return rnd.choice([y, y, p, u, ir.num(rnd.choice([0.5, 1.0, 2.0, -1.5, 3.0]))])

    def positive(e):  # keep log / sqrt / non-integer powers in their domain
        return ir.add(ir.mul(e, e), ir.num(0.7))

    def tree(depth):
        if depth == 0:
            return leaf()
        a, b = tree(depth - 1), tree(depth - 1)
        op = rnd.randrange(11)
        if op == 0:
            return ir.add(a, b)
        if op == 1:
            return ir.sub(a, b)
        if op == 2:
            return ir.mul(a, b)
        if op == 3:
            return ir.div(a, positive(b))
        if op == 4:
            return ir.div(ir.num(rnd.choice([1.0, -2.0, 0.25])), positive(b))  # constant numerator: the division-free rule
        if op == 5:
            return ir.call("exp", ir.mul(ir.num(0.2), a))
        if op == 6:
            return ir.call("log", positive(a))
        if op == 7:
            return ir.call("sqrt", positive(a))
        if op == 8:
            return ir.power(positive(a), ir.num(rnd.choice([2.0, 3.0, 1.4, 0.24])))
        if op == 9:
            return ir.call("abs", ir.sub(a, ir.num(0.123)))
        return ir.cond(ir.cmp("lt", a, ir.num(0.4)), b, ir.neg(a))

    du = ir.diff(u_def, "y")
    checked = 0
    for _ in range(300):
        e = tree(rnd.randint(1, 4))
        de = ir.diff(e, "y", sym_diff=lambda name: du if name == "u" else ir.ZERO)
        y0, p0, h = rnd.uniform(-1.2, 1.2), rnd.uniform(0.3, 1.5), 1e-3

        def f(yv):
            return _eval_node(e, {"y": yv, "p": p0, "u": u_def})

        try:
            vals = [f(y0 + k * h) for k in (-2, -1, 1, 2)]
            got = _eval_node(de, {"y": y0, "p": p0, "u": u_def})
        except (OverflowError, ZeroDivisionError, ValueError):
            continue
        fd = (vals[0] - 8 * vals[1] + 8 * vals[2] - vals[3]) / (12 * h)
        # skip points where a Conditional / abs switches branch inside the stencil (the derivative jumps there)
        branches = {repr(_branch_signature(e, {"y": y0 + k * h, "p": p0, "u": u_def})) for k in (-2, -1, 0, 1, 2)}
        if len(branches) > 1 or not all(map(lambda v: abs(v) < 1e6, vals + [got])):
            continue
        assert abs(got - fd) <= 1e-6 * max(1.0, abs(fd), *map(abs, vals)), (e, got, fd)
        checked += 1
    assert checked >= 200


def _branch_signature(x, env, memo=None, out=None):
    """Which way every Conditional / abs of the expression goes at this point."""
    out = [] if out is None else out
    seen = set() if memo is None else memo

    def walk(n):
        if id(n) in seen:
            return
        seen.add(id(n))
        if n.kind == "sym" and hasattr(env.get(n.value), "kind"):
            walk(env[n.value])
        if n.kind == "cond":
            out.append(bool(_eval_node(n.args[0], env)))
        if n.kind == "call" and n.value == "abs":
            out.append(_eval_node(n.args[0], env) >= 0)
        for a in n.args:
            walk(a)

    walk(x)
    return out
