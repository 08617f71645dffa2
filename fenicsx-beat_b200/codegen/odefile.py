"""Front-end for gotran ``.ode`` cell-model files.

The DSL is a subset of Python syntax (``states(...)``, ``parameters(...)``,
``expressions(...)`` calls and ``name = expr`` assignments), so the file is parsed
with :mod:`ast` and lowered to the IR in :mod:`codegen.ir`.

Inputs followed: /root/reference/odes/tentusscher_panfilov_2006/
tentusscher_panfilov_2006_epi_cell.ode:36-322, /root/reference/odes/torord/
ToRORd_dynCl_endo.ode:1-633.  The model files themselves are *not* copied into
this repository; the generator reads them where they lie and the generated
CUDA/Python artefacts are committed (see generate.py).
"""

from __future__ import annotations

import ast
from dataclasses import dataclass, field

from . import ir


@dataclass
class OdeModel:
    name: str
    states: dict[str, float] = field(default_factory=dict)  # insertion order = index order
    parameters: dict[str, float] = field(default_factory=dict)
    intermediates: dict[str, ir.Node] = field(default_factory=dict)  # source order
    derivatives: dict[str, ir.Node] = field(default_factory=dict)  # state name -> rhs

    @property
    def state_names(self) -> list[str]:
        return list(self.states)

    @property
    def parameter_names(self) -> list[str]:
        return list(self.parameters)

    def state_index(self, name: str) -> int:
        return self.state_names.index(name)

    def parameter_index(self, name: str) -> int:
        return self.parameter_names.index(name)


_FUNCS = {"exp", "log", "sqrt", "floor"}
_CMPS = {"Lt": "lt", "Gt": "gt", "Le": "le", "Ge": "ge", "Eq": "eq", "Ne": "ne"}


def _literal(node: ast.expr) -> float:
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return float(node.value)
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
        return -_literal(node.operand)
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.UAdd):
        return _literal(node.operand)
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "ScalarParam":
        return _literal(node.args[0])
    raise ValueError(f"unsupported initial value: {ast.dump(node)}")


def _lower(node: ast.expr) -> ir.Node:
    if isinstance(node, ast.Constant):
        if isinstance(node.value, bool) or not isinstance(node.value, (int, float)):
            raise ValueError(f"unsupported literal {node.value!r}")
        return ir.num(node.value)
    if isinstance(node, ast.Name):
        return ir.sym(node.id)
    if isinstance(node, ast.UnaryOp):
        a = _lower(node.operand)
        if isinstance(node.op, ast.USub):
            return ir.neg(a)
        if isinstance(node.op, ast.UAdd):
            return a
        raise ValueError("unsupported unary operator")
    if isinstance(node, ast.BinOp):
        a, b = _lower(node.left), _lower(node.right)
        if isinstance(node.op, ast.Add):
            return ir.add(a, b)
        if isinstance(node.op, ast.Sub):
            return ir.sub(a, b)
        if isinstance(node.op, ast.Mult):
            return ir.mul(a, b)
        if isinstance(node.op, ast.Div):
            return ir.div(a, b)
        if isinstance(node.op, ast.Pow):
            return ir.power(a, b)
        raise ValueError(f"unsupported operator {node.op}")
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
        fn = node.func.id
        args = [_lower(a) for a in node.args]
        if fn in _FUNCS:
            (a,) = args
            return ir.call(fn, a)
        if fn == "Abs":
            (a,) = args
            return ir.call("abs", a)
        if fn == "Conditional":
            c, a, b = args
            return ir.cond(c, a, b)
        if fn in _CMPS:
            a, b = args
            return ir.cmp(_CMPS[fn], a, b)
        if fn in ("And", "Or"):
            return ir.boolean(fn.lower(), args)
        raise ValueError(f"unknown function {fn}")
    raise ValueError(f"unsupported syntax: {ast.dump(node)}")


def parse_ode(text: str, name: str) -> OdeModel:
    model = OdeModel(name=name)
    tree = ast.parse(text)
    assigns: dict[str, ir.Node] = {}
    for stmt in tree.body:
        if isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Call):
            fn = stmt.value.func.id  # type: ignore[attr-defined]
            if fn in ("states", "parameters"):
                target = model.states if fn == "states" else model.parameters
                for kw in stmt.value.keywords:
                    if kw.arg in model.states or kw.arg in model.parameters:
                        raise ValueError(f"duplicate declaration of {kw.arg}")
                    target[kw.arg] = _literal(kw.value)
            elif fn == "expressions":
                pass  # component headers carry no arithmetic
            else:
                raise ValueError(f"unknown statement {fn}(...)")
        elif isinstance(stmt, ast.Assign):
            (tgt,) = stmt.targets
            if tgt.id in assigns:  # type: ignore[attr-defined]
                raise ValueError(f"duplicate assignment to {tgt.id}")  # type: ignore[attr-defined]
            assigns[tgt.id] = _lower(stmt.value)  # type: ignore[attr-defined]
        elif isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant):
            continue  # stray docstring
        else:
            raise ValueError(f"unsupported statement: {ast.dump(stmt)[:80]}")

    # d<State>_dt lines define derivatives (state names may themselves end in "_")
    for lhs, rhs in assigns.items():
        if lhs.startswith("d") and lhs.endswith("_dt") and lhs[1:-3] in model.states:
            model.derivatives[lhs[1:-3]] = rhs
        else:
            model.intermediates[lhs] = rhs
    missing = [s for s in model.states if s not in model.derivatives]
    if missing:
        raise ValueError(f"states without derivative: {missing}")
    # every free symbol must resolve
    known = set(model.states) | set(model.parameters) | set(model.intermediates) | {"time"}
    for lhs, rhs in assigns.items():
        unk = ir.free_symbols(rhs) - known
        if unk:
            raise ValueError(f"{lhs}: unknown symbols {sorted(unk)}")
    return model


def load_ode(path: str, name: str | None = None) -> OdeModel:
    import os

    with open(path) as fh:
        text = fh.read()
    return parse_ode(text, name or os.path.splitext(os.path.basename(path))[0])


# FitzHugh-Nagumo as written in the reference README (README.md:58-129): the
# forward-Euler ``fun`` there is hand-written NumPy; this is the same right-hand
# side in the .ode dialect so that it goes through the same generator.
# parameter order = README.md:66-78 ; state order (s, v) = README.md:65, v_index=1.
FITZHUGH_NAGUMO_ODE = """
states("fhn", s=0.0, v=-85.0)
parameters("fhn", c_1=0.26, c_2=0.1, c_3=1.0, a=0.13, b=0.013, v_amp=125.0,
           v_rest=-85.0, v_peak=40.0, stim_amplitude=100.0, stim_duration=1.0,
           stim_start=0.0)
expressions("fhn")
i_app = Conditional(And(Gt(time, stim_start), Lt(time, stim_start + stim_duration)), stim_amplitude, 0)
ds_dt = b*(-c_3*s + (v - v_rest))
v_th = v_amp*a + v_rest
I = -s*(c_2/v_amp)*(v - v_rest) + (((c_1/v_amp**2)*(v - v_rest))*(v - v_th))*(-v + v_peak)
dv_dt = I + i_app
"""


# The two-state linear test model of the reference's own splitting tests (tests/test_monodomain_solver.py:25-30,
# tests/test_odesolver.py:20-49: v' = -s, s' = v, advanced with forward Euler), so that those tests can be run
# through the device path.  omega = 1 reproduces them exactly (multiplication by 1.0 is exact).
SIMPLE_OSCILLATOR_ODE = """
states("simple", v=0.0, s=0.0)
parameters("simple", omega=1.0)
expressions("simple")
dv_dt = -omega*s
ds_dt = omega*v
"""

BUILTIN_ODES = {"fhn": ("fitzhugh_nagumo", FITZHUGH_NAGUMO_ODE), "simple": ("simple_oscillator", SIMPLE_OSCILLATOR_ODE)}
