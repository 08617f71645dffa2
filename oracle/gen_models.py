"""ORACLE (test infrastructure, never a product path): cell-model step functions on the CPU.

Independent of the product's code generator on purpose: the ``.ode`` file is *executed* as Python in a
namespace of SymPy symbols (the gotran DSL is Python syntax), derivatives for the generalized
Rush-Larsen linearisation come from ``sympy.diff``, and code is printed by SymPy's own printers.  This
is the route gotranx itself takes (it is SymPy-based), restated from its published scheme because
gotranx cannot be installed here.  No reference test exercises the generated
TP06 / ToR-ORd step; the one reference output that does - the Niederer activation-time table,
demos/niederer_benchmark.py:315-325 - is reproduced within one dt (tests/test_oracle_niederer.py), which is what fixed
the linearisation rule below (total derivative).

Follows:
  * /root/reference/odes/tentusscher_panfilov_2006/tentusscher_panfilov_2006_epi_cell.ode:36-322
  * /root/reference/odes/torord/ToRORd_dynCl_endo.ode:1-633
  * FitzHugh-Nagumo right-hand side: /root/reference/README.md:58-89 (hand-written there)
  * call convention fun(states=, t=, parameters=, dt=) -> new states: src/beat/odesolver.py:67-79

Run here (needs /root/reference):   python oracle/gen_models.py
Writes oracle/models/<tag>.py (NumPy, vectorised over nodes) and oracle/c/<tag>.c (C, one node per
call, used by oracle/c/oracle_ode.c's OpenMP loop as the multi-core CPU baseline).
"""

from __future__ import annotations

import os
import sys

import sympy as sp
from sympy.printing.c import C99CodePrinter
from sympy.printing.numpy import NumPyPrinter

HERE = os.path.dirname(os.path.abspath(__file__))
ODES = os.environ.get("BEAT_ODES", "/root/reference/odes")

MODELS = {
    "fhn": None,
    "tp06": "tentusscher_panfilov_2006/tentusscher_panfilov_2006_epi_cell.ode",
    "torord": "torord/ToRORd_dynCl_endo.ode",
}

# README.md:58-129 of the reference, re-written in the .ode dialect (states s, v; 11 parameters in
# the README's order; stimulus window is open: t > start and t < start + duration, README.md:79-83).
FHN_TEXT = """
states("fhn", s=0.0, v=-85.0)
parameters("fhn", c_1=0.26, c_2=0.1, c_3=1.0, a=0.13, b=0.013, v_amp=125.0, v_rest=-85.0,
           v_peak=40.0, stim_amplitude=100.0, stim_duration=1.0, stim_start=0.0)
expressions("fhn")
i_app = Conditional(And(Gt(time, stim_start), Lt(time, stim_start + stim_duration)), stim_amplitude, 0)
ds_dt = b*(-c_3*s + (v - v_rest))
v_th = v_amp*a + v_rest
I = -s*(c_2/v_amp)*(v - v_rest) + (((c_1/v_amp**2)*(v - v_rest))*(v - v_th))*(-v + v_peak)
dv_dt = I + i_app
"""


class _Namespace(dict):
    """exec() locals: every unknown name is a Symbol; assignments are recorded, and the assigned name
    keeps resolving to its Symbol so intermediates stay opaque (as gotranx keeps them)."""

    def __init__(self):
        super().__init__()
        self.states: dict[str, float] = {}
        self.params: dict[str, float] = {}
        self.assigned: dict[str, sp.Expr] = {}
        self.builtin = {
            "states": self._states,
            "parameters": self._parameters,
            "expressions": lambda *a, **k: None,
            "ScalarParam": lambda v, **k: v,
            "exp": sp.exp,
            "log": sp.log,
            "sqrt": sp.sqrt,
            "floor": sp.floor,
            "Abs": sp.Abs,
            "Conditional": lambda c, a, b: sp.Piecewise((a, c), (b, True)),
            "And": lambda *a: sp.And(*a),
            "Or": lambda *a: sp.Or(*a),
            "Eq": sp.Eq,
            "Lt": sp.Lt,
            "Gt": sp.Gt,
            "Le": sp.Le,
            "Ge": sp.Ge,
        }

    def _states(self, *comp, **kw):
        for k, v in kw.items():
            self.states[k] = float(v)

    def _parameters(self, *comp, **kw):
        for k, v in kw.items():
            self.params[k] = float(v)

    def __getitem__(self, key):
        if key in self.builtin:
            return self.builtin[key]
        return sp.Symbol(key, real=True)

    def __setitem__(self, key, value):
        if key in self.assigned:
            raise ValueError(f"duplicate assignment {key}")
        self.assigned[key] = sp.sympify(value)


def load(tag: str):
    rel = MODELS[tag]
    text = FHN_TEXT if rel is None else open(os.path.join(ODES, rel)).read()
    ns = _Namespace()
    exec(compile(text, tag, "exec"), {"__builtins__": {}}, ns)
    return ns


def build(tag: str, scheme: str):
    """Returns (ordered assignments [(Symbol, expr)], outputs [expr per state], states, params)."""
    ns = load(tag)
    S = lambda n: sp.Symbol(n, real=True)  # noqa: E731
    dt = S("dt")
    defs = dict(ns.assigned)
    outputs = []
    extra = {}
    dmemo: dict = {}

    def total_diff(expr, y):
        """d expr / d y with the chain rule through every intermediate that depends on y; the derivative of an
        intermediate u becomes its own named assignment d<u>_d<y> (so nothing is expanded symbolically)."""
        out = sp.diff(expr, y)
        for u in sorted(expr.free_symbols, key=lambda q: q.name):
            if u.name not in ns.assigned or u == y:
                continue
            key = (u.name, y.name)
            if key not in dmemo:
                dmemo[key] = sp.Integer(0)
                du = total_diff(ns.assigned[u.name], y)
                if du != 0 and not du.is_number:
                    dn = f"d{u.name}_d{y.name}"
                    extra[dn] = du
                    du = S(dn)
                dmemo[key] = du
            if dmemo[key] != 0:
                out = out + sp.diff(expr, u) * dmemo[key]
        return out

    for s in ns.states:
        dname = f"d{s}_dt"
        f = defs[dname]
        y, fs = S(s), S(dname)
        if scheme == "forward_explicit_euler":
            outputs.append(y + dt * fs)
            continue
        # gotranx linearises with the TOTAL derivative (settled against the published Niederer table: see
        # tests/test_oracle_niederer.py); with intermediates held fixed V would fall back to forward Euler
        lin = total_diff(f, y)
        if lin.is_zero or lin == 0:
            outputs.append(y + dt * fs)
            continue
        lname = f"{dname}_linearized"
        extra[lname] = lin
        ls = S(lname)
        rl = fs * (sp.exp(ls * dt) - 1) / ls
        num, _den = sp.fraction(lin)
        if not (num.is_number and num != 0):
            rl = sp.Piecewise((rl, sp.Abs(ls) > 1e-8), (dt * fs, True))
        outputs.append(y + rl)
    defs.update(extra)
    # topological order
    order, mark = [], {}

    def visit(n):
        if mark.get(n) == 1:
            return
        assert mark.get(n) != 0, f"cycle at {n}"
        mark[n] = 0
        for d in sorted(defs[n].free_symbols, key=lambda q: q.name):
            if d.name in defs:
                visit(d.name)
        mark[n] = 1
        order.append(n)

    for n in defs:
        visit(n)
    assigns = [(S(n), defs[n]) for n in order]
    return assigns, outputs, ns.states, ns.params


class _Np(NumPyPrinter):
    def _print_Symbol(self, e):
        return e.name


def emit_numpy(tag: str) -> str:
    out = [
        f'"""ORACLE - GENERATED by oracle/gen_models.py: NumPy step functions for \'{tag}\'.',
        "Test infrastructure only (tests/, smoke(), bench.py cpu_baseline).  Pinned on the published Niederer table (within one dt).",
        '"""',
        "import numpy",
        "",
    ]
    for scheme in ("forward_explicit_euler", "generalized_rush_larsen"):
        assigns, outputs, states, params = build(tag, scheme)
        pr = _Np({"fully_qualified_modules": True})
        out.append(f"def {scheme}(states, t, dt, parameters):")
        out.append("    time = t")
        for i, s in enumerate(states):
            out.append(f"    {s} = states[{i}]")
        for i, p in enumerate(params):
            out.append(f"    {p} = parameters[{i}]")
        for lhs, rhs in assigns:
            out.append(f"    {lhs.name} = {pr.doprint(rhs)}")
        out.append("    values = numpy.zeros_like(states, dtype=numpy.float64)")
        for i, e in enumerate(outputs):
            out.append(f"    values[{i}] = {pr.doprint(e)}")
        out.append("    return values")
        out.append("")
        out.append("")
    assigns, outputs, states, params = build(tag, "forward_explicit_euler")
    out.append("state = {" + ", ".join(f"{s!r}: {i}" for i, s in enumerate(states)) + "}")
    out.append("parameter = {" + ", ".join(f"{p!r}: {i}" for i, p in enumerate(params)) + "}")
    out.append("_state_defaults = " + repr(list(states.values())))
    out.append("_parameter_defaults = " + repr(list(params.values())))
    out.append(
        '''

def state_index(name):
    return state[name]


def parameter_index(name):
    return parameter[name]


def init_state_values(**values):
    out = numpy.array(_state_defaults, dtype=numpy.float64)
    for k, v in values.items():
        out[state[k]] = v
    return out


def init_parameter_values(**values):
    out = numpy.array(_parameter_defaults, dtype=numpy.float64)
    for k, v in values.items():
        out[parameter[k]] = v
    return out
'''
    )
    return "\n".join(out)


class _C(C99CodePrinter):
    def _print_Symbol(self, e):
        return e.name


def emit_c(tag: str) -> str:
    out = [
        f"/* ORACLE - GENERATED by oracle/gen_models.py: scalar C step functions for '{tag}'.",
        "   Test infrastructure only; the OpenMP driver in oracle_ode.c loops these over nodes. */",
        "#include <math.h>",
        "",
    ]
    for scheme, short in (("forward_explicit_euler", "fe"), ("generalized_rush_larsen", "grl1")):
        assigns, outputs, states, params = build(tag, scheme)
        pr = _C()
        out.append(f"void {tag}_{short}(const double *y_in, double *y_out, double t, double dt, const double *prm) {{")
        out.append("  const double time = t; (void)time;")
        for i, s in enumerate(states):
            out.append(f"  const double {s} = y_in[{i}];")
        for i, p in enumerate(params):
            out.append(f"  const double {p} = prm[{i}]; (void){p};")
        for lhs, rhs in assigns:
            out.append(f"  const double {lhs.name} = {pr.doprint(rhs)};")
        for i, e in enumerate(outputs):
            out.append(f"  y_out[{i}] = {pr.doprint(e)};")
        out.append("}")
        out.append("")
    out.append(f"const int {tag}_num_states = {len(states)};")
    out.append(f"const int {tag}_num_params = {len(params)};")
    return "\n".join(out) + "\n"


def main():
    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "c"), exist_ok=True)
    for tag in MODELS:
        with open(os.path.join(HERE, "models", f"{tag}.py"), "w") as fh:
            fh.write(emit_numpy(tag))
        with open(os.path.join(HERE, "c", f"{tag}.c"), "w") as fh:
            fh.write(emit_c(tag))
        print("generated", tag)


if __name__ == "__main__":
    sys.exit(main())
