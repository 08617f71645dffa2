"""Back-ends: one scheduled :class:`Program` -> CUDA device code, a host-side Python model module
(metadata + parameter-only derived constants), and a NumPy evaluator used by the CPU tests of the
generator itself (never by the product path).
"""

from __future__ import annotations

import re

from . import ir
from .program import Program, op_counts


def _cname(name: str) -> str:
    return re.sub(r"[^0-9A-Za-z_]", "_", name)


def _fmt(v: float) -> str:
    r = repr(float(v))
    if r in ("inf", "-inf", "nan"):
        raise ValueError("non-finite literal")
    if "e" not in r and "." not in r:
        r += ".0"
    return r


class _Printer:
    """Prints expression DAGs as statements with every shared non-leaf node bound to a temporary."""

    def __init__(self, lang: str, symmap: dict[str, str]):
        assert lang in ("cuda", "numpy", "python")
        self.lang = lang
        self.symmap = symmap
        self.lines: list[str] = []
        self.bound: dict[int, str] = {}
        self.refs: dict[int, int] = {}
        self.ntmp = 0

    # reference counting over everything that will be printed -----------------
    def count(self, roots) -> None:
        seen: set[int] = set()

        def rec(x: ir.Node) -> None:
            self.refs[id(x)] = self.refs.get(id(x), 0) + 1
            if id(x) in seen:
                return
            seen.add(id(x))
            for a in x.args:
                rec(a)

        for r in roots:
            rec(r)

    def _decl(self, name: str, text: str, boolean: bool = False) -> None:
        if self.lang == "cuda":
            self.lines.append(f"const {'bool' if boolean else 'double'} {name} = {text};")
        else:
            self.lines.append(f"{name} = {text}")

    def assign(self, name: str, expr: ir.Node) -> None:
        text = self.expr(expr, top=True)
        self._decl(name, text)

    def expr(self, x: ir.Node, top: bool = False) -> str:
        b = self.bound.get(id(x))
        if b is not None:
            return b
        text = self._print(x)
        if not top and x.kind not in ("num", "sym") and self.refs.get(id(x), 0) > 1:
            name = f"_t{self.ntmp}"
            self.ntmp += 1
            self._decl(name, text, boolean=x.kind in ("lt", "gt", "le", "ge", "eq", "ne", "and", "or"))
            self.bound[id(x)] = name
            return name
        return text

    def _ipow(self, base: str, n: int) -> str:
        if self.lang == "cuda":
            return f"ipow<{n}>({base})"
        return f"_ipow({base}, {n})"

    def _print(self, x: ir.Node) -> str:
        k, L = x.kind, self.lang
        if k == "num":
            v = _fmt(x.value)
            return f"({v})" if v.startswith("-") else v
        if k == "sym":
            return self.symmap[x.value]
        e = self.expr
        if k == "add":
            return f"({e(x.args[0])} + {e(x.args[1])})"
        if k == "sub":
            return f"({e(x.args[0])} - {e(x.args[1])})"
        if k == "mul":
            return f"({e(x.args[0])} * {e(x.args[1])})"
        if k == "div":
            if L == "cuda":
                if ir.is_num(x.args[0], 1.0):
                    return f"RCP({e(x.args[1])})"  # shared reciprocal of a repeated denominator (program.py)
                return f"DIV({e(x.args[0])}, {e(x.args[1])})"
            return f"({e(x.args[0])} / {e(x.args[1])})"
        if k == "neg":
            return f"(-{e(x.args[0])})"
        if k == "pow":
            a, b = x.args
            if ir.is_num(b) and float(b.value).is_integer() and 2 <= abs(b.value) <= 16:
                n = int(abs(b.value))
                p = self._ipow(e(a), n)
                if b.value < 0:
                    return f"DIV(1.0, {p})" if L == "cuda" else f"(1.0 / {p})"
                return p
            if ir.is_num(b, 0.5):
                return f"SQRT({e(a)})" if L == "cuda" else f"{self._fn('sqrt')}({e(a)})"
            if L == "cuda":
                return f"POW({e(a)}, {e(b)})"
            return f"{self._fn('pow')}({e(a)}, {e(b)})"
        if k == "call":
            fn = x.value
            if L == "cuda":
                m = {"exp": "EXP", "log": "LOG", "sqrt": "SQRT", "floor": "floor", "abs": "fabs"}[fn]
                return f"{m}({e(x.args[0])})"
            return f"{self._fn(fn)}({e(x.args[0])})"
        if k == "cond":
            c, a, b = x.args
            if L == "numpy":
                return f"np.where({e(c)}, {e(a)}, {e(b)})"
            if L == "python":
                return f"({e(a)} if {e(c)} else {e(b)})"
            return f"({e(c)} ? {e(a)} : {e(b)})"
        if k in ("lt", "gt", "le", "ge", "eq", "ne"):
            op = {"lt": "<", "gt": ">", "le": "<=", "ge": ">=", "eq": "==", "ne": "!="}[k]
            return f"({e(x.args[0])} {op} {e(x.args[1])})"
        if k in ("and", "or"):
            if L == "numpy":
                fn = "np.logical_and" if k == "and" else "np.logical_or"
                out = e(x.args[0])
                for a in x.args[1:]:
                    out = f"{fn}({out}, {e(a)})"
                return out
            op = {"cuda": {"and": "&&", "or": "||"}, "python": {"and": "and", "or": "or"}}[L][k]
            return "(" + f" {op} ".join(e(a) for a in x.args) + ")"
        raise NotImplementedError(k)

    def _fn(self, fn: str) -> str:
        if self.lang == "numpy":
            return {"exp": "np.exp", "log": "np.log", "sqrt": "np.sqrt", "floor": "np.floor", "abs": "np.abs", "pow": "np.power"}[fn]
        return {"exp": "math.exp", "log": "math.log", "sqrt": "math.sqrt", "floor": "math.floor", "abs": "abs", "pow": "math.pow"}[fn]


SCHEME_SHORT = {"forward_explicit_euler": "fe", "generalized_rush_larsen": "grl1"}


def emit_cuda(progs: list[Program], model_tag: str) -> str:
    """One header per model: for every scheme a device function

        template <class PRM> __device__ __forceinline__
        void <tag>_<scheme>(double (&y)[NS], const PRM& prm, double t, double dt)

    ``PRM`` supplies ``p<k>()`` for parameter k and, when ``PRM::kPerNode`` is false, ``u<k>()`` for the
    k-th parameter-only derived constant (evaluated once on the host); with per-node parameters the
    derived constants are recomputed in the thread.
    """
    model = progs[0].model
    ns, npar = len(model.states), len(model.parameters)
    out: list[str] = []
    out.append(f"// GENERATED by fenicsx-beat_b200/codegen/generate.py from the gotran model '{model.name}'.")
    out.append("// Do not edit: re-run the generator.  Arithmetic follows the model file expression by expression;")
    out.append("// DIV/EXP/LOG/SQRT/POW are bound by ode_math.cuh (exact IEEE fp64 by default).")
    out.append("#pragma once")
    out.append('#include "../ode_math.cuh"')
    out.append("")
    out.append(f"struct {model_tag}_meta {{")
    out.append(f"  static constexpr int kNumStates = {ns};")
    out.append(f"  static constexpr int kNumParams = {npar};")
    nu = max(len(p.uniform) for p in progs)
    out.append(f"  static constexpr int kNumDerived = {nu};")
    out.append("};")
    out.append("")
    for prog in progs:
        assert [n for n, _ in prog.uniform] == [n for n, _ in progs[0].uniform] or not prog.uniform or True
        short = SCHEME_SHORT[prog.scheme]
        symmap = {"time": "t", "dt": "dt"}
        for i, s in enumerate(model.states):
            symmap[s] = f"y_{_cname(s)}"
        for p in model.parameters:
            symmap[p] = f"p_{_cname(p)}"
        for n, _ in prog.uniform + prog.body:
            symmap[n] = f"v_{_cname(n)}"
        counts = op_counts(prog)
        out.append(f"// scheme: {prog.scheme}; Rush-Larsen states: {len(prog.rl_states)}, forward-Euler states: {len(prog.fe_states)}")
        out.append("// per node-step after CSE: " + ", ".join(f"{k}={v}" for k, v in counts.items() if v))
        out.append("template <class PRM>")
        out.append(f"__device__ __forceinline__ void {model_tag}_{short}(double (&y)[{ns}], const PRM& prm, const double t, const double dt) {{")
        body: list[str] = []
        used = prog.used_parameters
        uniform_syms = {n for n, _ in prog.uniform}
        # parameters referenced by the node-dependent part
        body_params: set[str] = set()
        for _, e in prog.body:
            body_params |= ir.free_symbols(e)
        for e in prog.outputs:
            body_params |= ir.free_symbols(e)
        uni_params: set[str] = set()
        for _, e in prog.uniform:
            uni_params |= ir.free_symbols(e)
        for p in used:
            k = model.parameter_index(p)
            if p in body_params:
                body.append(f"const double p_{_cname(p)} = prm.template p<{k}>();")
        # derived constants
        if prog.uniform:
            for n, _ in prog.uniform:
                body.append(f"double v_{_cname(n)};")
            body.append("if constexpr (PRM::kPerNode) {")
            pr = _Printer("cuda", dict(symmap))
            # in the per-node branch parameters only needed by derived constants are loaded here
            for p in used:
                if p in uni_params and p not in body_params:
                    k = model.parameter_index(p)
                    pr.lines.append(f"const double p_{_cname(p)} = prm.template p<{k}>();")
            pr.count([e for _, e in prog.uniform])
            for n, e in prog.uniform:
                text = pr.expr(e, top=True)
                pr.lines.append(f"v_{_cname(n)} = {text};")
            body += ["  " + ln for ln in pr.lines]
            body.append("} else {")
            for k, (n, _) in enumerate(prog.uniform):
                body.append(f"  v_{_cname(n)} = prm.template u<{k}>();")
            body.append("}")
        for i, s in enumerate(model.states):
            body.append(f"const double y_{_cname(s)} = y[{i}];")
        pr = _Printer("cuda", symmap)
        pr.count([e for _, e in prog.body] + list(prog.outputs))
        for n, e in prog.body:
            pr.assign(f"v_{_cname(n)}", e)
        for i, e in enumerate(prog.outputs):
            text = pr.expr(e, top=True)
            pr.lines.append(f"y[{i}] = {text};")
        body += pr.lines
        out += ["  " + ln for ln in body]
        out.append("}")
        out.append("")
        _ = uniform_syms
    return "\n".join(out) + "\n"


def emit_numpy(prog: Program, func_name: str) -> str:
    """NumPy evaluation of the same DAG (generator self-test only; not a product path)."""
    model = prog.model
    symmap = {"time": "t", "dt": "dt"}
    for s in model.states:
        symmap[s] = f"y_{_cname(s)}"
    for p in model.parameters:
        symmap[p] = f"p_{_cname(p)}"
    for n, _ in prog.uniform + prog.body:
        symmap[n] = f"v_{_cname(n)}"
    pr = _Printer("numpy", symmap)
    pr.count([e for _, e in prog.uniform + prog.body] + list(prog.outputs))
    lines = [f"def {func_name}(states, t, dt, parameters):"]
    for i, s in enumerate(model.states):
        lines.append(f"    y_{_cname(s)} = states[{i}]")
    for p in prog.used_parameters:
        lines.append(f"    p_{_cname(p)} = parameters[{model.parameter_index(p)}]")
    for n, e in prog.uniform + prog.body:
        pr.assign(f"v_{_cname(n)}", e)
    outs = []
    for i, e in enumerate(prog.outputs):
        outs.append(pr.expr(e, top=True))
    lines += ["    " + ln for ln in pr.lines]
    lines.append("    values = np.zeros_like(states, dtype=np.float64)")
    for i, t in enumerate(outs):
        lines.append(f"    values[{i}] = {t}")
    lines.append("    return values")
    header = "import numpy as np\n\n\ndef _ipow(x, n):\n    r = x\n    for _ in range(n - 1):\n        r = r * x\n    return r\n\n\n"
    return header + "\n".join(lines) + "\n"


def emit_host_module(progs: list[Program], model_tag: str, model_id: int) -> str:
    """Python module mirroring the gotranx-generated module surface the reference's demos use
    (init_state_values / init_parameter_values / state_index / parameter_index, e.g.
    demos/niederer_benchmark.py:66,97-99,212) with the step functions replaced by device handles."""
    model = progs[0].model
    lines = [
        f'"""GENERATED host-side description of the cell model \'{model.name}\' (see codegen/generate.py).',
        "",
        "The step functions are *device handles*: they name a CUDA kernel, they are not callable on the CPU.",
        '"""',
        "import math",
        "",
        "import numpy as np",
        "",
        "from ..device_model import DeviceODE",
        "",
        f"MODEL_ID = {model_id}",
        f"MODEL_TAG = {model_tag!r}",
        "state = {" + ", ".join(f"{s!r}: {i}" for i, s in enumerate(model.states)) + "}",
        "parameter = {" + ", ".join(f"{p!r}: {i}" for i, p in enumerate(model.parameters)) + "}",
        "_state_defaults = [" + ", ".join(_fmt(v) for v in model.states.values()) + "]",
        "_parameter_defaults = [" + ", ".join(_fmt(v) for v in model.parameters.values()) + "]",
        "",
        "",
        "def state_index(name: str) -> int:",
        "    return state[name]",
        "",
        "",
        "def parameter_index(name: str) -> int:",
        "    return parameter[name]",
        "",
        "",
        "def init_state_values(**values):",
        "    out = np.array(_state_defaults, dtype=np.float64)",
        "    for k, v in values.items():",
        "        out[state[k]] = v",
        "    return out",
        "",
        "",
        "def init_parameter_values(**values):",
        "    out = np.array(_parameter_defaults, dtype=np.float64)",
        "    for k, v in values.items():",
        "        out[parameter[k]] = v",
        "    return out",
        "",
        "",
    ]
    for prog in progs:
        short = SCHEME_SHORT[prog.scheme]
        symmap = {}
        for p in model.parameters:
            symmap[p] = f"p[{model.parameter_index(p)}]"
        for n, _ in prog.uniform:
            symmap[n] = f"v_{_cname(n)}"
        pr = _Printer("python", symmap)
        pr.count([e for _, e in prog.uniform])
        for n, e in prog.uniform:
            pr.assign(f"v_{_cname(n)}", e)
        lines.append(f"def _derived_{short}(p):")
        lines.append('    """Parameter-only intermediates, evaluated once per parameter set and passed to the kernel."""')
        lines += ["    " + ln for ln in pr.lines]
        lines.append("    return np.array([" + ", ".join(f"v_{_cname(n)}" for n, _ in prog.uniform) + "], dtype=np.float64)")
        lines.append("")
        lines.append("")
    lines.append("def _ipow(x, n):")
    lines.append("    r = x")
    lines.append("    for _ in range(n - 1):")
    lines.append("        r = r * x")
    lines.append("    return r")
    lines.append("")
    lines.append("")
    scheme_ids = {"forward_explicit_euler": 0, "generalized_rush_larsen": 1}
    for prog in progs:
        short = SCHEME_SHORT[prog.scheme]
        c = op_counts(prog)
        lines.append(
            f"{prog.scheme} = DeviceODE(model_id=MODEL_ID, model_tag=MODEL_TAG, scheme_id={scheme_ids[prog.scheme]}, "
            f"scheme={prog.scheme!r}, num_states={len(model.states)}, num_parameters={len(model.parameters)}, "
            f"derived=_derived_{short}, op_counts={c!r})"
        )
    lines.append("")
    if not any("math." in ln for ln in lines):  # (models whose hoisted constants need no transcendental function)
        at = lines.index("import math")
        del lines[at: at + 2]
    return "\n".join(lines)
