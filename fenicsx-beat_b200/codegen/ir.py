"""Expression IR for the cell-model code generator.

A deliberately tiny, hash-consed expression DAG.  It keeps the evaluation order
the ``.ode`` source wrote (no algebraic re-association), so the CUDA kernel and
any other back-end emitted from the same DAG perform the same floating-point
operations in the same order (up to FMA contraction by the compiler).

Node kinds
----------
num(value)            literal
sym(name)             state / parameter / intermediate / ``time`` / ``dt``
add sub mul div       binary arithmetic
neg                   unary minus
pow(a, b)             power (b usually a literal)
call(fn, args)        exp log sqrt floor abs
cond(c, a, b)         ternary (gotran ``Conditional``)
lt gt le ge eq ne     comparisons (boolean valued)
and or                boolean connectives (n-ary)
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable

_INTERN: dict[tuple, "Node"] = {}


@dataclass(frozen=True, eq=False)
class Node:
    kind: str
    args: tuple = ()
    value: float | str | None = None

    # identity hashing: nodes are interned, so `is` equality == structural equality
    def __hash__(self) -> int:  # pragma: no cover - trivial
        return id(self)

    def __eq__(self, other) -> bool:  # pragma: no cover - trivial
        return self is other

    # ---- operator sugar (used by the differentiator) -----------------------
    def __add__(self, o):
        return add(self, _lift(o))

    def __radd__(self, o):
        return add(_lift(o), self)

    def __sub__(self, o):
        return sub(self, _lift(o))

    def __rsub__(self, o):
        return sub(_lift(o), self)

    def __mul__(self, o):
        return mul(self, _lift(o))

    def __rmul__(self, o):
        return mul(_lift(o), self)

    def __truediv__(self, o):
        return div(self, _lift(o))

    def __rtruediv__(self, o):
        return div(_lift(o), self)

    def __neg__(self):
        return neg(self)

    def __repr__(self) -> str:
        if self.kind == "num":
            return repr(self.value)
        if self.kind == "sym":
            return str(self.value)
        if self.kind == "call":
            return f"{self.value}({', '.join(map(repr, self.args))})"
        return f"{self.kind}({', '.join(map(repr, self.args))})"


def _mk(kind: str, args: tuple = (), value=None) -> Node:
    key = (kind, tuple(id(a) for a in args), value if kind != "num" else repr(float(value)))
    n = _INTERN.get(key)
    if n is None:
        n = Node(kind, args, float(value) if kind == "num" else value)
        _INTERN[key] = n
    return n


def _lift(x) -> Node:
    return x if isinstance(x, Node) else num(x)


def num(v) -> Node:
    return _mk("num", (), float(v))


def sym(name: str) -> Node:
    return _mk("sym", (), name)


def is_num(n: Node, v: float | None = None) -> bool:
    return n.kind == "num" and (v is None or n.value == v)


ZERO = num(0.0)
ONE = num(1.0)


# ---- smart constructors (only identities that do not change rounding) -------
def add(a: Node, b: Node) -> Node:
    if is_num(a, 0.0):
        return b
    if is_num(b, 0.0):
        return a
    if is_num(a) and is_num(b):
        return num(a.value + b.value)
    if b.kind == "neg":
        return sub(a, b.args[0])
    return _mk("add", (a, b))


def sub(a: Node, b: Node) -> Node:
    if is_num(b, 0.0):
        return a
    if is_num(a, 0.0):
        return neg(b)
    if is_num(a) and is_num(b):
        return num(a.value - b.value)
    if b.kind == "neg":
        return add(a, b.args[0])
    return _mk("sub", (a, b))


def mul(a: Node, b: Node) -> Node:
    if is_num(a, 0.0) or is_num(b, 0.0):
        return ZERO
    if is_num(a, 1.0):
        return b
    if is_num(b, 1.0):
        return a
    if is_num(a, -1.0):
        return neg(b)
    if is_num(b, -1.0):
        return neg(a)
    if is_num(a) and is_num(b):
        return num(a.value * b.value)
    return _mk("mul", (a, b))


def div(a: Node, b: Node) -> Node:
    if is_num(a, 0.0):
        return ZERO
    if is_num(b, 1.0):
        return a
    if is_num(a) and is_num(b):
        return num(a.value / b.value)
    return _mk("div", (a, b))


def neg(a: Node) -> Node:
    if is_num(a):
        return num(-a.value)
    if a.kind == "neg":
        return a.args[0]
    return _mk("neg", (a,))


def power(a: Node, b: Node) -> Node:
    if is_num(b, 1.0):
        return a
    if is_num(b, 0.0):
        return ONE
    return _mk("pow", (a, b))


def call(fn: str, *args: Node) -> Node:
    return _mk("call", tuple(args), fn)


def cond(c: Node, a: Node, b: Node) -> Node:
    if a is b:
        return a
    return _mk("cond", (c, a, b))


def cmp(kind: str, a: Node, b: Node) -> Node:
    assert kind in ("lt", "gt", "le", "ge", "eq", "ne")
    return _mk(kind, (a, b))


def boolean(kind: str, args: Iterable[Node]) -> Node:
    assert kind in ("and", "or")
    args = tuple(args)
    if len(args) == 1:
        return args[0]
    return _mk(kind, args)


# ---- queries ---------------------------------------------------------------
def free_symbols(n: Node, _seen: dict | None = None) -> set[str]:
    memo: dict[int, set[str]] = {} if _seen is None else _seen

    def rec(x: Node) -> set[str]:
        r = memo.get(id(x))
        if r is not None:
            return r
        if x.kind == "sym":
            r = {x.value}
        else:
            r = set()
            for a in x.args:
                r |= rec(a)
        memo[id(x)] = r
        return r

    return rec(n)


def diff(n: Node, s: str, sym_diff=None) -> Node:
    """d n / d sym(s).  ``sym_diff(name)`` supplies d(name)/d(s) for every other symbol (chain rule through named
    intermediates); without it other symbols are independent of s."""
    memo: dict[int, Node] = {}

    def d(x: Node) -> Node:
        r = memo.get(id(x))
        if r is not None:
            return r
        k = x.kind
        if k == "num":
            r = ZERO
        elif k == "sym":
            r = ONE if x.value == s else (sym_diff(x.value) if sym_diff is not None else ZERO)
        elif k == "add":
            r = add(d(x.args[0]), d(x.args[1]))
        elif k == "sub":
            r = sub(d(x.args[0]), d(x.args[1]))
        elif k == "neg":
            r = neg(d(x.args[0]))
        elif k == "mul":
            a, b = x.args
            r = add(mul(d(a), b), mul(a, d(b)))
        elif k == "div":
            a, b = x.args
            da, db = d(a), d(b)
            if is_num(db, 0.0):
                r = div(da, b)
            elif is_num(da, 0.0) and is_num(a) and a.value != 0.0:
                # (c/b)' = -(c/b)^2 b' / c: no division at all (x = c/b is already there)
                r = mul(num(-1.0 / a.value), mul(mul(x, x), db))
            else:
                # (a/b)' = (a' - (a/b) b') / b: one division, and the quotient x is already there
                r = div(sub(da, mul(x, db)), b)
        elif k == "pow":
            a, b = x.args
            da, db = d(a), d(b)
            if not is_num(db, 0.0):
                raise NotImplementedError("state in exponent")
            if is_num(da, 0.0):
                r = ZERO
            else:
                r = mul(mul(b, power(a, sub(b, ONE))), da)
        elif k == "call":
            fn = x.value
            a = x.args[0]
            da = d(a)
            if is_num(da, 0.0):
                r = ZERO
            elif fn == "exp":
                r = mul(x, da)
            elif fn == "log":
                r = div(da, a)
            elif fn == "sqrt":
                r = div(da, mul(num(2.0), x))
            elif fn == "floor":
                r = ZERO
            elif fn == "abs":
                r = cond(cmp("ge", a, ZERO), da, neg(da))
            else:
                raise NotImplementedError(f"d/dx {fn}")
        elif k == "cond":
            c, a, b = x.args
            r = cond(c, d(a), d(b))
        else:
            raise NotImplementedError(f"diff of {k}")
        memo[id(x)] = r
        return r

    return d(n)


def rebuild(n: Node, leaf_map, memo: dict | None = None) -> Node:
    """Rebuild a DAG bottom-up through the smart constructors; ``leaf_map(node)`` may return a
    replacement for any node (checked top-down first) or None to recurse."""
    memo = {} if memo is None else memo

    def rec(x: Node) -> Node:
        r = memo.get(id(x))
        if r is not None:
            return r
        rep = leaf_map(x)
        if rep is not None:
            r = rep
        elif not x.args:
            r = x
        else:
            a = tuple(rec(c) for c in x.args)
            if all(p is q for p, q in zip(a, x.args)):
                r = x
            else:
                r = _mk(x.kind, a, x.value)
        memo[id(x)] = r
        return r

    return rec(n)


def numerator_is_nonzero_constant(n: Node) -> bool:
    """True for  c/expr , -(c/expr), -c/expr  with a non-zero literal c: the
    Rush-Larsen divisor can then never vanish and the |lin| > delta guard is
    dropped (what gotranx calls ``fraction_numerator_is_nonzero``)."""
    if n.kind == "neg":
        return numerator_is_nonzero_constant(n.args[0])
    if n.kind == "div":
        a = n.args[0]
        if a.kind == "neg":
            a = a.args[0]
        return is_num(a) and a.value != 0.0
    return False
