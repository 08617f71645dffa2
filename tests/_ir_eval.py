"""TEST INFRASTRUCTURE: a NumPy interpreter of the device code generator's scheduled program (codegen/program.py).

It evaluates exactly the DAG the CUDA emitter prints - uniform (parameter-only) part, node-dependent body, outputs - so
the generator's semantics (parsing, scheduling, hoisting, reciprocal pass, total-derivative Rush-Larsen) can be checked
against the independent SymPy-generated oracle models on the CPU, without a GPU.  Never imported by the product."""
import numpy as np


def evaluate(prog, states, t, dt, parameters):
    env = {"time": t, "dt": dt}
    for i, name in enumerate(prog.model.states):
        env[name] = states[i]
    for i, name in enumerate(prog.model.parameters):
        env[name] = parameters[i]
    memo = {}

    def ev(x):
        r = memo.get(id(x))
        if r is not None:
            return r
        k = x.kind
        if k == "num":
            r = x.value
        elif k == "sym":
            r = env[x.value]
        elif k == "add":
            r = ev(x.args[0]) + ev(x.args[1])
        elif k == "sub":
            r = ev(x.args[0]) - ev(x.args[1])
        elif k == "mul":
            r = ev(x.args[0]) * ev(x.args[1])
        elif k == "div":
            r = ev(x.args[0]) / ev(x.args[1])
        elif k == "neg":
            r = -ev(x.args[0])
        elif k == "pow":
            r = np.power(ev(x.args[0]), ev(x.args[1]))
        elif k == "call":
            a = ev(x.args[0])
            r = {"exp": np.exp, "log": np.log, "sqrt": np.sqrt, "floor": np.floor, "abs": np.abs}[x.value](a)
        elif k == "cond":
            r = np.where(ev(x.args[0]), ev(x.args[1]), ev(x.args[2]))
        elif k in ("lt", "gt", "le", "ge", "eq", "ne"):
            a, b = ev(x.args[0]), ev(x.args[1])
            r = {"lt": np.less, "gt": np.greater, "le": np.less_equal, "ge": np.greater_equal, "eq": np.equal, "ne": np.not_equal}[k](a, b)
        elif k == "and":
            r = np.logical_and.reduce([ev(a) for a in x.args])
        elif k == "or":
            r = np.logical_or.reduce([ev(a) for a in x.args])
        else:
            raise NotImplementedError(k)
        memo[id(x)] = r
        return r

    with np.errstate(all="ignore"):
        for name, e in prog.uniform:
            env[name] = ev(e)
        for name, e in prog.body:
            env[name] = ev(e)
        out = [np.broadcast_to(ev(e), states[0].shape) for e in prog.outputs]
    return np.array(out, dtype=np.float64)
