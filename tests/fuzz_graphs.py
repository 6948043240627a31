"""Seeded random op-graph programs over the reference's Tensor API (test infrastructure, not product).

A program is data - leaf arrays, a list of (op, input nodes, arguments) and weighted terminal nodes - generated from a
seed with a NumPy float32 shadow evaluation that keeps every intermediate finite, away from the singularities of
log / sqrt / division / fractional powers and free of (near-)ties under max / maximum / minimum.  `run_program` executes a program with ANY Tensor
class that has the reference's surface (src/tensor/tensor.py): the live reference (build container), this package's
NumPy path, or this package's CUDA path.  The loss is sum_k (node_k * w_k).sum(); `run_program` returns the terminal
values, the loss and every leaf gradient as host arrays.

Used by tests/test_fuzz_cpu.py (package CPU path == live reference, fixtures == reference) and tests/test_fuzz_gpu.py
(CUDA path == CPU path == reference fixtures).  The ops and argument forms follow the reference's signatures:
overload.py:52-164 (+ - * / @ and their scalar / reflected forms, get_item), math.py (pow, log, exp, tanh, sqrt),
reduce.py (sum with int / tuple axes, mean, max), base.py (view / reshape, transpose, squeeze, unsqueeze, broadcast_to),
elementwise.py (abs, maximum, minimum) and - in the `select=True` variant - tensor.py:467-503 (comparisons with where,
clip, threshold, masked_fill, sign).
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np

DIMS = [1, 2, 3, 4, 5, 7, 8, 16, 33]
MAX_ELEMS = 1 << 16
MAX_ABS = 1.0e3


class Program:
    def __init__(self, seed: int):
        self.seed = seed
        self.leaves: List[np.ndarray] = []
        self.ops: List[Tuple[str, Tuple[int, ...], Dict[str, Any]]] = []
        self.terminals: List[Tuple[int, np.ndarray]] = []

    def describe(self) -> str:
        lines = [f"program seed={self.seed}"]
        for i, a in enumerate(self.leaves):
            lines.append(f"  n{i} = leaf{a.shape}")
        for k, (op, ins, args) in enumerate(self.ops):
            shown = {key: (f"array{np.shape(v)}" if isinstance(v, np.ndarray) else v) for key, v in args.items()}
            lines.append(f"  n{len(self.leaves) + k} = {op}({', '.join('n%d' % i for i in ins)}, {shown})")
        lines.append("  terminals: " + ", ".join("n%d" % k for k, _ in self.terminals))
        return "\n".join(lines)


# --------------------------------------------------------------------------------------------------- shadow semantics
def _shadow(op: str, xs: List[np.ndarray], args: Dict[str, Any]) -> np.ndarray:
    a = xs[0]
    if op in ("add", "sub", "mul", "div"):
        b = xs[1]
        return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b}[op]
    if op == "sadd": return a + np.float32(args["c"])
    if op == "sradd": return np.float32(args["c"]) + a
    if op == "ssub": return a - np.float32(args["c"])
    if op == "srsub": return np.float32(args["c"]) - a
    if op == "smul": return a * np.float32(args["c"])
    if op == "srmul": return np.float32(args["c"]) * a
    if op == "sdiv": return a / np.float32(args["c"])
    if op == "srdiv": return np.float32(args["c"]) / a
    if op == "neg": return -a
    if op == "pow": return a ** np.float32(args["p"])
    if op == "log": return np.log(a)
    if op == "exp": return np.exp(a)
    if op == "tanh": return np.tanh(a)
    if op == "sqrt": return np.sqrt(a)
    if op == "abs": return np.abs(a)
    if op == "maximum": return np.maximum(a, xs[1])
    if op == "minimum": return np.minimum(a, xs[1])
    if op == "sum": return np.asarray(a.sum(axis=args["axis"], keepdims=args["keepdims"]))
    if op == "mean": return np.asarray(a.sum() * np.float32(1.0 / a.size))
    if op == "max": return np.asarray(a.max(axis=args["axis"], keepdims=args["keepdims"]))
    if op == "transpose": return np.transpose(a, args["axes"])
    if op == "T": return a.T
    if op in ("reshape", "view"): return a.reshape(args["shape"])
    if op == "squeeze": return np.squeeze(a, axis=args["dim"])
    if op == "unsqueeze": return np.expand_dims(a, axis=args["dim"])
    if op == "broadcast_to": return np.broadcast_to(a, args["shape"])
    if op == "getitem": return np.asarray(a[args["index"]])
    if op == "matmul": return np.asarray(a @ xs[1])
    if op == "clip": return np.clip(a, np.float32(args["lo"]), np.float32(args["hi"]))
    if op == "threshold": return np.where(a > np.float32(args["thr"]), a, np.float32(args["val"]))
    if op == "where_gt": return np.where(a > np.float32(args["c"]), a, xs[1])
    if op == "masked_fill": return np.where(xs[1] > np.float32(args["c"]), np.float32(args["val"]), a)
    if op == "sign": return np.sign(a)
    raise KeyError(op)


def _apply(op: str, ts: list, args: Dict[str, Any], Tensor):
    a = ts[0]
    if op == "add": return a + ts[1]
    if op == "sub": return a - ts[1]
    if op == "mul": return a * ts[1]
    if op == "div": return a / ts[1]
    if op == "sadd": return a + args["c"]
    if op == "sradd": return args["c"] + a
    if op == "ssub": return a - args["c"]
    if op == "srsub": return args["c"] - a
    if op == "smul": return a * args["c"]
    if op == "srmul": return args["c"] * a
    if op == "sdiv": return a / args["c"]
    if op == "srdiv": return args["c"] / a
    if op == "neg": return -a
    if op == "pow": return a ** args["p"]
    if op == "log": return a.log()
    if op == "exp": return a.exp()
    if op == "tanh": return a.tanh()
    if op == "sqrt": return a.sqrt()
    if op == "abs": return a.abs()
    if op == "maximum": return Tensor.maximum(a, ts[1])
    if op == "minimum": return Tensor.minimum(a, ts[1])
    if op == "sum": return a.sum(axis=args["axis"], keepdims=args["keepdims"])
    if op == "mean": return a.mean()
    if op == "max": return a.max(axis=args["axis"], keepdims=args["keepdims"])
    if op == "transpose": return a.transpose(args["axes"])
    if op == "T": return a.T
    if op == "reshape": return a.reshape(args["shape"])
    if op == "view": return a.view(args["shape"])
    if op == "squeeze": return a.squeeze(args["dim"])
    if op == "unsqueeze": return a.unsqueeze(args["dim"])
    if op == "broadcast_to": return a.broadcast_to(args["shape"])
    if op == "getitem": return a[args["index"]]
    if op == "matmul": return a @ ts[1]
    if op == "clip": return a.clip(args["lo"], args["hi"])
    if op == "threshold": return a.threshold(args["thr"], args["val"])
    if op == "where_gt": return Tensor.where(a > args["c"], a, ts[1])
    if op == "masked_fill": return a.masked_fill(ts[1] > args["c"], args["val"])
    if op == "sign": return a.sign()
    raise KeyError(op)


# --------------------------------------------------------------------------------------------------- generation
_OPS = [("add", 6), ("sub", 4), ("mul", 7), ("div", 4), ("sadd", 2), ("sradd", 1), ("ssub", 1), ("srsub", 2), ("smul", 2),
        ("srmul", 1), ("sdiv", 1), ("srdiv", 2), ("neg", 2), ("pow", 5), ("log", 3), ("exp", 3), ("tanh", 4), ("sqrt", 2),
        ("abs", 1), ("maximum", 1), ("minimum", 1), ("sum", 6), ("mean", 1), ("max", 2), ("transpose", 3), ("T", 2),
        ("reshape", 3), ("view", 1), ("squeeze", 2), ("unsqueeze", 2), ("broadcast_to", 2), ("getitem", 5), ("matmul", 5)]
_NAMES = [n for n, _ in _OPS]
_WEIGHTS = np.array([w for _, w in _OPS], dtype=np.float64)
_WEIGHTS /= _WEIGHTS.sum()
_BINARY = ("add", "sub", "mul", "div", "maximum", "minimum")
# the select family (tensor.py:467-503 of the reference: comparisons, where, clip, threshold, masked_fill, sign) - a separate
# variant (`select=True`) so that the programs of the base variant, and their fixtures, do not change
_SELECT_OPS = [("clip", 5), ("threshold", 4), ("where_gt", 5), ("masked_fill", 4), ("sign", 2)]
_NAMES_SEL = _NAMES + [n for n, _ in _SELECT_OPS]
_WEIGHTS_SEL = np.array([w for _, w in _OPS] + [w for _, w in _SELECT_OPS], dtype=np.float64)
_WEIGHTS_SEL /= _WEIGHTS_SEL.sum()


def _rand_shape(rng, dims) -> Tuple[int, ...]:
    rank = int(rng.choice([1, 2, 2, 2, 3, 3, 4]))
    return tuple(int(rng.choice(dims)) for _ in range(rank))


def _variant_shape(rng, base: Tuple[int, ...]) -> Tuple[int, ...]:
    """A shape that broadcasts with `base`: same, 1-dims, leading dims dropped, or a scalar."""
    kind = rng.integers(0, 6)
    if kind == 0:
        return ()
    s = list(base)
    if kind in (1, 2):
        for i in range(len(s)):
            if rng.random() < 0.5:
                s[i] = 1
    if kind in (2, 3) and len(s) > 1:
        s = s[int(rng.integers(1, len(s))):]
    return tuple(s)


def _pick_args(rng, op: str, a: np.ndarray, vals: List[np.ndarray], max_elems: int = MAX_ELEMS):
    """Returns (input indices beyond the first, args) or None when `op` does not apply to `a`."""
    nd = a.ndim
    if op in _BINARY:
        cands = []
        for j, b in enumerate(vals):
            try:
                shp = np.broadcast_shapes(a.shape, b.shape)
            except ValueError:
                continue
            if int(np.prod(shp)) <= max_elems:
                cands.append(j)
        if not cands:
            return None
        if op in ("maximum", "minimum"):      # the reference's maximum / minimum do not un-broadcast their gradients
            cands = [j for j in cands if vals[j].shape == a.shape]
            if not cands:
                return None
        j = int(rng.choice(cands))
        if op in ("sub", "div") and vals[j] is a:
            return None                       # x - x, x / x: the gradient is pure cancellation noise
        if op == "div" and np.abs(vals[j]).min(initial=1.0) < 0.1:
            return None
        if op in ("maximum", "minimum") and np.abs(a - vals[j]).min(initial=1.0) <= 1e-4 * (1.0 + np.abs(a).max(initial=0.0)):
            return None                       # a (near-)tie: which side wins is decided by the last ulp of the operands
        return (j,), {}
    if op in ("sadd", "sradd", "ssub", "srsub", "smul", "srmul", "sdiv", "srdiv"):
        if op == "srdiv" and np.abs(a).min(initial=1.0) < 0.1:
            return None
        return (), {"c": float(np.round(rng.uniform(0.5, 2.0), 3))}
    if op in ("neg", "tanh", "T", "mean"):
        return (), {}
    if op == "abs":
        return ((), {}) if np.abs(a).min(initial=1.0) > 1e-3 else None
    if op == "pow":
        p = [2, 3, -1, 0.5, 1, 2.5, 0][int(rng.integers(0, 7))]
        if p in (0.5, 2.5) and a.min(initial=1.0) < 0.05:
            return None
        if p in (-1, 0) and np.abs(a).min(initial=1.0) < 0.1:
            return None
        return (), {"p": p}
    if op in ("log", "sqrt"):
        return ((), {}) if a.min(initial=1.0) > 0.05 else None
    if op == "exp":
        return ((), {}) if a.max(initial=0.0) < 5.0 else None
    if op == "sum":
        keep = bool(rng.integers(0, 2))
        if nd == 0 or rng.random() < 0.3:
            return (), {"axis": None, "keepdims": keep}
        if nd >= 2 and rng.random() < 0.25:
            ax = tuple(sorted(int(v) for v in rng.choice(nd, size=2, replace=False)))
            return (), {"axis": ax, "keepdims": keep}
        return (), {"axis": int(rng.integers(-nd, nd)), "keepdims": keep}
    if op == "max":
        if a.size == 0:
            return None
        if nd == 0 or rng.random() < 0.4:
            axis, keep = None, False
        else:
            axis, keep = int(rng.integers(-nd, nd)), True      # the reference's mask needs keepdims along an axis
        m = a.max(axis=axis, keepdims=True)
        if np.any((a == m).sum(axis=axis, keepdims=True) != 1):
            return None
        runner_up = np.where(a == m, -np.inf, a).max(axis=axis, keepdims=True)
        if np.any(m - runner_up <= 1e-4 * (1.0 + np.abs(m))):
            return None                       # the winner must not depend on the last ulp of the producer (tanh, exp ...)
        return (), {"axis": axis, "keepdims": keep}
    if op == "transpose":
        if nd < 2 or rng.random() < 0.3:
            return (), {"axes": None}
        return (), {"axes": tuple(int(v) for v in rng.permutation(nd))}
    if op in ("reshape", "view"):
        if a.size == 0:
            return None
        if op == "view" and not a.flags["C_CONTIGUOUS"]:
            return None
        kind = int(rng.integers(0, 4))
        shape = list(a.shape)
        if kind == 0 or nd == 0:
            shape = [a.size]
        elif kind == 1 and nd >= 2:
            i = int(rng.integers(0, nd - 1))
            shape = shape[:i] + [shape[i] * shape[i + 1]] + shape[i + 2:]
        elif kind == 2:
            i = int(rng.integers(0, nd))
            for f in (2, 3, 4, 11):
                if shape[i] % f == 0 and shape[i] > f:
                    shape = shape[:i] + [f, shape[i] // f] + shape[i + 1:]
                    break
        else:
            shape.insert(int(rng.integers(0, nd + 1)), 1)
        return (), {"shape": tuple(int(v) for v in shape)}
    if op == "squeeze":
        ones = [i for i, d in enumerate(a.shape) if d == 1]
        if not ones:
            return None
        return (), {"dim": int(rng.choice(ones))}
    if op == "unsqueeze":
        return (), {"dim": int(rng.integers(0, nd + 1))}
    if op == "broadcast_to":
        shape = [int(rng.choice([2, 3, 5])) if (d == 1 and rng.random() < 0.7) else d for d in a.shape]
        if rng.random() < 0.5 and nd < 4:
            shape = [int(rng.choice([2, 3]))] + shape
        if tuple(shape) == a.shape or int(np.prod(shape)) > max_elems:
            return None
        return (), {"shape": tuple(shape)}
    if op == "getitem":
        if nd == 0 or a.size == 0:
            return None
        if rng.random() < 0.3:                                   # integer-array rows (unique: get_item's backward assigns)
            n = a.shape[0]
            k = int(rng.integers(1, n + 1))
            return (), {"index": rng.permutation(n)[:k].astype(np.int64)}
        idx = []
        for d in a.shape[: int(rng.integers(1, nd + 1))]:
            r = rng.random()
            if r < 0.3:
                idx.append(int(rng.integers(-d, d)))
            elif r < 0.45:
                idx.append(slice(None))
            else:
                lo = int(rng.integers(0, d))
                hi = int(rng.integers(lo + 1, d + 1))
                idx.append(slice(lo, hi, int(rng.choice([1, 1, 2]))))
        return (), {"index": tuple(idx) if len(idx) > 1 or rng.random() < 0.5 else idx[0]}
    if op in ("clip", "threshold", "where_gt", "masked_fill", "sign"):
        if a.size == 0:
            return None

        def clear_of(x, c):                   # no element within reach of the comparison's flip
            return np.abs(x - np.float32(c)).min(initial=1.0) > 1e-4 * (1.0 + abs(c))

        lo_v, hi_v = float(a.min()), float(a.max())
        if op == "sign":
            return ((), {}) if clear_of(a, 0.0) else None
        if op == "clip":
            lo = float(np.round(rng.uniform(lo_v - 0.2, 0.5 * (lo_v + hi_v)), 3))
            hi = float(np.round(rng.uniform(0.5 * (lo_v + hi_v), hi_v + 0.2), 3))
            return ((), {"lo": lo, "hi": hi}) if lo < hi and clear_of(a, lo) and clear_of(a, hi) else None
        if op == "threshold":
            thr = float(np.round(rng.uniform(lo_v - 0.1, hi_v + 0.1), 3))
            return ((), {"thr": thr, "val": float(np.round(rng.uniform(-1.0, 1.0), 3))}) if clear_of(a, thr) else None
        cands = [j for j, b in enumerate(vals) if b.shape == a.shape]      # where() does not un-broadcast its gradients
        if not cands:
            return None
        j = int(rng.choice(cands))
        if op == "where_gt":
            c = float(np.round(rng.uniform(lo_v, hi_v), 3))
            return ((j,), {"c": c}) if clear_of(a, c) else None
        b = vals[j]
        c = float(np.round(rng.uniform(float(b.min()), float(b.max())), 3))
        return ((j,), {"c": c, "val": float(np.round(rng.uniform(-1.0, 1.0), 3))}) if clear_of(b, c) else None
    if op == "matmul":
        if nd == 0 or nd > 2:
            return None
        k = a.shape[-1]
        cands = [j for j, b in enumerate(vals) if b.ndim in (1, 2) and b.shape[0] == k and not (nd == 1 and b.ndim == 1)]
        if nd == 1:                           # 1-D operands: the reference squeezes the outer product, so no 1-sized dims
            cands = [j for j in cands if k > 1 and vals[j].shape[-1] > 1]
        else:
            cands = [j for j in cands if vals[j].ndim == 2 or (k > 1 and a.shape[0] > 1)]
        if not cands:
            return None
        return (int(rng.choice(cands)),), {}
    raise KeyError(op)


LARGE_DIMS = [1, 2, 7, 64, 100, 257, 512, 1024]
LARGE_MAX_ELEMS = 1 << 22
LARGE = dict(dims=LARGE_DIMS, max_elems=LARGE_MAX_ELEMS, min_base=1 << 16)


def make_program(seed: int, n_ops: int = 12, dims=DIMS, max_elems: int = MAX_ELEMS, min_base: int = 0,
                 select: bool = False) -> Program:
    """`**LARGE` gives the programs of the large variant: odd extents next to powers
    of two, up to 4 M elements per node - the vectorised / grid-stride / split-reduction paths and, through matmul, the
    tensor-core engines."""
    rng = np.random.default_rng(seed)
    prog = Program(seed)
    base = _rand_shape(rng, dims)
    while int(np.prod(base)) > max_elems or int(np.prod(base)) < min_base:
        base = _rand_shape(rng, dims)
    shapes = [base] + [_variant_shape(rng, base) for _ in range(int(rng.integers(1, 4)))]
    if rng.random() < 0.6:                                        # a partner for matmul
        shapes.append((base[-1], int(rng.choice(dims))))
    vals: List[np.ndarray] = []
    for shp in shapes:
        lo, hi = (-1.5, 1.5) if rng.random() < 0.25 else (0.5, 1.5)
        arr = rng.uniform(lo, hi, size=shp).astype(np.float32)
        prog.leaves.append(arr)
        vals.append(arr)
    used = set()
    grad = [True] * len(vals)                 # does the node require a gradient?  (sign() of anything does not)
    for _ in range(n_ops):
        for _attempt in range(30):
            op = _NAMES_SEL[int(rng.choice(len(_NAMES_SEL), p=_WEIGHTS_SEL))] if select else _NAMES[int(rng.choice(len(_NAMES), p=_WEIGHTS))]
            recent = rng.random() < 0.6
            i = int(rng.integers(max(0, len(vals) - 3), len(vals))) if recent else int(rng.integers(0, len(vals)))
            picked = _pick_args(rng, op, vals[i], vals, max_elems)
            if picked is None:
                continue
            more, args = picked
            ins = (i,) + tuple(more)
            try:
                with np.errstate(all="raise"):
                    out = np.asarray(_shadow(op, [vals[j] for j in ins], args), dtype=np.float32)
            except (FloatingPointError, ValueError, IndexError, ZeroDivisionError):
                continue
            if out.size == 0 or out.size > max_elems or not np.all(np.isfinite(out)) or np.abs(out).max() > MAX_ABS:
                continue
            prog.ops.append((op, ins, args))
            vals.append(out)
            grad.append(op != "sign" and any(grad[j] for j in (ins[:1] if op == "masked_fill" else ins)))
            used.update(ins)
            break
    n = len(vals)
    loose = [k for k in range(len(prog.leaves), n) if k not in used and grad[k]] or [max(k for k in range(n) if grad[k])]
    for k in loose[-3:]:
        w = rng.uniform(0.5, 1.5, size=vals[k].shape).astype(np.float32)
        prog.terminals.append((k, w))
    return prog


# --------------------------------------------------------------------------------------------------- execution
def to_host(a) -> np.ndarray:
    return np.asarray(a.get() if hasattr(a, "get") else a)


def run_program(prog: Program, Tensor, device: str = "cpu", dtype=None) -> Dict[str, Any]:
    kw = {} if dtype is None else {"dtype": dtype}
    cast = (lambda a: a.copy()) if dtype is None else (lambda a: a.astype(np.float64))
    nodes = [Tensor(cast(a), requires_grad=True, device=device, **kw) for a in prog.leaves]
    n_leaves = len(nodes)
    for op, ins, args in prog.ops:
        nodes.append(_apply(op, [nodes[j] for j in ins], args, Tensor))
    loss = None
    for k, w in prog.terminals:
        term = (nodes[k] * Tensor(cast(w), device=device, **kw)).sum()
        loss = term if loss is None else loss + term
    loss.backward()
    grads = []
    for t in nodes[:n_leaves]:
        g = t.grad
        grads.append(np.zeros(t.shape, np.float32) if g is None else to_host(g).reshape(t.shape))
    return {"values": [to_host(nodes[k].data) for k, _ in prog.terminals],
            "loss": float(to_host(loss.data)),
            "grads": grads}


def compare(got: Dict[str, Any], ref: Dict[str, Any], tol: float, what: str) -> None:
    """Norm-wise: max|got - ref| <= tol * max|ref| per array (cancellation inside one array is judged against that
    array's scale; a gradient that cancels as a whole against a twentieth of the largest leaf gradient)."""
    floor = [0.0]

    def close(g, r, name):
        g, r = np.asarray(g, dtype=np.float64), np.asarray(r, dtype=np.float64)
        assert g.shape == r.shape, f"{what}: {name} shape {g.shape} != {r.shape}"
        if r.size == 0:
            return
        scale = max(np.abs(r).max(), floor[0]) + 1e-30
        err = np.abs(g - r).max()
        assert err <= tol * scale + 1e-12, f"{what}: {name} off by {err / scale:.3e} (tol {tol:.1e})"
    for i, (g, r) in enumerate(zip(got["values"], ref["values"])):
        close(g, r, f"terminal {i}")
    close(got["loss"], ref["loss"], "loss")
    assert len(got["grads"]) == len(ref["grads"])
    floor[0] = 0.05 * max([float(np.abs(r).max()) for r in ref["grads"] if np.size(r)] + [1.0])   # leaves and weights are O(1)
    for i, (g, r) in enumerate(zip(got["grads"], ref["grads"])):
        close(g, r, f"grad of leaf {i}")


def well_conditioned(prog: Program, Tensor, base: Dict[str, Any], tol: float, f64=None) -> bool:
    """Re-runs the program on the CPU path in float64 (`f64` = the Tensor class's DType.FLOAT64): a program whose float32
    results sit further than `tol` (norm-wise) from the float64 ones is dominated by cancellation - it measures rounding
    order, not a backend."""
    try:
        exact = run_program(prog, Tensor, "cpu", dtype=f64)
        compare(base, exact, tol, "conditioning")
    except AssertionError:
        return False
    return True
