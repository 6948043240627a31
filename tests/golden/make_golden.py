#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference (nickovchinnikov/microtorch, NumPy backend) is imported unmodified under its
own package name ``src``; the three repairs of SURVEY.md section 8(c) are monkey-patched
in (never edited on disk):
  S1  np.generic gradients -> 0-d ndarray before Tensor.backward   (Q2, always)
  S2  Tensor.unsqueeze(dim=0, axis=None) alias                      (Q4a, 3-D only)
  S3  matmul grad_fns post-composed with BaseOps.bkwd_broadcast     (Q4b, 3-D only)
Outputs are small .npz/.json fixtures; inputs are regenerated from seeds by
microtorch_b200.workloads, and stored too where they are small.
"""

from __future__ import annotations

import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MICROTORCH_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True


def import_reference(repair_3d: bool = False):
    """Import the reference as `src` (fresh) and apply the shims."""
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import src.tensor.tensor as tmod
        import src.tensor.ops.overload as omod
        from src.tensor.ops.base import BaseOps
        from src.tensor.types import Leaf
        import src.nn.layer.linear  # noqa: F401  (pull nn in while REF is on the path)
        import src.nn.loss, src.nn.optimizer, src.nn.activation, src.nn.sequential  # noqa
    finally:
        sys.path.remove(REF)
    Tensor = tmod.Tensor

    orig_backward = Tensor.backward

    def backward_s1(self, grad=None):
        if isinstance(grad, np.generic):
            grad = np.asarray(grad)
        return orig_backward(self, grad)

    Tensor.backward = backward_s1

    if repair_3d:
        orig_unsq = Tensor.unsqueeze

        def unsqueeze_s2(self, dim=0, axis=None):
            return orig_unsq(self, dim if axis is None else axis)

        Tensor.unsqueeze = unsqueeze_s2

        orig_matmul = omod.OverloadOps.matmul

        def matmul_s3(a, b):
            props = orig_matmul(a, b)
            props.dependencies = [
                Leaf(value=l.value,
                     grad_fn=(lambda g, f=l.grad_fn, v=l.value: BaseOps.bkwd_broadcast(v)(f(g))))
                for l in props.dependencies
            ]
            return props

        omod.OverloadOps.matmul = staticmethod(matmul_s3)
    return tmod


def ref_classes():
    from src.nn.layer.linear import Linear
    from src.nn.activation import Tanh
    from src.nn.sequential import Sequential
    from src.nn.loss import MSELoss, CrossEntropyLoss, BCELoss
    from src.nn.optimizer import SGD
    return Linear, Tanh, Sequential, {"mse": MSELoss, "ce": CrossEntropyLoss, "bce": BCELoss}, SGD


def ref_run(w, steps):
    """Run workload `w` on the reference for `steps` steps (the notebook loop)."""
    from microtorch_b200 import workloads as W
    tmod = import_reference(repair_3d=w.input_shape is not None and len(w.input_shape) == 3)
    Tensor = tmod.Tensor
    Linear, Tanh, Sequential, losses, SGD = ref_classes()
    model, x, t = W.setup(w, Linear, Tanh, Sequential)
    X, T = Tensor(x), Tensor(t)
    loss_fn = losses[w.loss]()
    opt = SGD(model.parameters, lr=w.lr)
    out_losses = []
    first_pred = None
    for _ in range(steps):
        model.zero_grad()
        pred = model(X)
        if first_pred is None:
            first_pred = np.ascontiguousarray(pred.data).copy()
        if w.squeeze_dim is not None:
            pred = pred.squeeze(w.squeeze_dim)
        if w.pred_map is not None:
            pred = pred * w.pred_map[0] + w.pred_map[1]
        loss = loss_fn(pred, T)
        loss.backward()
        grads = [p.grad.copy() for p in model.parameters()]
        opt.step()
        out_losses.append(np.asarray(loss.data).copy())
    params = [p.data.copy() for p in model.parameters()]
    return {"losses": np.array(out_losses, np.float32), "grads": grads, "params": params,
            "first_pred": first_pred, "x": x, "t": t}


def save_run(name, res, with_inputs=True):
    d = {"losses": res["losses"], "first_pred": res["first_pred"]}
    for i, (g, p) in enumerate(zip(res["grads"], res["params"])):
        d[f"grad{i}"] = g
        d[f"param{i}"] = p
    if with_inputs:
        d["x"] = np.asarray(res["x"])
        d["t"] = np.asarray(res["t"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def golden_ops():
    """Op-level forward/backward vectors: the C3 expressions at small sizes with every
    broadcast pattern, plus reductions, getitem and shape ops."""
    tmod = import_reference()
    Tensor = tmod.Tensor
    rng = np.random.default_rng(0)
    out = {}
    shapes = {"same": ((6, 8), (6, 8)), "row": ((6, 8), (1, 8)), "col": ((6, 8), (6, 1)),
              "scalar": ((6, 8), ()), "lead": ((3, 6, 8), (8,)), "mid": ((3, 6, 8), (1, 6, 1)),
              "flat": ((48,), (48,))}
    for tag, (sx, sy) in shapes.items():
        xv = rng.uniform(0.5, 1.5, sx).astype(np.float32)
        yv = rng.uniform(0.5, 1.5, sy).astype(np.float32)
        x, y = Tensor(xv, requires_grad=True), Tensor(yv, requires_grad=True)
        z = x * y + x ** 2 - y
        z.sum().backward()
        out[f"poly_{tag}_x"], out[f"poly_{tag}_y"] = xv, yv
        out[f"poly_{tag}_z"], out[f"poly_{tag}_gx"], out[f"poly_{tag}_gy"] = z.data, x.grad, y.grad
    xv = rng.uniform(0.5, 1.5, (5, 7)).astype(np.float32)
    out["unary_x"] = xv
    for op in ("log", "tanh", "exp"):
        x = Tensor(xv, requires_grad=True)
        r = getattr(x, op)()
        r.sum().backward()
        out[f"{op}_y"], out[f"{op}_gx"] = r.data, x.grad
    for p in (2, 3, -1, 0.5, 1, 0, 2.5):
        x = Tensor(xv, requires_grad=True)
        r = x ** p
        r.sum().backward()
        out[f"pow_{p}_y"], out[f"pow_{p}_gx"] = r.data, x.grad
    x = Tensor(xv, requires_grad=True)
    m = x.mean()
    m.backward()
    out["mean_y"], out["mean_gx"] = np.asarray(m.data), x.grad
    for axis in (0, 1):
        for keep in (False, True):
            x = Tensor(xv, requires_grad=True)
            s = x.sum(axis=axis, keepdims=keep)
            w = rng.uniform(0.5, 1.5, s.shape).astype(np.float32)
            (s * Tensor(w)).sum().backward()
            out[f"sum_a{axis}_k{int(keep)}_w"] = w
            out[f"sum_a{axis}_k{int(keep)}_y"], out[f"sum_a{axis}_k{int(keep)}_gx"] = s.data, x.grad
    # division, rsub, neg chains (tensor.py:556-562, 601-607)
    yv = rng.uniform(0.5, 1.5, (5, 7)).astype(np.float32)
    x, y = Tensor(xv, requires_grad=True), Tensor(yv, requires_grad=True)
    r = (1 - x) / y + (-x)
    r.sum().backward()
    out["div_y_in"], out["div_r"], out["div_gx"], out["div_gy"] = yv, r.data, x.grad, y.grad
    # matmul 2-D fwd + both grads (overload.py:130-164)
    av = rng.standard_normal((5, 7)).astype(np.float32)
    bv = rng.standard_normal((7, 4)).astype(np.float32)
    a, b = Tensor(av, requires_grad=True), Tensor(bv, requires_grad=True)
    c = a @ b
    gw = rng.standard_normal((5, 4)).astype(np.float32)
    (c * Tensor(gw)).sum().backward()
    out.update(mm_a=av, mm_b=bv, mm_gw=gw, mm_c=c.data, mm_ga=a.grad, mm_gb=b.grad)
    # getitem forward + scatter-ASSIGN backward incl. duplicate fancy indices (Q10)
    x = Tensor(xv, requires_grad=True)
    s1 = x[1:4]
    s1.sum().backward()
    out["gi_slice_y"], out["gi_slice_gx"] = s1.data, x.grad
    x = Tensor(xv, requires_grad=True)
    s2 = x[[0, 2, 2, 4]]
    (s2 * Tensor(np.arange(28, dtype=np.float32).reshape(4, 7))).sum().backward()
    out["gi_fancy_y"], out["gi_fancy_gx"] = s2.data, x.grad
    x = Tensor(xv, requires_grad=True)
    s3 = x[2, 1:5]
    s3.sum().backward()
    out["gi_tuple_y"], out["gi_tuple_gx"] = s3.data, x.grad
    # shape ops: transpose / view / squeeze / unsqueeze round trips (base.py:82-187)
    x3 = rng.standard_normal((2, 3, 4)).astype(np.float32)
    x = Tensor(x3, requires_grad=True)
    t = x.transpose((2, 0, 1))
    gw = rng.standard_normal(t.shape).astype(np.float32)
    (t * Tensor(gw)).sum().backward()
    out.update(tr_x=x3, tr_gw=gw, tr_y=np.ascontiguousarray(t.data), tr_gx=x.grad)
    x = Tensor(x3, requires_grad=True)
    v = x.view((6, 4)).unsqueeze(1).squeeze(1).T
    gw = rng.standard_normal(v.shape).astype(np.float32)
    (v * Tensor(gw)).sum().backward()
    out.update(vw_gw=gw, vw_y=np.ascontiguousarray(v.data), vw_gx=x.grad)
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **{k: np.asarray(v) for k, v in out.items()})


def golden_losses():
    """Loss forward values + dL/dpred for MSE / CE(ref formula) / BCE, and the raw-CE NaN case."""
    tmod = import_reference()
    Tensor = tmod.Tensor
    _, _, _, losses, _ = ref_classes()
    rng = np.random.default_rng(1)
    out = {}
    p = rng.uniform(0.05, 0.95, (16, 10)).astype(np.float32)
    t = (rng.uniform(0, 1, (16, 10)) > 0.5).astype(np.float32)
    out["p"], out["t"] = p, t
    for k, cls in losses.items():
        pt = Tensor(p, requires_grad=True)
        l = cls()(pt, Tensor(t))
        l.backward()
        out[f"{k}_loss"], out[f"{k}_dp"] = np.asarray(l.data), pt.grad
    pn = rng.uniform(-0.9, 0.9, (16, 10)).astype(np.float32)   # tanh-range preds: NaN loss (Q3)
    pt = Tensor(pn, requires_grad=True)
    with np.errstate(all="ignore"):
        l = losses["ce"]()(pt, Tensor(t))
        l.backward()
    out["ce_nan_p"], out["ce_nan_loss"], out["ce_nan_dp"] = pn, np.asarray(l.data), pt.grad
    np.savez_compressed(os.path.join(HERE, "losses_small.npz"), **out)


def notebook_losses():
    """The reference's own published golden vector: demo.ipynb's printed loss stream."""
    nb = json.load(open(os.path.join(REF, "demo.ipynb")))
    text = ""
    for cell in nb["cells"]:
        for o in cell.get("outputs", []):
            if "text" in o:
                text += "".join(o["text"])
    vals = {int(m.group(1)): m.group(2) for m in re.finditer(r"Step (\d+), loss: ([0-9.eE+-]+)", text)}
    assert len(vals) == 5000, len(vals)
    keep = {str(i): vals[i] for i in list(range(0, 201)) + [1000, 4999]}
    json.dump({"source": "demo.ipynb stream output (CuPy run), seed 1337, MLP(2,[64,32,16,1]), lr 0.5",
               "losses": keep}, open(os.path.join(HERE, "c1a_notebook_losses.json"), "w"), indent=0)


def main():
    from microtorch_b200 import workloads as W
    notebook_losses()
    golden_ops()
    golden_losses()
    res = ref_run(W.C1A, 120)
    save_run("c1a_ref_120steps", res)
    save_run("c1b_ref_20steps", ref_run(W.C1B, 20))
    c2s = W.scaled(W.C2, 96, (48, 80, 64, 10))
    save_run("c2_small_3steps", ref_run(c2s, 3))
    with np.errstate(all="ignore"):
        c2r = W.scaled(W.C2_RAW, 96, (48, 80, 64, 10))
        save_run("c2raw_small_1step", ref_run(c2r, 1))
    c4s = W.scaled(W.C4, 4, (32, 48, 32), input_shape=(4, 6, 32))
    save_run("c4_small_2steps", ref_run(c4s, 2))
    c5s = W.scaled(W.C5, 64, (32,) * 5)
    save_run("c5_small_2steps", ref_run(c5s, 2))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
