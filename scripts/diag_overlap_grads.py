"""Diagnostic: gradient of one full-size pass under the overlap mechanisms vs the plain single-stream pass."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pn2_b200
from pn2_b200 import _ffi
from pn2_b200.train_step import Trainer

A = [torch.as_tensor(x).cuda() for x in bench.make_batch(16, 8192, 100)]
B = [torch.as_tensor(x).cuda() for x in bench.make_batch(16, 8192, 1100)]


def grads(tag, budget=0, **kw):
    tr = Trainer(bench.HP, bench.NUM_CLASS, device="cuda", seed=0, world_size=1, **kw)
    tr._seed_dev.add_(1)
    tr.forward_backward(*A)          # creates the variables
    if budget:
        _ffi.lib().pn2_set_sm_budget(budget)
        real = tr._sm_budget
    loss = float(tr.forward_backward(*A).item())
    _ffi.lib().pn2_set_sm_budget(0)
    torch.cuda.synchronize()
    return tr, loss, tr.grads.clone()


def report(tag, tr, l, g, l0, g0):
    names = [(k, v) for k, v in tr.store.vars.items() if v.trainable]
    off, worst = 0, []
    for k, v in names:
        n = v.data.numel()
        a, b = g[off:off + n], g0[off:off + n]
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        worst.append((rel, k))
        off += n
    worst.sort(reverse=True)
    print("%-28s loss diff %.2e  max|dg| %.2e (gmax %.2e)  worst rel-L2: %s" % (
        tag, abs(l - l0), float((g - g0).abs().max()), float(g0.abs().max()),
        ", ".join("%s %.1e" % (k, r) for r, k in worst[:4])), flush=True)



import pn2_b200.util.tf_util as tfu
real_call = _ffi.call
only = [None, 0]


def patched(name, *a):
    if only[0] is not None and name in only[0]:
        _ffi.lib().pn2_set_sm_budget(only[1])
        try:
            return real_call(name, *a)
        finally:
            _ffi.lib().pn2_set_sm_budget(0)
    return real_call(name, *a)


for mod in (tfu, _ffi):
    mod.call = patched


store = {}
cur = [None]
real_bwd = tfu._MLPChain.backward


def spy(ctx, d_out):
    key = ctx.layers[0].w.name
    rec = {"d_out": d_out.clone(), "x": ctx.x.clone()}
    for i, (Y, sc, sh, sv) in enumerate(zip(ctx.Ys, ctx.scs, ctx.shs, ctx.saveds)):
        for nm, t in (("Y", Y), ("sc", sc), ("sh", sh), ("saved", sv)):
            if t is not None:
                rec["%s%d" % (nm, i)] = t.clone()
    if ctx.arg is not None:
        rec["arg"] = ctx.arg.clone()
    cur[0][key] = rec
    return real_bwd(ctx, d_out)


tfu._MLPChain.backward = staticmethod(spy)
cur[0] = store.setdefault("plain", {})
tr0, l0, g0 = grads("plain", wgrad_sms=0)
only[0], only[1] = ("pn2_linear_fwd_bn", "pn2_linear_fwd"), 132
cur[0] = store.setdefault("b132", {})
t, l, g = grads("x", wgrad_sms=0)
only[0] = None
report("fwd @ 132", t, l, g, l0, g0)
# the passes ran twice (variable creation + measured): the records hold the second
for key in store["plain"]:
    a, b = store["plain"][key], store["b132"][key]
    bad = []
    for nm in a:
        if not torch.equal(a[nm], b[nm]):
            d = (a[nm].double() - b[nm].double()).abs().max().item()
            bad.append("%s max|d| %.2e (max %.2e)" % (nm, d, a[nm].double().abs().max().item()))
    print(key, "DIFF: " + "; ".join(bad) if bad else "identical", flush=True)
