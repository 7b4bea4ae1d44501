"""Diagnostic: capture the training step in stages to find what invalidates the capture."""
import os, sys, traceback
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pn2_b200
from pn2_b200 import model
from pn2_b200.train_step import Trainer
from pn2_b200.util import tf_util

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pc, labels, smpw = bench.make_batch(b, 8192, 100)
dev = torch.device("cuda", 0)
d_pc, d_lab, d_w = (torch.as_tensor(x).to(dev) for x in (pc, labels, smpw))
tr = Trainer(bench.HP, bench.NUM_CLASS, device=dev, seed=0, world_size=1)
tr.step(d_pc, d_lab, d_w); tr.step(d_pc, d_lab, d_w)
torch.cuda.synchronize()

def attempt(name, fn, mode="global"):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("%-28s capture OK" % name)
    except Exception:
        print("%-28s FAILED\n%s" % (name, "".join(traceback.format_exc().splitlines(True)[-6:])))
        try: torch.cuda.synchronize()
        except Exception: pass

def fwd_only():
    with torch.no_grad():
        return model.get_model(d_pc, True, 9, bench.HP, bn_decay=0.5)[0]
def fwd_grad():
    return model.get_model(d_pc, True, 9, bench.HP, bn_decay=0.5)[0]
def fwd_loss():
    pred, _ = model.get_model(d_pc, True, 9, bench.HP, bn_decay=0.5)
    return model.get_loss(pred, d_lab, d_w)
def fwd_bwd():
    tr.store.zero_grad()
    loss = fwd_loss(); loss.backward(); return loss

attempt("forward + loss + backward", fwd_bwd)
def tr_fb():
    return tr.forward_backward(d_pc, d_lab, d_w)
attempt("Trainer.forward_backward", tr_fb)
stat = [t.clone() for t in (d_pc, d_lab, d_w)]
def tr_fb_static():
    return tr.forward_backward(*stat)
attempt("same on cloned inputs", tr_fb_static)
tf_util.set_dropout_seed_device(torch.zeros(1, dtype=torch.int64, device=dev))
attempt("with device dropout seed", tr_fb_static)
tf_util.set_dropout_seed_device(None)

print("---- Trainer.capture()")
tr2 = Trainer(bench.HP, bench.NUM_CLASS, device=dev, seed=0, world_size=1)
tr2.step(d_pc, d_lab, d_w)
ok = tr2.capture(d_pc, d_lab, d_w)
print("capture ->", ok, getattr(tr2, "_capture_error", None))
if ok:
    for _ in range(3):
        l = tr2.step_graph(d_pc, d_lab, d_w)
    torch.cuda.synchronize(); print("graph steps ok, loss", float(l))
