"""Slowest conv_launch / conv_wgrad calls of one distillation step (shapes + CUDA-event times)."""
import sys, os, torch
sys.path.insert(0, '/root/repo')
import bench
from sparse2dense_b200 import ops, autograd as AG, _lib
wl = bench.TrainWorkload("train", 4, ops.PRECISION_AUTO, torch.device("cuda", 0), 0, 1)
for _ in range(2): wl.step()
torch.cuda.synchronize()
log = []
orig_w = AG.conv_wgrad
def w(g, d, tbl, n_rows, d_rows=None, precision=ops.PRECISION_AUTO):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig_w(g, d, tbl, n_rows, d_rows, precision); e1.record()
    log.append(("wgrad", g.shape[1], d.shape[1], tbl.shape[0], n_rows, e0, e1)); return out
AG.conv_wgrad = w
orig_l = ops.conv_launch
def l(x, w_arg, tbl, n_out, cin, cout, k, *a, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prec = kw.get("precision", a[8] if len(a) > 8 else None)
    prec = prec if isinstance(prec, int) else -1
    e0.record(); r = orig_l(x, w_arg, tbl, n_out, cin, cout, k, *a, **kw); e1.record()
    log.append(("conv", cin, cout, k, n_out, e0, e1, prec)); return r
ops.conv_launch = l
import sparse2dense_b200.dense as D
D.ops.conv_launch = l
wl.step(); torch.cuda.synchronize()
rows = []
for t in log:
    ms = t[5].elapsed_time(t[6])
    rows.append((ms, t[0], t[1], t[2], t[3], t[4], t[7] if len(t) > 7 else None))
rows.sort(key=lambda r: -r[0])
for r in rows[:40]:
    print("%8.3f ms %s Cin/Cg %4d Cout/Cd %4d K %3d rows %9d prec %s" % r)
