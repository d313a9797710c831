#!/usr/bin/env python
"""Run one tcgen05 conv case per subprocess (a trap kills the CUDA context): tc_matrix.py [n_rows]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from sparse2dense_b200 import ops
cin, cout, prec, n = %d, %d, %d, %d
rng = np.random.default_rng(0)
vol = 2 * 9 * 64 * 64
lin = rng.permutation(rng.choice(vol, n, replace=False))
coors = np.stack([lin // (9*4096), (lin // 4096) %% 9, (lin // 64) %% 64, lin %% 64], 1).astype(np.int32)
c = torch.from_numpy(coors).cuda()
tbl = ops.rulebook_subm(c, ops.build_grid_index(c, 2, (9, 64, 64)), 3)
f = torch.randn(n, cin, device="cuda"); w = torch.randn(3, 3, 3, cin, cout, device="cuda") / (27 * cin) ** 0.5
ref = ops.spconv_fwd(f, w, tbl, n, precision=ops.PRECISION_FP32)
out = ops.spconv_fwd(f, w, tbl, n, precision=prec)
torch.cuda.synchronize()
print("err %%.2e" %% float((out - ref).abs().max() / ref.abs().max()))
'''


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3001
    shapes = [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)]
    if len(sys.argv) > 2:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2].split(",")]
    for cin, cout in shapes:
        for prec in (1, 2, 3):
            r = subprocess.run([sys.executable, "-c", CASE % (ROOT, cin, cout, prec, n)], capture_output=True, text=True,
                               timeout=120)
            msg = r.stdout.strip() if r.returncode == 0 else "FAIL " + r.stderr.strip().splitlines()[-1][:100]
            print(f"{cin:3d}->{cout:3d} prec {prec} n {n}: {msg}", flush=True)


if __name__ == "__main__":
    main()
