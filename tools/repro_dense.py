import sys, subprocess, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, torch
sys.path.insert(0, %r)
from sparse2dense_b200 import dense, ops
from torch import nn
B, H, W, cin, cout, prec = %d, %d, %d, %d, %d, %d
torch.manual_seed(0)
conv = nn.Conv2d(cin, cout, 3, 1, 1).cuda().eval()
x = torch.randn(B * H * W, cin, device="cuda")
D = dense.DenseOps(prec)
y, _, _ = D.conv("c", x, B, H, W, conv, None, dense.ACT_RELU)
torch.cuda.synchronize()
D2 = dense.DenseOps(ops.PRECISION_FP32)
r, _, _ = D2.conv("c", x, B, H, W, conv, None, dense.ACT_RELU)
print("err %%.2e" %% float((y - r).abs().max() / r.abs().max()))
'''
for (B, H, W, cin, cout) in [(1, 188, 188, 512, 64), (2, 188, 188, 512, 64), (3, 188, 188, 512, 64), (4, 188, 188, 512, 64),
                             (4, 188, 188, 64, 64), (4, 188, 188, 128, 64), (4, 188, 188, 512, 128)]:
    for prec in (2, 4):
        r = subprocess.run([sys.executable, "-c", CASE % (ROOT, B, H, W, cin, cout, prec)], capture_output=True, text=True, timeout=120)
        msg = r.stdout.strip() if r.returncode == 0 else "FAIL " + r.stderr.strip().splitlines()[-1][:120]
        print(B, H, W, cin, cout, "prec", prec, msg, flush=True)
