import sys, time, torch
sys.path.insert(0, '/root/repo')
from sparse2dense_b200 import ops, synth
from sparse2dense_b200.hotpath import VoxelBackbonePath, concat_clouds
path = VoxelBackbonePath(state=synth.backbone_state(0), precision=ops.PRECISION_AUTO)
pts_host, offs = concat_clouds(synth.lidar_batch(1, 4), pin=True)
pts_dev = pts_host.cuda()
outs = [torch.empty((4, 256, 188, 188), pin_memory=True) for _ in range(2)]
print("pinned:", outs[0].is_pinned(), pts_host.is_pinned())
def run(name, fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); print(f"{name}: {(time.perf_counter() - t) / n * 1e3:.2f} ms/step")
run("device only", lambda i: path.forward_points(pts_dev, offs))
run("h2d + compute", lambda i: path.forward_points(pts_host.to('cuda', non_blocking=True), offs))
run("sync d2h", lambda i: path.forward_host(pts_host, offs, outs[i % 2]))
run("async d2h", lambda i: path.forward_host_async(pts_host, offs, outs[i % 2]))
cs = torch.cuda.Stream()
bev = path.forward_points(pts_dev, offs)
def only_copy(i):
    with torch.cuda.stream(cs):
        outs[i % 2].copy_(bev, non_blocking=True)
run("d2h copy alone (copy stream)", only_copy)
