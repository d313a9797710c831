"""TEST INFRASTRUCTURE ONLY: numpy restatement of ``AssignLabel.__call__`` (det3d/datasets/pipelines/preprocess.py:489-653,
train mode, Waymo / nuScenes 10-column ``anno_box``) with ``gaussian_radius`` / ``gaussian2D`` / ``draw_umich_gaussian``
(det3d/core/utils/center_utils.py:18-64), ``limit_period`` (det3d/core/bbox/box_np_ops.py:360-361), ``flatten`` and
``merge_multi_group_label`` (preprocess.py:465-476).  Pinned against the reference's own code executed in the authoring
container (tests/golden/assign_label.npz, ``make_golden.py assign``).

Arithmetic notes kept from the reference (as it evaluates under the numpy of this image, where a float32 scalar times a
python float stays float32): box sizes, centres and the gaussian radius are float32 expressions in the written order;
the gaussian itself is float64 (``np.ogrid`` of python floats) rounded to float32 by ``np.maximum(..., out=float32)``."""
import numpy as np

f32 = np.float32


def gaussian_radius(height, width, min_overlap):
    """center_utils.py:18-39 on float32 scalars."""
    height, width = f32(height), f32(width)
    mo = float(min_overlap)
    b1 = height + width
    c1 = width * height * f32(1 - mo) / f32(1 + mo)
    sq1 = np.sqrt(b1 ** 2 - f32(4) * c1)
    r1 = (b1 + sq1) / f32(2)
    b2 = f32(2) * (height + width)
    c2 = f32(1 - mo) * width * height
    sq2 = np.sqrt(b2 ** 2 - f32(16) * c2)
    r2 = (b2 + sq2) / f32(2)
    a3 = f32(4 * mo)
    b3 = f32(-2 * mo) * (height + width)
    c3 = f32(mo - 1) * width * height
    sq3 = np.sqrt(b3 ** 2 - f32(4) * a3 * c3)
    r3 = (b3 + sq3) / f32(2)
    return min(r1, r2, r3)


def draw_gaussian(hm, cx, cy, radius):
    """draw_umich_gaussian + gaussian2D (center_utils.py:41-64): float64 gaussian, float32 running maximum."""
    H, W = hm.shape
    sigma = (2 * radius + 1) / 6
    left, right = min(cx, radius), min(W - cx, radius + 1)
    top, bottom = min(cy, radius), min(H - cy, radius + 1)
    if left + right <= 0 or top + bottom <= 0:
        return
    ys = np.arange(-top, bottom, dtype=np.float64)[:, None]
    xs = np.arange(-left, right, dtype=np.float64)[None, :]
    g = np.exp(-(xs * xs + ys * ys) / (2 * sigma * sigma))
    win = hm[cy - top:cy + bottom, cx - left:cx + right]
    np.maximum(win, g.astype(np.float32), out=win)


def assign_label(gt_boxes, gt_classes, class_counts, grid_xy, pc_range, voxel_size, out_size_factor, gaussian_overlap=0.1,
                 max_objs=500, min_radius=2):
    """One sample.  gt_boxes f32 [n,9] (x,y,z,w,l,h,vx,vy,rot), gt_classes int [n] (1-based over all tasks),
    class_counts: classes per task -> dict(hm, anno_box, ind, mask, cat: lists per task; gt_boxes_and_cls)."""
    gt_boxes = np.asarray(gt_boxes, np.float32)
    gt_classes = np.asarray(gt_classes)
    W, H = int(grid_xy[0]) // out_size_factor, int(grid_xy[1]) // out_size_factor
    pc_range, voxel_size = np.asarray(pc_range, np.float32), np.asarray(voxel_size, np.float32)
    out = dict(hm=[], anno_box=[], ind=[], mask=[], cat=[])
    flag = 0
    all_boxes, all_cls = [], []
    for ncls in class_counts:
        sel = [np.where(gt_classes == c + 1 + flag)[0] for c in range(ncls)]              # class-major order inside a task
        boxes = np.concatenate([gt_boxes[m] for m in sel], 0).copy()
        cls = np.concatenate([gt_classes[m] - flag for m in sel])
        period = f32(np.pi * 2)
        boxes[:, -1] = boxes[:, -1] - np.floor(boxes[:, -1] / period + f32(0.5)) * period
        hm = np.zeros((ncls, H, W), np.float32)
        anno = np.zeros((max_objs, 10), np.float32)
        ind = np.zeros((max_objs,), np.int64)
        mask = np.zeros((max_objs,), np.uint8)
        cat = np.zeros((max_objs,), np.int64)
        for k in range(min(boxes.shape[0], max_objs)):
            b = boxes[k]
            w = b[3] / voxel_size[0] / f32(out_size_factor)
            l = b[4] / voxel_size[1] / f32(out_size_factor)
            if not (w > 0 and l > 0):
                continue
            radius = max(min_radius, int(gaussian_radius(l, w, gaussian_overlap)))
            cx = (b[0] - pc_range[0]) / voxel_size[0] / f32(out_size_factor)
            cy = (b[1] - pc_range[1]) / voxel_size[1] / f32(out_size_factor)
            ix, iy = int(np.int32(cx)), int(np.int32(cy))
            if not (0 <= ix < W and 0 <= iy < H):
                continue
            draw_gaussian(hm[int(cls[k]) - 1], ix, iy, radius)
            cat[k], ind[k], mask[k] = int(cls[k]) - 1, iy * W + ix, 1
            anno[k] = [cx - f32(ix), cy - f32(iy), b[2], np.log(b[3]), np.log(b[4]), np.log(b[5]), b[6], b[7],
                       np.sin(b[8]), np.cos(b[8])]
        for key, v in zip(("hm", "anno_box", "ind", "mask", "cat"), (hm, anno, ind, mask, cat)):
            out[key].append(v)
        all_boxes.append(boxes)
        all_cls.append(cls + flag)
        flag += ncls
    boxes = np.concatenate(all_boxes, 0)
    classes = np.concatenate(all_cls).reshape(-1, 1).astype(np.float32)
    bc = np.zeros((max_objs, 10), np.float32)
    both = np.concatenate([boxes, classes], 1)[:, [0, 1, 2, 3, 4, 5, 8, 6, 7, 9]]
    assert len(both) <= max_objs
    bc[:len(both)] = both
    out["gt_boxes_and_cls"] = bc
    return out
