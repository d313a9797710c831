"""CPU: the decode / rotated-IoU / NMS oracle against fixtures generated from the reference's own code
(tests/golden/make_golden.py predict: iou3d_cpu.cpp compiled unmodified into oracle/_ref, CenterHead.predict imported
through the shim)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ref_ops as R

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
TEST_RANGE = [-80, -80, -10.0, 80, 80, 10.0]


def test_iou_oracle_bitwise_equals_reference_fixture():
    d = np.load(os.path.join(G, "iou_bev_pairs.npz"))
    got = R.iou_bev_matrix(d["boxes_a"], d["boxes_b"])
    assert (got.view(np.uint32) == d["iou"].view(np.uint32)).all()
    assert (d["iou"] > 0.7).sum() > 10 and (d["iou"] > 0).sum() > 500          # the fixture exercises real overlaps


def test_iou_oracle_equals_compiled_reference_when_present():
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_iou3d.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    lib = ctypes.CDLL(so)
    rng = np.random.default_rng(11)
    n = 200
    a = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(-1, 1, n), rng.uniform(0.5, 5, n),
                  rng.uniform(0.5, 2.5, n), rng.uniform(1, 2, n), rng.uniform(-4, 4, n)], 1).astype(np.float32)
    f32p = ctypes.POINTER(ctypes.c_float)
    ref = np.empty((n, n), np.float32)
    lib.ref_iou_bev_matrix(a.ctypes.data_as(f32p), n, a.ctypes.data_as(f32p), n, ref.ctypes.data_as(f32p))
    got = R.iou_bev_matrix(a, a)
    assert (got.view(np.uint32) == ref.view(np.uint32)).all()
    assert (ref > 0).mean() > 0.03


def test_iou_basic_properties():
    box = np.array([[1.0, 2.0, 0.0, 4.0, 2.0, 1.5, 0.3]], np.float32)
    assert abs(R.iou_bev_matrix(box, box)[0, 0] - 1.0) < 1e-5
    far = box.copy(); far[0, 0] += 100
    assert R.iou_bev_matrix(box, far)[0, 0] == 0.0
    half = box.copy(); half[0, 6] = 0; b2 = half.copy(); b2[0, 0] += 2.0       # axis-aligned, shifted by half the length
    assert abs(R.iou_bev_matrix(half, b2)[0, 0] - (4.0 / 12.0)) < 1e-3       # MARGIN widens the corner test slightly


def test_predict_oracle_reproduces_reference_predict_fixture():
    d = np.load(os.path.join(G, "centerhead_predict.npz"))
    maps = {k: np.ascontiguousarray(d["in_" + k].transpose(0, 2, 3, 1)) for k in ("reg", "height", "dim", "rot", "hm")}
    boxes, hm = R.centerhead_decode(maps, 8, [0.1, 0.1], [-75.2, -75.2])
    for i in range(boxes.shape[0]):
        det, margin = R.post_processing(boxes[i], hm[i], 0.1, TEST_RANGE, 0.7, 4096, 500)
        assert np.array_equal(det["label_preds"], d[f"labels_{i}"])
        np.testing.assert_allclose(det["box3d_lidar"], d[f"boxes_{i}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(det["scores"], d[f"scores_{i}"], rtol=1e-6, atol=1e-7)
        assert len(det["scores"]) > 100 and margin > 1e-5


def test_nms_oracle_protocol():
    rng = np.random.default_rng(2)
    n = 300
    b = np.stack([rng.uniform(-8, 8, n), rng.uniform(-8, 8, n), np.zeros(n), rng.uniform(2, 5, n), rng.uniform(1, 2.5, n),
                  np.ones(n), rng.uniform(-3, 3, n)], 1).astype(np.float32)
    keep, _ = R.nms_sorted(b, 0.3)
    assert keep[0] == 0 and (np.diff(keep) > 0).all() and len(keep) < n
    iou = R.iou_bev_matrix(b[keep], b[keep])
    assert (iou[np.triu_indices(len(keep), 1)] <= 0.3).all()                  # survivors do not suppress each other
    keep2, _ = R.nms_sorted(b[keep], 0.3)
    assert np.array_equal(keep2, np.arange(len(keep)))                        # idempotent
    removed = np.setdiff1d(np.arange(n), keep)
    full = R.iou_bev_matrix(b[removed], b[keep])
    for r, row in zip(removed, full):                                         # every removed box has an earlier suppressor
        assert (row[keep < r] > 0.3).any()
    assert len(R.nms_sorted(np.zeros((0, 7), np.float32), 0.5)[0]) == 0
