"""CPU: second-stage oracle (box sample points, bilinear BEV features, RoI MLP, box refinement) against the fixture
generated with the reference's own TwoStageDetector / BEVFeatureExtractor / RoIHead code (make_golden.py second)."""
import os

import numpy as np

from oracle import ref_ops as R

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "two_stage.npz")
PC, VS, STRIDE = [-75.2, -75.2], [0.1, 0.1], 8


def load():
    d = np.load(G)
    state = {k[4:]: d[k] for k in d.files if k.startswith("roi.")}
    return d, state


def test_second_stage_oracle_reproduces_reference_fixture():
    d, state = load()
    for b in range(2):
        boxes = d[f"in_boxes_{b}"]
        f = R.roi_features(np.ascontiguousarray(d["bev"][b].transpose(1, 2, 0)), boxes, PC, VS, STRIDE)
        ref = d[f"roi_features_{b}"]
        assert f.shape == ref.shape == (len(boxes), 5 * d["bev"].shape[1])
        assert np.abs(f - ref).max() <= 1e-5 * np.abs(ref).max()
        cls, reg = R.roi_head_forward(state, f)
        ob, sc = R.roi_refine(boxes, d[f"in_scores_{b}"], cls, reg)
        np.testing.assert_allclose(ob, d[f"boxes_{b}"], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(sc, d[f"scores_{b}"], rtol=1e-5, atol=1e-6)


def test_sample_points_geometry():
    box = np.array([[10.0, -3.0, 0.5, 4.0, 2.0, 1.5, 0.0]], np.float32)
    pts = R.box_sample_points(box)[:, 0]                        # heading 0: corners are axis aligned
    np.testing.assert_allclose(pts, [[10, -3], [8, -3], [12, -3], [10, -4], [10, -2]], atol=1e-6)
    im = np.arange(12, dtype=np.float32).reshape(3, 4, 1)
    np.testing.assert_allclose(R.bilinear_sample(im, np.array([1.5]), np.array([0.5]))[0, 0], 3.5)
    # clamp-before-weights quirk (center_utils.py:101-120): outside the map the weights no longer sum to one
    v = R.bilinear_sample(im, np.array([-0.5]), np.array([0.0]))[0, 0]
    assert v == 0.0 and R.bilinear_sample(im, np.array([3.5]), np.array([2.0]))[0, 0] != im[2, 3, 0]
