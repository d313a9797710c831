"""GPU parity at the BASELINE config sizes: the whole two-stage forward at batch 8 (configs[2]) and the
pillar student at batch 16 (configs[3]) against the CPU restatement (oracle/full_forward.py) on detections.

The per-stage parity tests (voxel indices / rulebook bit-exact, conv <= 2e-5, neck/head vs the reference fixture, NMS keep
sets exact on a fixed input) are the tight ones; here the stages are chained end to end with random weights, so two effects
remain and the bars are written for them: (1) a candidate whose score sits within fp32 rounding of the 0.1 threshold or
whose IoU with a kept box sits within rounding of 0.7 can flip, (2) random head weights produce a few overflowing box
dimensions (exp of a large logit).  Bar: >= 97 % of the oracle's detections of a scene are found by the GPU path (same label,
centre within 2 cm, score within 2e-3) and vice versa, and the number of detections differs by <= 3 %."""
import numpy as np
import pytest
import torch

from oracle import full_forward as FF
from oracle import ref_ops as R
from sparse2dense_b200 import ops, synth
from sparse2dense_b200.hotpath import FullForwardPath, PillarForwardPath, concat_clouds

pytestmark = pytest.mark.gpu


def _match(got, ref, tol_xy=2e-2, tol_s=2e-3):
    """Fraction of ``ref`` detections that have a partner in ``got`` (label equal, finite centre within tol, score within tol)."""
    gb, gs, gl = got
    rb, rs, rl = ref
    if len(rb) == 0:
        return 1.0
    used = np.zeros(len(gb), bool)
    hit = 0
    for i in range(len(rb)):
        if not np.all(np.isfinite(rb[i, :2])):
            hit += 1
            continue
        d = np.abs(gb[:, :2] - rb[i, :2]).max(1)
        ok = (d < tol_xy) & (gl == rl[i]) & (np.abs(gs - rs[i]) < tol_s) & ~used
        j = np.flatnonzero(ok)
        if j.size:
            used[j[0]] = True
            hit += 1
    return hit / len(rb)


def _check_scene(out, b, ref, roi_state, pc, vs, stride):
    """First stage (decode + NMS output) against the oracle's first stage by matching; second stage against the oracle's
    second stage run on the GPU path's OWN first-stage boxes and BEV rows (so that the comparison is not dominated by the
    random RoI head amplifying a first-stage difference)."""
    boxes, scores, labels, counts, rois, roi_scores, bev = out
    k = int(counts[b])
    got = (rois[b, :k].cpu().numpy(), roi_scores[b, :k].cpu().numpy(), labels[b, :k].cpu().numpy())
    refs = (np.asarray(ref["boxes"]), np.asarray(ref["scores"]), np.asarray(ref["labels"]))
    assert abs(k - len(refs[0])) <= max(2, 0.03 * len(refs[0])), (k, len(refs[0]))
    f1, f2 = _match(got, refs), _match(refs, got)
    assert f1 >= 0.97 and f2 >= 0.97, (f1, f2)
    f = R.roi_features(bev[b].cpu().numpy(), got[0], pc, vs, stride)
    cls, reg = R.roi_head_forward(roi_state, f)
    ob, sc = R.roi_refine(got[0], got[1], cls, reg)
    gb = boxes[b, :k].cpu().numpy()
    # random head weights: some boxes carry exp(large logit) dimensions (1e4 .. inf); rows finite on both sides (>= 90 %) are
    # compared with a per-row scale (the refinement couples a row's entries through the box diagonal)
    fin = np.isfinite(ob).all(1) & np.isfinite(gb).all(1)
    assert fin.mean() > 0.9, fin.mean()
    scale = np.maximum(1.0, np.abs(ob[fin]).max(1, keepdims=True))
    assert (np.abs(gb[fin] - ob[fin]) <= 1e-3 * scale).all(), float((np.abs(gb[fin] - ob[fin]) / scale).max())
    np.testing.assert_allclose(scores[b, :k].cpu().numpy()[fin], sc[fin], rtol=0, atol=1e-3)
    return f1, f2


def test_full_forward_batch8_detections_vs_oracle():
    """BASELINE configs[2]: batch 8 x ~180k points through FullForwardPath (the object bench.py times)."""
    R.build()
    path = FullForwardPath.synthetic(precision=ops.PRECISION_AUTO)
    clouds = synth.lidar_batch(1, 8)
    pts, offs = concat_clouds(clouds)
    before = ops.kernel_launches()
    out = path.forward_points(pts.cuda(), offs, return_first_stage=True)
    torch.cuda.synchronize()
    assert ops.kernel_launches() - before > 100
    counts = out[3].cpu().numpy()
    assert out[0].shape == (8, 500, 7) and counts.min() > 50, counts
    states = path.states_numpy()
    for b in (0, 7):
        ref = FF.scene_forward(states, clouds[b], second_stage=False)
        print("scene", b, "gpu", int(counts[b]), "oracle", len(ref["scores"]),
              _check_scene(out, b, ref, states["roi"], (-75.2, -75.2), (0.1, 0.1), 8))


def test_pillar_forward_batch16_detections_vs_oracle():
    """BASELINE configs[3]: the pillar student's global batch of 16 on one GPU."""
    R.build()
    path = PillarForwardPath(precision=ops.PRECISION_AUTO)
    clouds = synth.lidar_batch(3, 16)
    pts, offs = concat_clouds(clouds)
    out = path.forward_points(pts.cuda(), offs, return_first_stage=True)
    torch.cuda.synchronize()
    counts = out[3].cpu().numpy()
    assert out[0].shape == (16, 500, 7) and counts.min() > 20, counts
    states = path.states_numpy()
    for b in (0, 15):
        ref = FF.pillar_scene_forward(states, clouds[b], second_stage=False)
        print("scene", b, "gpu", int(counts[b]), "oracle", len(ref["scores"]),
              _check_scene(out, b, ref, states["roi"], (-74.88, -74.88), (0.32, 0.32), 1))
