"""TEST INFRASTRUCTURE -- CPU oracle for K10 (fm3d_detect_fast), numpy only.

Restates what feature_detector_->detect computes in DescriptorsMatcher::compareWithNNDR / compare /
crosscompare (DescriptorsMatcher/descriptorsmatcher.cpp:110-111, :92-93, :77-78) when DetectorType is
FAST with DetectorMode STATIC (:215-222: cv::FastFeatureDetector(FeatureOptions.FastDetector.Threshold,
FeatureOptions.FastDetector.NonMaxSuppression > 0)).  OpenCV is a third-party dependency of the
reference (unpinned, 2.4.x era); the published algorithm restated here is cv::FAST's 9-16 variant
(modules/features2d/src/fast.cpp FAST_t<16>, fast_score.cpp cornerScore<16>; unchanged between 2.4 and
the 4.13 installed in this image):

  * the threshold is clamped to [0, 255];
  * a pixel with 3 <= x < w - 3, 3 <= y < h - 3 is a corner iff 9 contiguous pixels of the 16-pixel
    circle of radius 3 all satisfy x_k < v - t, or all satisfy x_k > v + t;
  * cornerScore = max(t, max_arcs min_k (v - x_k), max_arcs min_k (x_k - v)) - 1 over the 16 arcs of 9;
  * with non-maximum suppression a corner is kept iff its score is strictly greater than the score of
    each of its 8 neighbours (non-corners count as 0) and response = score; without it every corner is
    kept with response 0;
  * keypoints are reported in row-major order (ascending y, then ascending x), size 7, angle -1.

Pinned by: cv2.FastFeatureDetector_create itself (tests/test_oracle_pins.py, where cv2 is importable) and
the committed golden vectors tests/golden/fast_keypoints.npz written from cv2 by tools/make_golden.py;
integer work, agreement is exact.  Only tests/ (and __graft_entry__.smoke) may import this module.
"""
from __future__ import annotations

import numpy as np

# circle offsets (dx, dy) in cv::FAST's order (fast_score.cpp makeOffsets, patternSize 16)
CIRCLE = ((0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
          (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3))


def fast_score_map(img, threshold):
    """H x W int32: cornerScore + 1 at corners, 0 elsewhere (so that a score of 0 is still a corner)."""
    img = np.asarray(img)
    assert img.ndim == 2 and img.dtype == np.uint8
    h, w = img.shape
    t = int(min(max(int(threshold), 0), 255))
    out = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return out
    v = img[3:h - 3, 3:w - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int32) for dx, dy in CIRCLE])     # v - x_k
    arc_min = np.full(v.shape, -1000, np.int32)      # max over arcs of min(v - x)
    arc_max = np.full(v.shape, 1000, np.int32)       # min over arcs of max(v - x)
    for k in range(16):
        idx = [(k + j) & 15 for j in range(9)]
        arc_min = np.maximum(arc_min, d[idx].min(0))
        arc_max = np.minimum(arc_max, d[idx].max(0))
    corner = (arc_min > t) | (arc_max < -t)          # 9 contiguous darker / brighter, strict
    score = np.maximum(np.maximum(t, arc_min), -arc_max) - 1
    out[3:h - 3, 3:w - 3] = np.where(corner, score + 1, 0)
    return out


def detect_fast(img, threshold, nonmax=True):
    """-> (n x 2 float32 xy, n float32 response), row-major order."""
    smap = fast_score_map(img, threshold)
    keep = smap > 0
    if nonmax:
        s = np.maximum(smap - 1, 0)                 # non-corners (and score-0 corners) compare as 0
        p = np.pad(s, 1)
        h, w = s.shape
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx or dy:
                    keep &= s > p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
    ys, xs = np.nonzero(keep)                       # np.nonzero is row-major
    xy = np.stack([xs, ys], 1).astype(np.float32).reshape(-1, 2)
    resp = (smap[ys, xs] - 1).astype(np.float32) if nonmax else np.zeros(len(xs), np.float32)
    return xy, resp
