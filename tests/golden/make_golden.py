#!/usr/bin/env python3
"""Generates the committed golden vectors.  Primitive outputs come from cv2 (4.13.0 here) -- the
third-party library the reference calls (cv::resize / GaussianBlur / FAST / fastAtan2); the
full-frame vector comes from the oracle itself once the primitives are pinned."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from manhattanslam_b200 import synthetic as S  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fast(img, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FastFeatureDetector_TYPE_9_16)
    return np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)], np.int32).reshape(-1, 3)


def main():
    seed = 42
    img = S.gray_frame(seed)
    r = np.random.default_rng(9)
    atan_in = r.integers(-200000, 200000, (512, 2)).astype(np.float32)
    atan_out = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in atan_in], np.float32)
    np.savez_compressed(os.path.join(HERE, "orb_primitives.npz"), seed=seed, cv2_version=cv2.__version__,
                        resized=cv2.resize(img, (533, 400), interpolation=cv2.INTER_LINEAR),
                        blurred=cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101),
                        fast20=fast(img, 20), fast7_roi=fast(np.ascontiguousarray(img[100:137, 200:237]), 7),
                        atan_in=atan_in, atan_out=atan_out)
    kps, desc = ob.OrbOracle()(img)
    np.savez_compressed(os.path.join(HERE, "orb_frame.npz"), seed=seed, kps=kps.view(np.uint8).reshape(len(kps), -1),
                        desc=desc)
    print("golden written; cv2", cv2.__version__, "kps", len(kps))


if __name__ == "__main__":
    main()
