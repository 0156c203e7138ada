#!/usr/bin/env python
"""Known answers for rt_preprocess_bgr8 from OpenCV itself (cv2, the library sample_app/main.cpp:83-98 calls):
tests/golden/preprocess_cv2.npz holds small random 8-bit BGR images and what
    img.convertTo(CV_32F); cv::resize(.., INTER_AREA); cv::cvtColor(BGR2RGB); reshape(1, w*h).t(); res /= 255.0
gives for them (float32 [3,h,w]).  Needs only cv2 (present in this image), not the reference checkout."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def read_img_file(img_u8, w, h):
    img = img_u8.astype(np.float32)
    img = cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    res = img.reshape(w * h, 3).T.copy()
    res = cv2.multiply(res, 1.0 / 255.0)            # Mat /= 255.0  ==  convertTo(-1, 1/255.0)
    return res.reshape(3, h, w)


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    out = {}
    for i, ((sh, sw), (dh, dw)) in enumerate((((97, 131), (48, 64)), ((97, 131), (97, 131)), ((97, 131), (97, 65)),
                                               ((96, 128), (48, 64)), ((75, 249), (64, 205)), ((33, 47), (1, 1)))):
        src = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
        out["src_%d" % i] = src
        out["dst_%d" % i] = read_img_file(src, dw, dh)
    np.savez_compressed(os.path.join(HERE, "preprocess_cv2.npz"), **out)
    print({k: v.shape for k, v in out.items()})
