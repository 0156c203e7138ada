"""Generates the TrailNet fixtures (run where /root/reference and cv2 exist; the GPU box only reads the committed files):
  tests/golden/trailnet/sresnet18_deploy.prototxt.gz, sresnet18_weights.caffemodel   the reference's model files
                                        (models/pretrained/TrailNet_SResNet-18.{prototxt,caffemodel}), content unchanged
  tests/golden/trailnet/inputs.npz      the five test images of ros/packages/caffe_ros/tests/data as the network sees them:
                                        cv::imread -> float -> cv::resize(320x180, INTER_CUBIC) -> CHW, BGR, 0..255
                                        (ros/packages/caffe_ros/src/tensor_net.cpp:303-336 with the test node's defaults)
  tests/golden/trailnet/frames_rgb8.npz two of the images as 8-bit RGB camera frames [2,504,640,3] (what tests.cpp publishes)
  tests/golden/trailnet/expected.npz    `tests_cpp`: the predictions ros/packages/caffe_ros/tests/tests.cpp:64-69 expects (1e-3);
                                        `oracle_f64`: oracle/caffe.py in float64 on the same inputs
"""
import os
import shutil
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import caffe

REF = "/root/reference"
OUT = os.path.join(HERE, "trailnet")
os.makedirs(OUT, exist_ok=True)
import gzip
with open(os.path.join(REF, "models", "pretrained", "TrailNet_SResNet-18.prototxt"), "rb") as f:
    proto_bytes = f.read()
with gzip.GzipFile(os.path.join(OUT, "sresnet18_deploy.prototxt.gz"), "wb", mtime=0) as f:
    f.write(proto_bytes)
shutil.copyfile(os.path.join(REF, "models", "pretrained", "TrailNet_SResNet-18.caffemodel"), os.path.join(OUT, "sresnet18_weights.caffemodel"))
names = ["rot_l.jpg", "rot_c.jpg", "rot_r.jpg", "tran_l.jpg", "tran_r.jpg"]
tests_cpp = np.array([[0.932, 0.060, 0.006, 0.080, 0.848, 0.071],
                      [0.040, 0.958, 0.001, 0.488, 0.375, 0.135],
                      [0.000, 0.027, 0.971, 0.036, 0.407, 0.555],
                      [0.011, 0.988, 0.000, 0.981, 0.008, 0.009],
                      [0.000, 0.855, 0.144, 0.013, 0.031, 0.954]], np.float64)
x = np.stack([caffe.preprocess_bgr8(cv2.imread(os.path.join(REF, "ros/packages/caffe_ros/tests/data", n)), 320, 180) for n in names])
np.savez_compressed(os.path.join(OUT, "inputs.npz"), images=x.astype(np.float32), names=np.array(names))
# two camera frames as the test publishes them (tests.cpp:28-48: cv::imread, BGR -> RGB, encoding rgb8), for the drop-in run of the
# reference's unchanged tensor_net.cpp (tools/dropin/trailnet_driver.cpp): rows 0 and 4 of the tables above
frames = np.stack([cv2.cvtColor(cv2.imread(os.path.join(REF, "ros/packages/caffe_ros/tests/data", n)), cv2.COLOR_BGR2RGB) for n in (names[0], names[4])])
np.savez_compressed(os.path.join(OUT, "frames_rgb8.npz"), frames=frames, rows=np.array([0, 4]))
blobs = caffe.read_caffemodel(os.path.join(OUT, "sresnet18_weights.caffemodel"))
proto = proto_bytes.decode()
y = caffe.run_net(proto, blobs, x.astype(np.float64))
print("max |oracle - tests.cpp| =", np.abs(y - tests_cpp).max())
np.savez(os.path.join(OUT, "expected.npz"), tests_cpp=tests_cpp, oracle_f64=y)
