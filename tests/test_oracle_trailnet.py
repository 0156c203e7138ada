"""CPU: the Caffe oracle (oracle/caffe.py) pinned to the reference's own TrailNet test
(ros/packages/caffe_ros/tests/tests.cpp:50-106: five images -> six softmax outputs, EXPECT_NEAR 1e-3)."""
import os

import numpy as np
import pytest

from oracle import caffe
from tests.util import trailnet_model_files

TN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trailnet")
PROTO, MODEL = trailnet_model_files()


@pytest.fixture(scope="module")
def model():
    return open(PROTO).read(), caffe.read_caffemodel(MODEL)


def test_caffemodel_reader_shapes(model):
    proto, blobs = model
    assert blobs["conv1"][0].shape == (64, 3, 7, 7) and blobs["conv1"][1].shape == (64,)
    assert blobs["fc3"][0].shape == (3, 16384) and blobs["fc3_t"][0].shape == (3, 16384)
    assert blobs["res4_1_proj"][0].shape[:2] == (512, 256) if "res4_1_proj" in blobs else True
    # S-ReLU = Scale(+1) -> ReLU -> Scale(-1): the stored scale blobs are (1, +-1) per channel (prototxt:54-105)
    np.testing.assert_array_equal(blobs["conv1_srelu1_1"][0], 1.0)
    np.testing.assert_array_equal(blobs["conv1_srelu1_1"][1], 1.0)
    np.testing.assert_array_equal(blobs["conv1_srelu1_3"][1], -1.0)
    net = caffe.parse_prototxt(proto)
    assert len(net["layer"]) == 87 and net["input_shape"][0]["dim"] == [1, 3, 180, 320]


def test_trailnet_reference_predictions(model):
    """The reference's expected predictions for its five test images, its tolerance."""
    proto, blobs = model
    x = np.load(os.path.join(TN, "inputs.npz"))["images"]
    exp = np.load(os.path.join(TN, "expected.npz"))
    y = caffe.run_net(proto, blobs, x.astype(np.float64))
    assert y.shape == (5, 6)
    np.testing.assert_allclose(y, exp["tests_cpp"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(y, exp["oracle_f64"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(y[:, :3].sum(axis=1), 1.0, atol=1e-9)      # two 3-way softmax heads
    np.testing.assert_allclose(y[:, 3:].sum(axis=1), 1.0, atol=1e-9)


def test_caffe_ceil_mode_pooling():
    """Caffe's pooling output size is ceil((in - k) / stride) + 1 (87x157 -> 43x78 after conv1), windows clipped at the border."""
    x = np.arange(2 * 7 * 8, dtype=np.float64).reshape(1, 2, 7, 8)
    y = caffe._pool(x, "MAX", 3, 2, 0)
    assert y.shape == (1, 2, 3, 4)                       # ceil((7-3)/2)+1 = 3, ceil((8-3)/2)+1 = 4 (last window has 2 columns)
    assert y[0, 0, 2, 3] == x[0, 0, 4:7, 6:8].max()
    a = caffe._pool(x, "AVE", 3, 1, 0)
    np.testing.assert_allclose(a[0, 1, 0, 0], x[0, 1, 0:3, 0:3].mean())


def test_cpp_caffe_parser_plan_matches_oracle(model):
    """The C++ side on the CPU: nvcaffeparser1 (redtail_b200/csrc/host/caffe_parser.cpp) reads the same prototxt + caffemodel into an
    INetworkDefinition, the host-only dump writes its plan, and the graph-level oracle (oracle/plan.py) executes that plan:
    same numbers as the Python Caffe oracle, i.e. the parser wires the reference's model correctly (no GPU involved)."""
    import ctypes as C
    from redtail_b200._lib import engine_lib
    from oracle import plan as P
    lib = engine_lib()
    args = (PROTO.encode(), MODEL.encode(), b"out", 1)
    n = lib.rt_caffe_dump_plan(*args, None, 0)
    assert n > 40e6, lib.rt_stereo_last_error()
    buf = C.create_string_buffer(n)
    assert lib.rt_caffe_dump_plan(*args, buf, n) == n
    pl = P.parse(buf.raw)
    assert len(pl["layers"]) == 87 and pl["inputs"][0][1:] == ("data", (3, 180, 320))
    kinds = [L["kind"] for L in pl["layers"]]
    assert kinds.count(P.K_POOLING) == 2 and kinds.count(P.K_FC) == 2 and kinds.count(P.K_SOFTMAX) == 2
    pools = [L for L in pl["layers"] if L["kind"] == P.K_POOLING]
    assert (pools[0]["oh"], pools[0]["ow"]) == (43, 78) and (pools[1]["oh"], pools[1]["ow"]) == (4, 8)      # Caffe's ceil-mode extents
    x = np.load(os.path.join(TN, "inputs.npz"))["images"]
    exp = np.load(os.path.join(TN, "expected.npz"))
    (y,) = P.execute(pl, {"data": x.astype(np.float64)}).values()
    np.testing.assert_allclose(y.reshape(5, 6), exp["oracle_f64"], rtol=0, atol=1e-12)
    assert lib.rt_caffe_dump_plan(b"/nonexistent.prototxt", args[1], b"out", 1, None, 0) == 0          # loud failure, no crash


def test_cpp_caffe_parser_rejects_malformed_models(tmp_path):
    """Truncated / corrupted model files and unsupported layers fail loudly (return value 0 + message), they do not crash."""
    from redtail_b200._lib import engine_lib
    lib = engine_lib()
    proto = open(PROTO).read()
    model = open(MODEL, "rb").read()
    good_p, good_m = PROTO.encode(), MODEL.encode()

    def dump(p, m, blob=b"out"):
        return lib.rt_caffe_dump_plan(p, m, blob, 1, None, 0)

    cases = {"trunc.caffemodel": model[:len(model) // 3], "garbage.caffemodel": bytes(range(256)) * 64, "empty.caffemodel": b""}
    for name, data in cases.items():
        f = tmp_path / name
        f.write_bytes(data)
        assert dump(good_p, str(f).encode()) == 0, name
    bad = {"unbalanced.prototxt": proto[:len(proto) // 2],                                   # ends inside a layer { ... }
           "lrn.prototxt": proto.replace('type: "Softmax"', 'type: "LRN"', 1),                # a layer type the parser does not map
           "nodim.prototxt": proto.replace("input_shape", "input_shapeX", 1),
           "badbottom.prototxt": proto.replace('bottom: "pool1"', 'bottom: "no_such_blob"', 1)}
    for name, text in bad.items():
        f = tmp_path / name
        f.write_text(text)
        assert dump(str(f).encode(), good_m) == 0, name
    assert dump(good_p, good_m, b"no_such_output") == 0
    assert dump(good_p, good_m) > 0
