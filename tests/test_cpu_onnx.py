"""ONNX weight ingestion: the reference loads its model by ``.onnx`` path
(``_script/gpu_handler.py:61-65``, ``simple_detector.py:39-46``).  No real checkpoint exists here
(``.MISSING_LARGE_BLOBS``), so the files are written by ``onnx_reader.write_conv_onnx`` in the layout of
an Ultralytics export and read back with the package-free protobuf walker."""
import numpy as np
import pytest

from aerial_image_recognition_b200 import graph as G, onnx_reader as R, session as S, weights as W


@pytest.fixture(scope="module")
def v8():
    g = G.build("yolov8m", imgsz=64)
    return g, W.make_synthetic_weights(g, 5, calibrate=False)


@pytest.mark.parametrize("named,half", [(True, False), (False, False), (True, True)])
def test_roundtrip_by_name_by_order_and_fp16(tmp_path, v8, named, half):
    g, w = v8
    path = str(tmp_path / "yolov8_tokyo_checkpoint.onnx")
    R.write_conv_onnx(path, g, w, named=named, half=half)
    got = R.load_onnx_weights(path, g)
    assert set(got) == set(w)
    for k in w:
        ref = w[k].astype(np.float16).astype(np.float32) if half else w[k]
        assert got[k].dtype == np.float32 and np.array_equal(got[k], ref), k


def test_nodes_and_initializers_are_parsed(tmp_path, v8):
    g, w = v8
    path = str(tmp_path / "m.onnx")
    R.write_conv_onnx(path, g, w)
    nodes, inits = R.read_onnx(path)
    convs = [n for n in nodes if n.op_type == "Conv"]
    assert len(convs) == 89 + 1                                   # 83 dense + 6 depthwise + the DFL arange conv
    assert convs[0].name == "/model/0/conv/Conv" and convs[0].inputs[1] == "model.0.conv.weight"
    assert inits["model.0.conv.weight"].shape == (48, 3, 3, 3)
    assert inits["model.22.cv3.0.0.0.conv.weight"].shape == (192, 1, 3, 3)       # depthwise cls branch (Ultralytics 8.3.x)
    assert np.array_equal(inits["model.22.dfl.conv.weight"].ravel(), np.arange(16, dtype=np.float32))


def test_wrong_architecture_fails_loudly(tmp_path, v8):
    g, w = v8
    path = str(tmp_path / "m.onnx")
    R.write_conv_onnx(path, g, w, named=False)
    with pytest.raises(ValueError, match="Conv nodes|expects"):
        R.load_onnx_weights(path, G.build("yolov7", imgsz=64))
    w2 = dict(w)
    w2["model.1.weight"] = w["model.1.weight"][:, :, :1, :1]
    R.write_conv_onnx(path, g, w2)
    with pytest.raises(ValueError, match="expects"):
        R.load_onnx_weights(path, g)


def test_session_load_weights_dispatches_on_extension(tmp_path, v8):
    g, w = v8
    path = str(tmp_path / "yolov8_tokyo_checkpoint.onnx")
    R.write_conv_onnx(path, g, w)
    got = S.load_weights(path)
    assert np.array_equal(got["model.21.cv2.weight"], w["model.21.cv2.weight"])
    np.savez(str(tmp_path / "m.npz"), **w)
    assert np.array_equal(S.load_weights(str(tmp_path / "m.npz"))["model.0.bias"], w["model.0.bias"])
    with pytest.raises(FileNotFoundError, match="synthetic"):          # as ort.InferenceSession(model_path) does; no silent random weights
        S.load_weights(str(tmp_path / "absent.onnx"))
    with pytest.raises(FileNotFoundError):
        S.resolve_weights(str(tmp_path / "absent.onnx"), "yolov8m", None)
    assert S.resolve_weights(str(tmp_path / "absent.onnx"), "yolov8m", "synthetic") is None      # the explicit opt-in
    assert S.resolve_weights(None, "yolov8m", w) is w
    (tmp_path / "m.bin").write_bytes(b"x")
    with pytest.raises(ValueError):
        S.load_weights(str(tmp_path / "m.bin"))


# ---- a real exporter's file (torch's TorchScript ONNX exporter, the one Ultralytics' export drives) ---------------------
@pytest.fixture(scope="module")
def exported(tmp_path_factory, v8):
    import _export_onnx as E
    g, w = v8
    m = E.YoloV8m()
    E.load_deploy_weights(m, w)
    d = tmp_path_factory.mktemp("onnx")
    E.export(m, str(d / "yolov8_tokyo_checkpoint.onnx"))
    return g, w, m, d


def test_reader_on_a_torch_exported_yolov8m(exported):
    """Conv -> Sigmoid -> Mul, Slice (chunk), Concat, MaxPool, Resize, Softmax, the DFL arange conv: the graph a real export of
    the architecture contains.  Weights come back bit-identical by initializer name and, with the names ignored, by order."""
    g, w, m, d = exported
    path = str(d / "yolov8_tokyo_checkpoint.onnx")
    nodes, inits = R.read_onnx(path)
    kinds = {nd.op_type for nd in nodes}
    assert {"Conv", "Sigmoid", "Mul", "Concat", "MaxPool", "Resize", "Softmax"} <= kinds and "BatchNormalization" not in kinds
    assert sum(nd.op_type == "Conv" for nd in nodes) == 90          # 89 layers + the DFL arange conv
    assert "model.22.cv3.0.0.0.conv.weight" in inits and "model.22.dfl.conv.weight" in inits
    for by_name in (True, False):
        got = R.load_onnx_weights(path, g, by_name=by_name)
        assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w), by_name
    assert S.load_weights(path)["model.9.cv2.bias"].dtype == np.float32


def test_reader_on_a_torch_exported_fp16_model(exported, v8):
    import copy
    import _export_onnx as E
    g, w, m, d = exported
    path = str(d / "half.onnx")
    E.export(copy.deepcopy(m), path, half=True)
    _nodes, inits = R.read_onnx(path)
    assert inits["model.0.conv.weight"].dtype == np.float16
    got = R.load_onnx_weights(path, g)
    assert all(got[k].dtype == np.float32 and np.array_equal(got[k], w[k].astype(np.float16).astype(np.float32)) for k in w)


def test_reader_rejects_unfolded_batchnorm_and_wrong_architecture(exported, tmp_path):
    import torch
    import _export_onnx as E
    g, w, m, d = exported

    class TwoLayers(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.ModuleList([E.Conv(3, 48, 3, 2, bn=True), E.Conv(48, 96, 3, 2, bn=True)])

        def forward(self, x):
            return self.model[1](self.model[0](x))
    path = str(tmp_path / "bn.onnx")
    E.export(TwoLayers(), path, train_mode=True)                  # training-mode export keeps the BatchNormalization nodes
    assert any(nd.op_type == "BatchNormalization" for nd in R.read_onnx(path)[0])
    with pytest.raises(ValueError, match="BatchNormalization"):
        R.load_onnx_weights(path, g)
    # same layer names, different stride: a by-name match must still be refused
    class WrongStride(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = torch.nn.ModuleList([E.Conv(3, 48, 3, 1)])

        def forward(self, x):
            return self.model[0](x)
    path = str(tmp_path / "stride.onnx")
    E.export(WrongStride(), path)
    with pytest.raises(ValueError, match="strides|convolutions found"):
        R.load_onnx_weights(path, g)
    with pytest.raises(ValueError):                                # a YOLOv8m file is not a YOLOv7 model
        R.load_onnx_weights(str(d / "yolov8_tokyo_checkpoint.onnx"), G.build("yolov7"))


def test_reader_rejects_external_data_initializers(tmp_path):
    # TensorProto with data_location = EXTERNAL (field 14 = 1): weights outside the file
    t = R._enc_varint((1 << 3) | 0) + R._enc_varint(4) + R._enc_varint((2 << 3) | 0) + R._enc_varint(1) + R._enc_field(8, b"w")
    t += R._enc_varint((14 << 3) | 0) + R._enc_varint(1)
    graph = R._enc_field(5, t)
    (tmp_path / "ext.onnx").write_bytes(R._enc_varint((1 << 3) | 0) + R._enc_varint(8) + R._enc_field(7, graph))
    with pytest.raises(ValueError, match="external data"):
        R.read_onnx(str(tmp_path / "ext.onnx"))
