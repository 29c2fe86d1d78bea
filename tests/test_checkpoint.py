"""aldi_b200/checkpoint.py: the {model, ema} file layout and the load-from-EMA rule of aldi/checkpoint.py:19-31,
exercised on CPU through FlatLayout with a stand-in for the device step; the GPU round trip through B200TrainStep
is in tests/test_gpu_graph.py::test_checkpoint_resume_is_bit_identical."""
import os
import pickle

import pytest
import torch

from aldi_b200 import arch
from aldi_b200.checkpoint import DetectionCheckpointerWithEMA
from aldi_b200.detector import FlatLayout


class HostStep:
    """state_dict / load_state_dict over flat buffers exactly like B200TrainStep, without a GPU."""

    def __init__(self, sd_s, sd_t):
        self.layout = FlatLayout(8)
        self.flat = {"student": self.layout.pack_state_dict(sd_s), "teacher": self.layout.pack_state_dict(sd_t)}
        self.momentum = torch.zeros(self.layout.num_trainable)
        self.iter = 0

    def state_dict(self, which="student"):
        return self.layout.unpack_state_dict(self.flat[which])

    def load_state_dict(self, sd, which="student", strict=True):
        known = {key: (layer, field, off, n, shape) for (layer, field), (off, n, key, shape) in self.layout.entries.items()}
        missing = [k for k in known if k not in sd]
        unexpected = [k for k in sd if k not in known]
        for k, (layer, field, off, n, shape) in known.items():
            if k in sd:
                self.flat[which][off:off + n] = self.layout.to_internal(layer, field, sd[k].float())
        return missing, unexpected

    def optimizer_state(self):
        return {"momentum_buffer": self.momentum.clone(), "iteration": self.iter}

    def load_optimizer_state(self, st):
        self.momentum.copy_(st["momentum_buffer"])
        self.iter = st["iteration"]


def sds():
    return arch.synthetic_state_dict(seed=3), arch.synthetic_state_dict(seed=4)


def test_save_layout_matches_reference(tmp_path):
    s, t = sds()
    ck = DetectionCheckpointerWithEMA(HostStep(s, t), str(tmp_path))
    path = ck.save("model_0000099", iteration=99)
    data = torch.load(path, weights_only=False)
    assert set(data) >= {"model", "ema", "optimizer", "iteration"}
    assert all(k.startswith("model.") for k in data["ema"])          # EMA(nn.Module).model.* (aldi/ema.py:13)
    assert set(data["model"]) == set(s)
    assert torch.equal(data["model"]["backbone.fpn_lateral2.weight"], s["backbone.fpn_lateral2.weight"])
    assert torch.equal(data["ema"]["model.roi_heads.box_head.fc1.weight"], t["roi_heads.box_head.fc1.weight"])
    assert open(os.path.join(tmp_path, "last_checkpoint")).read() == "model_0000099.pth"


def test_fresh_start_loads_ema_weights_into_the_model(tmp_path):
    """aldi/checkpoint.py:19-31: not resuming from a .pth that has an "ema" entry -> the model starts from the EMA."""
    s, t = sds()
    path = DetectionCheckpointerWithEMA(HostStep(s, t), str(tmp_path / "burnin")).save("model_final")
    other = arch.synthetic_state_dict(seed=9)
    step = HostStep(other, other)
    DetectionCheckpointerWithEMA(step, str(tmp_path / "run")).resume_or_load(path, resume=False)
    got = step.state_dict("student")
    assert all(torch.equal(got[k], t[k]) for k in t)                  # EMA weights, not the "model" entry
    assert all(torch.equal(step.state_dict("teacher")[k], other[k]) for k in t)   # teacher untouched (trainer re-inits it)


def test_resume_restores_model_ema_and_optimizer(tmp_path):
    s, t = sds()
    src = HostStep(s, t)
    src.momentum.fill_(0.25)
    src.iter = 7
    DetectionCheckpointerWithEMA(src, str(tmp_path)).save("model_0000006")
    other = arch.synthetic_state_dict(seed=9)
    step = HostStep(other, other)
    DetectionCheckpointerWithEMA(step, str(tmp_path)).resume_or_load("", resume=True)
    assert all(torch.equal(step.state_dict("student")[k], s[k]) for k in s)
    assert all(torch.equal(step.state_dict("teacher")[k], t[k]) for k in t)
    assert float(step.momentum[0]) == 0.25 and step.iter == 7


def test_pkl_with_detectron2_names_and_missing_keys(tmp_path, caplog):
    s, _ = sds()
    partial = {k: v.numpy() for k, v in s.items() if not k.startswith("roi_heads.box_predictor")}
    p = tmp_path / "zoo.pkl"
    with open(p, "wb") as fh:
        pickle.dump({"model": partial, "__author__": "test", "matching_heuristics": False}, fh)
    other = arch.synthetic_state_dict(seed=9)
    step = HostStep(other, other)
    with caplog.at_level("WARNING", logger="aldi_b200"):
        DetectionCheckpointerWithEMA(step).resume_or_load(str(p), resume=False)
    got = step.state_dict("student")
    assert torch.equal(got["backbone.bottom_up.res2.0.conv1.weight"], s["backbone.bottom_up.res2.0.conv1.weight"])
    assert torch.equal(got["roi_heads.box_predictor.cls_score.weight"], other["roi_heads.box_predictor.cls_score.weight"])
    assert "not found in the checkpoint" in caplog.text


def test_missing_file_raises(tmp_path):
    s, t = sds()
    with pytest.raises(AssertionError):
        DetectionCheckpointerWithEMA(HostStep(s, t)).resume_or_load(str(tmp_path / "nope.pth"), resume=False)
