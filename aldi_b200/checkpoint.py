"""Checkpoint I/O on the flat parameter buffers, in the file layout the reference reads and writes.

Mirror of `aldi.checkpoint.DetectionCheckpointerWithEMA` (aldi/checkpoint.py:8-31) over Detectron2's
`DetectionCheckpointer` / fvcore `Checkpointer` for the path this repository covers:

  * `save(name)` writes `<save_dir>/<name>.pth` = {"model": <student state_dict>, "ema": {"model.<key>": tensor},
    "optimizer": ..., "iteration": n} with Detectron2 key names, plus the `last_checkpoint` tag file — the layout
    `tools/visualize_featurespace.py:58-63` and the reference's own resume path read;
  * `resume_or_load(path, resume)`: resume -> the last checkpoint of `save_dir` incl. EMA / optimizer state;
    otherwise load `path`, and — the reason the reference subclasses the checkpointer — if that file is a `.pth`
    that carries an "ema" entry, start the model FROM THE EMA WEIGHTS (keys with the leading "model." removed,
    strict=False, incompatible keys logged);
  * `.pkl` files in Detectron2's model-zoo layout ({"model": {key: ndarray}, "matching_heuristics": ...}) load when
    their keys are already Detectron2 names; Caffe2-name matching heuristics are out of scope (they need the zoo).

The state dicts are views of `B200TrainStep`'s flat buffers via FlatLayout (detector.py), so released burn-in
checkpoints with these keys load unchanged and files written here load in the reference.
"""
import logging
import os
import pickle

import torch

logger = logging.getLogger("aldi_b200")


class DetectionCheckpointerWithEMA:
    def __init__(self, step, save_dir="", *, save_to_disk=True):
        self.step = step            # B200TrainStep (or anything with state_dict/load_state_dict(which=...))
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk

    # ---- fvcore Checkpointer.save ------------------------------------------------------------------------
    def save(self, name, **extra):
        if not self.save_dir or not self.save_to_disk:
            return None
        data = {"model": self.step.state_dict("student"),
                "ema": {"model." + k: v for k, v in self.step.state_dict("teacher").items()}}
        if hasattr(self.step, "optimizer_state"):
            data["optimizer"] = self.step.optimizer_state()
        data.update(extra)
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, "%s.pth" % name)
        torch.save(data, path)
        with open(os.path.join(self.save_dir, "last_checkpoint"), "w") as fh:
            fh.write(os.path.basename(path))
        return path

    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, "last_checkpoint"))

    def get_checkpoint_file(self):
        with open(os.path.join(self.save_dir, "last_checkpoint")) as fh:
            return os.path.join(self.save_dir, fh.read().strip())

    # ---- file -> dict -------------------------------------------------------------------------------------
    @staticmethod
    def _load_file(path):
        if path.endswith(".pkl"):
            with open(path, "rb") as fh:
                data = pickle.load(fh, encoding="latin1")
            if "model" in data and "__author__" in data:
                model = {k: torch.as_tensor(v) for k, v in data["model"].items() if not k.endswith("_momentum")}
                return {"model": model, "__author__": data["__author__"], "matching_heuristics": data.get("matching_heuristics", False)}
            if "blobs" in data:
                data = data["blobs"]
            return {"model": {k: torch.as_tensor(v) for k, v in data.items() if not k.endswith("_momentum")},
                    "__author__": "Caffe2", "matching_heuristics": True}
        data = torch.load(path, map_location="cpu", weights_only=False)
        if "model" not in data:
            data = {"model": data}
        return data

    def _log_incompatible(self, missing, unexpected):
        if missing:
            logger.warning("Some model parameters or buffers are not found in the checkpoint: %s", sorted(missing)[:20])
        if unexpected:
            logger.warning("The checkpoint state_dict contains keys that are not used by the model: %s", sorted(unexpected)[:20])

    # ---- fvcore Checkpointer.load + the EMA override of aldi/checkpoint.py:19-31 ------------------------------------
    def load(self, path, with_state=True):
        if not path:
            logger.info("No checkpoint found. Initializing model from scratch")
            return {}
        if not os.path.isfile(path):
            raise AssertionError("Checkpoint %s not found!" % path)
        data = self._load_file(path)
        model = {k[len("module."):] if k.startswith("module.") else k: v for k, v in data.pop("model").items()}
        if data.get("matching_heuristics") and not any(k.startswith("backbone.") for k in model):
            raise NotImplementedError("%s uses Caffe2 / torchvision parameter names; the name-matching heuristics of "
                                      "detectron2.checkpoint.c2_model_loading are not rebuilt here" % path)
        self._log_incompatible(*self.step.load_state_dict(model, which="student", strict=False))
        if with_state:
            if "ema" in data:
                ema = {k.replace("model.", "", 1): v for k, v in data["ema"].items()}
                self._log_incompatible(*self.step.load_state_dict(ema, which="teacher", strict=False))
            if "optimizer" in data and hasattr(self.step, "load_optimizer_state"):
                self.step.load_optimizer_state(data["optimizer"])
        return data

    def resume_or_load(self, path, *, resume=True):
        if resume and self.has_checkpoint():
            return self.load(self.get_checkpoint_file(), with_state=True)
        ret = self.load(path, with_state=False)
        if (not resume) and path and path.endswith(".pth") and "ema" in ret:
            logger.info("Loading EMA weights as model starting point.")
            ema = {k.replace("model.", "", 1): v for k, v in ret["ema"].items()}
            self._log_incompatible(*self.step.load_state_dict(ema, which="student", strict=False))
        return ret
