"""Golden vectors for the strong-augmentation transforms from the REFERENCE's own classes (authoring container only).

aldi/aug.py is executed with minimal stand-ins for the Detectron2 / fvcore names it imports (Transform base class,
NoOpTransform, the `T` namespace) — the three transforms pinned here (RandomBlurTransform, RandomEraseTransform,
MICTransform, aldi/aug.py:81-186) only use numpy, scipy and cv2, which are installed.  Inputs are regenerated from
seeds by `aug_case`; the outputs are stored as uint8 arrays on small images.

    python tests/golden/make_aug_golden.py     ->  tests/golden/aug_golden.pt
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/aldi/aug.py"

CASES = {"a": (11, 96, 160), "b": (12, 75, 101), "c": (13, 128, 64)}


def aug_case(name):
    seed, h, w = CASES[name]
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    img[h // 4:h // 2, w // 4:w // 2] = rng.randint(0, 256, size=3)      # a flat patch: blur edges, erase borders
    return img


def load_reference():
    class Transform:
        def _set_attributes(self, params=None):
            if params:
                for k, v in params.items():
                    if k != "self" and not k.startswith("_"):
                        setattr(self, k, v)

    class NoOpTransform(Transform):
        pass

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Aug:
        def _init(self, params=None):
            pass

    mod("detectron2"); mod("detectron2.data")
    mod("detectron2.data.transforms", Augmentation=Aug, RandomApply=object)
    mod("detectron2.data.transforms.augmentation", _get_aug_input_args=None)
    mod("detectron2.data.transforms.augmentation_impl", RandomApply=object)
    mod("detectron2.data.detection_utils")
    sys.modules["detectron2.data"].transforms = sys.modules["detectron2.data.transforms"]
    sys.modules["detectron2.data"].detection_utils = sys.modules["detectron2.data.detection_utils"]
    mod("fvcore"); mod("fvcore.transforms")
    mod("fvcore.transforms.transform", Transform=Transform, NoOpTransform=NoOpTransform)
    ns = {"__name__": "ref_aldi_aug"}
    exec(compile(open(REF).read(), REF, "exec"), ns)
    return ns


def main():
    ref = load_reference()
    out = {}
    for name in CASES:
        img = aug_case(name)
        seed = CASES[name][0]
        g = {}
        random.seed(seed); np.random.seed(seed)
        g["blur"] = ref["RandomBlurTransform"]((0.1, 2.0)).apply_image(img.copy())
        random.seed(seed + 100); np.random.seed(seed + 100)
        g["erase"] = ref["RandomEraseTransform"](sl=0.05, sh=0.2, r1=0.3, r2=3.3, value="random").apply_image(img.copy())
        random.seed(seed + 200); np.random.seed(seed + 200)
        g["erase_thin"] = ref["RandomEraseTransform"](sl=0.02, sh=0.2, r1=0.05, r2=8, value="random").apply_image(img.copy())
        random.seed(seed + 300); np.random.seed(seed + 300)
        g["mic"] = ref["MICTransform"](0.5, 32).apply_image(img.copy())
        random.seed(seed + 400); np.random.seed(seed + 400)
        g["mic16"] = ref["MICTransform"](0.3, 16).apply_image(img.copy())
        out[name] = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g.items()}
        print(name, {k: tuple(v.shape) for k, v in g.items()})
    torch.save(out, os.path.join(HERE, "aug_golden.pt"))


if __name__ == "__main__":
    main()
