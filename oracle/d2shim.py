"""ORACLE support (test infrastructure): a stand-in `detectron2` / `fvcore` package tree so that the REAL
reference modules under /root/reference/aldi can be imported in the authoring container (Detectron2 is
not installable here, SURVEY.md "three facts" #2).

Names the ALDI hot path actually calls are bound to the restatement in oracle/d2_rcnn.py; every other
attribute of every `detectron2.*`, `fvcore.*`, `timm.*`, `pycocotools.*`, `iopath.*`, `yacs.*` module
resolves to an inert mock class (usable as a base class / decorator), which is enough for the import-time
side of files like aldi/trainer.py and aldi/backbone.py.

Used only by tests/golden/make_golden.py to generate golden vectors from the reference's own code.
"""
import functools
import importlib.abc
import importlib.machinery
import inspect
import sys
import types

from . import d2_rcnn as d2

_MOCK_ROOTS = ("detectron2", "fvcore", "timm", "pycocotools", "iopath", "yacs", "cv2")


class _MockMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_mock(name)


class _MockBase(metaclass=_MockMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # decorator use
        return _MockBase()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _MockBase()


def _make_mock(name):
    return _MockMeta(name, (_MockBase,), {})


class _MockModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        m = _make_mock(name)
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _MOCK_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _MockModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ---- the pieces of detectron2 the hot path really uses ------------------------------------------
class CfgNode(dict):
    """Attribute-access dict (yacs.CfgNode surface needed by aldi/config.py and from_config)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.pop("cfg", None), CfgNode)


def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable for __init__ methods."""
    assert init_func is not None and inspect.isfunction(init_func) and init_func.__name__ == "__init__"

    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        if _called_with_cfg(*args, **kwargs):
            explicit = type(self).from_config(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)

    return wrapped


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(fn):
                self._obj_map[fn.__name__] = fn
                return fn
            return deco
        self._obj_map[obj.__name__] = obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return self._obj_map[name]


META_ARCH_REGISTRY = Registry("META_ARCH")


@META_ARCH_REGISTRY.register()
class GeneralizedRCNN(d2.GeneralizedRCNN):
    """oracle/d2_rcnn.GeneralizedRCNN behind detectron2's @configurable / from_config protocol."""

    @configurable
    def __init__(self, *, num_classes=8, freeze_at=2, **kw):
        super().__init__(num_classes=num_classes, freeze_at=freeze_at)

    @classmethod
    def from_config(cls, cfg):
        return {"num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES, "freeze_at": cfg.MODEL.BACKBONE.FREEZE_AT}


class _Comm(types.ModuleType):
    @staticmethod
    def get_world_size():
        return 1

    @staticmethod
    def is_main_process():
        return True


def get_cfg():
    """The subset of detectron2 defaults that build_aldi / from_config read."""
    c = CfgNode()
    c.MODEL = CfgNode(META_ARCHITECTURE="GeneralizedRCNN", DEVICE="cpu",
                      BACKBONE=CfgNode(FREEZE_AT=2, NAME="build_resnet_fpn_backbone"),
                      ROI_HEADS=CfgNode(NUM_CLASSES=8))
    c.DATASETS = CfgNode()
    c.SOLVER = CfgNode()
    return c


def install(reference_root="/root/reference"):
    """Put the stand-in packages on sys.modules / sys.meta_path and the reference on sys.path."""
    if any(isinstance(f, _Finder) for f in sys.meta_path):
        return
    sys.meta_path.insert(0, _Finder())
    import importlib

    def mod(name):
        return importlib.import_module(name)

    mod("detectron2.config").configurable = configurable
    mod("detectron2.config").CfgNode = CfgNode
    mod("detectron2.config").get_cfg = get_cfg
    mod("detectron2.utils.registry").Registry = Registry
    mod("detectron2.modeling").GeneralizedRCNN = GeneralizedRCNN
    mod("detectron2.modeling.meta_arch.rcnn").GeneralizedRCNN = GeneralizedRCNN
    mod("detectron2.modeling.meta_arch.build").META_ARCH_REGISTRY = META_ARCH_REGISTRY
    mod("detectron2.layers").cat = d2.cat
    mod("detectron2.layers.wrappers").cross_entropy = d2.cross_entropy
    mod("detectron2.modeling.sampling").subsample_labels = d2.subsample_labels
    mod("detectron2.modeling.box_regression")._dense_box_regression_loss = d2._dense_box_regression_loss
    mod("detectron2.structures").Boxes = d2.Boxes
    mod("detectron2.structures").Instances = d2.Instances
    mod("detectron2.structures.boxes").Boxes = d2.Boxes
    mod("detectron2.structures.instances").Instances = d2.Instances
    mod("detectron2.utils.events").get_event_storage = d2.get_event_storage
    mod("detectron2.utils.events").EventStorage = d2.EventStorage
    mod("detectron2.utils.logger")._log_api_usage = lambda *a, **k: None
    comm = _Comm("detectron2.utils.comm")
    sys.modules["detectron2.utils.comm"] = comm
    mod("detectron2.utils").comm = comm
    mod("fvcore.nn").smooth_l1_loss = d2.smooth_l1_loss
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
