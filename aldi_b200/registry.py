"""Plug-in registries keyed by cfg strings — the reference's plugin API for this path
(detectron2 Registry; aldi/align.py:11, aldi/distill.py:17,33; aldi/model.py:14-16)."""


class Registry:
    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, "An object named '{}' was already registered in '{}' registry!".format(
            name, self._name)
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._do_register(o.__name__, o)
                return o
            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return ret

    def __contains__(self, name):
        return name in self._obj_map


META_ARCH_REGISTRY = Registry("META_ARCH")
ALIGN_MIXIN_REGISTRY = Registry("ALIGN_MIXIN")
DISTILL_MIXIN_REGISTRY = Registry("DISTILL_MIXIN")
DISTILLER_REGISTRY = Registry("DISTILLER")
