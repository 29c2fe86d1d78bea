"""The reference's plug-in surface: registries keyed by cfg strings, `build_aldi`, `build_distiller` /
`Distiller.from_config` / `distill_enabled` -- on CPU, without an engine.  Where the reference tree is present
(authoring container) the distiller selection is ALSO held against the reference's own classes
(aldi/distill.py:17-85 imported through oracle/d2shim.py)."""
import itertools
import os
import types

import pytest

from aldi_b200 import model as M
from aldi_b200.config import add_aldi_config, get_cfg
from aldi_b200.registry import ALIGN_MIXIN_REGISTRY, DISTILL_MIXIN_REGISTRY, DISTILLER_REGISTRY, META_ARCH_REGISTRY
from aldi_b200.train_step import StepConfig

FLAGS = ("HARD_ROIH_CLS_ENABLED", "HARD_ROIH_REG_ENABLED", "HARD_OBJ_ENABLED", "HARD_RPN_REG_ENABLED", "ROIH_CLS_ENABLED",
         "ROIH_REG_ENABLED", "OBJ_ENABLED", "RPN_REG_ENABLED")


def _cfg(name="ALDIDistiller", **flags):
    cfg = get_cfg()
    add_aldi_config(cfg)
    cfg.DOMAIN_ADAPT.DISTILL.DISTILLER_NAME = name
    for k, v in flags.items():
        cfg.DOMAIN_ADAPT.DISTILL[k] = v
    return cfg


def test_registries_hold_the_reference_names():
    assert "GeneralizedRCNN" in META_ARCH_REGISTRY                       # cfg.MODEL.META_ARCHITECTURE
    assert "AlignMixin" in ALIGN_MIXIN_REGISTRY                          # cfg.DOMAIN_ADAPT.ALIGN.MIXIN_NAME default
    assert "DistillMixin" in DISTILL_MIXIN_REGISTRY                      # cfg.DOMAIN_ADAPT.DISTILL.MIXIN_NAME default
    for name in ("Distiller", "HardDistiller", "ALDIDistiller"):         # aldi/distill.py:44,60,87
        assert name in DISTILLER_REGISTRY
    cfg = _cfg()
    assert cfg.DOMAIN_ADAPT.ALIGN.MIXIN_NAME in ALIGN_MIXIN_REGISTRY and cfg.DOMAIN_ADAPT.DISTILL.MIXIN_NAME in DISTILL_MIXIN_REGISTRY
    with pytest.raises(KeyError, match="No object named 'Nope' found in 'DISTILLER' registry"):
        DISTILLER_REGISTRY.get("Nope")


def test_build_aldi_composes_the_class_from_the_three_registries(monkeypatch):
    """aldi/model.py:12-34 without a device: the engine constructor is stubbed, the class composition is real."""
    made = {}

    def fake_init(self, cfg=None, **kw):
        made["cfg"], made["kw"] = cfg, kw
        self.engine, self.which, self.training = types.SimpleNamespace(cfg=StepConfig()), "student", True

    monkeypatch.setattr(M.GeneralizedRCNN, "__init__", fake_init)
    cfg = _cfg()
    model = M.build_aldi(cfg, state_dict={"x": 1})
    mro = [c.__name__ for c in type(model).__mro__]
    assert mro[:4] == ["ALDI", "AlignMixin", "DistillMixin", "GeneralizedRCNN"], mro
    assert made["cfg"] is cfg and made["kw"] == {"state_dict": {"x": 1}}
    cfg.MODEL.META_ARCHITECTURE = "DeformableDETR"
    with pytest.raises(KeyError, match="META_ARCH"):
        M.build_aldi(cfg)


def _fake_pair():
    eng = types.SimpleNamespace(cfg=StepConfig())
    return types.SimpleNamespace(engine=eng), types.SimpleNamespace(engine=eng), eng


@pytest.mark.parametrize("name", ["Distiller", "HardDistiller", "ALDIDistiller"])
def test_build_distiller_selects_by_cfg_string_and_sets_the_engine_flags(name):
    teacher, student, eng = _fake_pair()
    cfg = _cfg(name, HARD_OBJ_ENABLED=True, ROIH_CLS_ENABLED=True, CLS_TMP=2.0)
    cfg.DOMAIN_ADAPT.TEACHER.THRESHOLD = 0.7
    d = M.build_distiller(cfg, teacher, student)
    assert type(d).__name__ == name
    if name == "Distiller":
        assert not d.distill_enabled() and d(None, None) == {}
        return
    assert d.distill_enabled() and d.pseudo_label_threshold == 0.7
    if name == "HardDistiller":
        # the student's standard losses come back unmasked (aldi/distill.py:78-81); no soft losses exist
        assert (eng.cfg.do_hard_cls, eng.cfg.do_hard_obj, eng.cfg.do_hard_rpn_reg, eng.cfg.do_hard_roi_reg) == (True,) * 4
        assert not any((eng.cfg.do_cls_dst, eng.cfg.do_obj_dst, eng.cfg.do_rpn_reg_dst, eng.cfg.do_roih_reg_dst))
    else:
        assert eng.cfg.do_hard_obj and eng.cfg.do_cls_dst and not eng.cfg.do_hard_cls and not eng.cfg.do_obj_dst
        assert d.cls_temperature == 2.0
    assert eng.cfg.distill_enabled


def test_hard_distiller_ignores_the_soft_flags():
    """configs/Base-DETR.yaml:76-81 + ALDI-Best-DETR-Cityscapes.yaml:10-13: HardDistiller with only soft
    flags set never distils."""
    teacher, student, eng = _fake_pair()
    d = M.build_distiller(_cfg("HardDistiller", ROIH_CLS_ENABLED=True, OBJ_ENABLED=True), teacher, student)
    assert not d.distill_enabled() and not eng.cfg.distill_enabled


@pytest.mark.skipif(not os.path.isdir("/root/reference/aldi"), reason="reference tree not present (GPU box)")
def test_distiller_selection_agrees_with_the_reference_classes():
    import make_golden as mg                       # installs oracle/d2shim and imports the REAL aldi.* modules
    import aldi.distill as ref
    import cases
    sd_s, sd_t = cases.student_teacher_state(next(iter(cases.CASES.values())))
    student, rcfg = mg.build_reference_model(sd_s)
    teacher, _ = mg.build_reference_model(sd_t)
    for name in ("Distiller", "HardDistiller", "ALDIDistiller"):
        for on in itertools.chain([()], [(f,) for f in FLAGS], [("HARD_OBJ_ENABLED", "OBJ_ENABLED")]):
            flags = {f: f in on for f in FLAGS}
            rcfg.DOMAIN_ADAPT.DISTILL.DISTILLER_NAME = name
            for k, v in flags.items():
                rcfg.DOMAIN_ADAPT.DISTILL[k] = v
            want = ref.build_distiller(rcfg, teacher, student)
            t, s, _ = _fake_pair()
            got = M.build_distiller(_cfg(name, **flags), t, s)
            assert type(got).__name__ == type(want).__name__ == name
            assert got.distill_enabled() == want.distill_enabled(), (name, on)
