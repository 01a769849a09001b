"""Registration shim: make the reference's own entry points build the B200 classes, with zero YAML / script changes.

    import mml_b200.shim as shim; shim.install()      # before StandardMultimodalConfig.load(...)

What it re-binds (reference file:line):
  * YAML tags ``!ResNet18`` / ``!ResNet34`` / ``!ResNetEncoder`` (MML_Suite/config/yaml_constructors.py:159-178) ->
    ``mml_b200.resnet`` factories (``yaml.SafeLoader.add_constructor`` replaces the earlier registration);
  * YAML tags ``!MNISTAudio`` / ``!MNISTImage`` / ``!ConvBlockArgs`` / ``!ConvBlock`` (yaml_constructors.py:70-86) -> ``mml_b200.convblock``;
  * ``resolve_model_name("avmnist")`` (MML_Suite/config/resolvers.py:18-22 does ``from models.avmnist import AVMNIST``
    at call time) -> the attribute ``models.avmnist.AVMNIST`` is replaced by ``mml_b200.avmnist.AVMNIST``;
  * ``models.msa.networks.resnet.ResNet18/ResNet34/ResNetEncoder`` for scripts that import them directly
    (MML_Suite/train_monomodal.py);
  * config 3: YAML tags ``!MMIMDb`` / ``!MMIMDbModalityEncoder`` / ``!GatedBiModalNetwork`` / ``!MLPGenreClassifier``
    (yaml_constructors.py:126-142) and the attributes ``models.mmimdb.MMIMDb`` (+ the three part classes) ->
    ``mml_b200.mmimdb``.
  * ``install(datasets=True)`` (opt-in): ``config.resolvers.AVMNIST / MOSI / MOSEI / MMIMDb`` -- the names ``resolve_dataset_name`` looks up at call
    time (config/resolvers.py:192-221) -- and ``data.AVMNIST / MOSI / MOSEI / MMIMDb`` -> ``mml_b200.datasets``: ``dataset: "AVMNIST"`` in a YAML
    then builds the pinned in-memory dataset (same items through a DataLoader, plus ``fused_loader()`` for the fused step).
Works without the reference on the path too (then only the YAML tags are registered).
"""
from __future__ import annotations

import sys
from typing import Dict

_installed: Dict[str, object] = {}


def install(patch_reference_modules: bool = True, datasets: bool = False) -> Dict[str, object]:
    import yaml

    from .avmnist import AVMNIST
    from .resnet import ResNet18, ResNet34, ResNetEncoder

    def register(tag, factory):
        def construct(loader, node):
            return factory(**loader.construct_mapping(node, deep=True))

        yaml.SafeLoader.add_constructor(tag, construct)
        _installed[tag] = factory

    register("!ResNet18", ResNet18)
    register("!ResNet34", ResNet34)
    register("!ResNetEncoder", ResNetEncoder)
    from . import convblock as _cb

    # yaml_constructors.py:70-86 (train_avmnist.yaml builds AVMNIST from these four tags)
    for cname in ("MNISTAudio", "MNISTImage", "ConvBlockArgs", "ConvBlock"):
        register("!" + cname, getattr(_cb, cname))
    from . import mmimdb as _mm

    gated = {"MMIMDb": _mm.MMIMDb, "MMIMDbModalityEncoder": _mm.MMIMDbModalityEncoder, "GatedBiModalNetwork": _mm.GatedBiModalNetwork,
             "MLPGenreClassifier": _mm.MLPGenreClassifier}
    for cname, cls in gated.items():
        register("!" + cname, cls)
    if patch_reference_modules:
        ref_av = sys.modules.get("models.avmnist")
        if ref_av is None:
            try:
                import models.avmnist as ref_av  # noqa: F401  (only importable when MML_Suite is on sys.path)
            except Exception:
                ref_av = None
        if ref_av is not None:
            _installed["reference.AVMNIST"] = getattr(ref_av, "AVMNIST", None)
            ref_av.AVMNIST = AVMNIST
            ref_av.MNISTAudio, ref_av.MNISTImage = _cb.MNISTAudio, _cb.MNISTImage
        ref_conv = sys.modules.get("models.conv")
        if ref_conv is not None:
            ref_conv.ConvBlock, ref_conv.ConvBlockArgs = _cb.ConvBlock, _cb.ConvBlockArgs
        ref_mm = sys.modules.get("models.mmimdb")
        if ref_mm is not None:
            _installed["reference.MMIMDb"] = getattr(ref_mm, "MMIMDb", None)
            for cname, cls in gated.items():
                setattr(ref_mm, cname, cls)
        ref_rn = sys.modules.get("models.msa.networks.resnet")
        if ref_rn is not None:
            ref_rn.ResNet18, ref_rn.ResNet34, ref_rn.ResNetEncoder = ResNet18, ResNet34, ResNetEncoder
    _installed["AVMNIST"] = AVMNIST
    _installed["MMIMDb"] = _mm.MMIMDb
    from . import utt_fusion as _uf

    utt = {"UttFusionModel": _uf.UttFusionModel, "LSTMEncoder": _uf.LSTMEncoder, "TextCNN": _uf.TextCNN, "FcClassifier": _uf.FcClassifier}
    for cname, cls in utt.items():  # yaml_constructors.py registers the same tags for models/msa/*
        register("!" + cname, cls)
    if patch_reference_modules:
        for modname, names in (("models.msa.utt_fusion", ("UttFusionModel",)), ("models.msa.networks.lstm", ("LSTMEncoder",)),
                               ("models.msa.networks.textcnn", ("TextCNN",)), ("models.msa.networks.classifier", ("FcClassifier",))):
            mod = sys.modules.get(modname)
            if mod is not None:
                for n in names:
                    setattr(mod, n, utt[n])
    _installed["UttFusionModel"] = _uf.UttFusionModel
    from .mono import MonomodalEncoder

    ref_tm = sys.modules.get("train_monomodal")
    if patch_reference_modules and ref_tm is not None:  # MonomodalEncoder lives in the training script itself (train_monomodal.py:64)
        _installed["reference.MonomodalEncoder"] = getattr(ref_tm, "MonomodalEncoder", None)
        ref_tm.MonomodalEncoder = MonomodalEncoder
    _installed["MonomodalEncoder"] = MonomodalEncoder
    if datasets:
        from . import datasets as _ds

        for modname in ("config.resolvers", "data"):
            mod = sys.modules.get(modname)
            if patch_reference_modules and mod is not None:
                for n in ("AVMNIST", "MOSI", "MOSEI", "MMIMDb"):  # module-level names of the DATASET classes (models are imported locally)
                    _installed.setdefault(f"reference.{modname}.{n}", getattr(mod, n, None))
                    setattr(mod, n, getattr(_ds, n))
        _installed["datasets"] = (_ds.AVMNIST, _ds.MOSI, _ds.MOSEI, _ds.MMIMDb)
    return dict(_installed)
