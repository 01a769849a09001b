"""Import the UNMODIFIED reference (MML_Suite) in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under task-specific-pretraining-multimodal_b200/ may import
this file.  It is used by ``oracle/make_golden.py`` (run here, where
``/root/reference`` is mounted) to pin the oracle restatement against the
reference's own modules.  It cannot travel to the GPU box (the reference is not
there), which is why the golden vectors it produces are committed under
``tests/golden/``.

Recipe: three third-party modules the reference imports are
absent from this image (``modalities`` -- an un-pinned git dependency,
pyproject.toml:13 --, ``matplotlib`` and ``h5py``); they are only needed for
enums / annotations / loaders, so attribute-permissive stubs are enough.  The
``config`` package has to be imported before ``models`` because of a circular
import in the reference (models/avmnist.py:9 -> experiment_utils/metric_recorder.py:11
-> config/__init__.py:40 -> ...).
"""
from __future__ import annotations

import enum
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MML_REFERENCE_ROOT", "/root/reference")
SUITE = os.path.join(REFERENCE_ROOT, "MML_Suite")


def reference_available() -> bool:
    return os.path.isdir(SUITE)


class _Permissive(types.ModuleType):
    """Module stub: any attribute resolves to a dummy callable/class."""

    def __getattr__(self, name):  # pragma: no cover - trivial
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


def _install_stubs() -> None:
    if "modalities" not in sys.modules:
        mod = types.ModuleType("modalities")

        class Modality(enum.Enum):
            AUDIO = "audio"
            IMAGE = "image"
            TEXT = "text"
            VIDEO = "video"
            MULTIMODAL = "multimodal"

            def __str__(self) -> str:
                return self.value

            # The reference sorts Modality members (config/data_config.py:67,74); the pattern names it
            # documents ("ai", "it", "atv": data/mosi.py:62-70, data/mmimdb.py:73-77) are alphabetical in
            # the modality name, so the stub orders by name.  ASSUMPTION about the un-vendored package.
            def __lt__(self, other) -> bool:
                return self.value < other.value

            @classmethod
            def from_str(cls, s: str) -> "Modality":
                return cls(str(s).lower())

        def add_modality(name: str):
            return Modality.from_str(name) if str(name).lower() in Modality._value2member_map_ else None

        def create_missing_mask(n_modalities, batch_size, missing_rates):
            # Contract observed at the only call site (data/base_dataset.py:53-57):
            # returns [batch_size, n_modalities] of 0/1, arg 3 = per-modality MISSING rate.
            import torch

            keep = 1.0 - torch.tensor(list(missing_rates), dtype=torch.float32)
            return torch.bernoulli(keep.expand(batch_size, n_modalities)).float()

        mod.Modality = Modality
        mod.add_modality = add_modality
        mod.create_missing_mask = create_missing_mask
        sys.modules["modalities"] = mod
    for name in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "h5py", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Permissive(name)
    if isinstance(sys.modules.get("matplotlib"), _Permissive):
        sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def import_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if SUITE not in sys.path:
        sys.path.insert(0, SUITE)
    import config.multimodal_training_config  # noqa: F401  (must precede models.*)
    from config.data_config import MissingPatternConfig, ModalityConfig
    from experiment_utils.loss import LossFunctionGroup
    from models.avmnist import AVMNIST
    from models.msa.networks.resnet import ResNet18, ResNet34
    from modalities import Modality
    from models.gates import GatedBiModalNetwork
    from models.mmimdb import MLPGenreClassifier, MMIMDb, MMIMDbModalityEncoder

    from models.avmnist import MNISTAudio, MNISTImage
    from models.conv import ConvBlockArgs
    from models.msa.networks.classifier import FcClassifier
    from models.msa.networks.lstm import LSTMEncoder
    from models.msa.networks.textcnn import TextCNN
    from models.msa.utt_fusion import UttFusionModel

    ns = types.SimpleNamespace(
        MNISTAudio=MNISTAudio, MNISTImage=MNISTImage, ConvBlockArgs=ConvBlockArgs,
        UttFusionModel=UttFusionModel, LSTMEncoder=LSTMEncoder, TextCNN=TextCNN, FcClassifier=FcClassifier,
        MMIMDb=MMIMDb,
        MMIMDbModalityEncoder=MMIMDbModalityEncoder,
        MLPGenreClassifier=MLPGenreClassifier,
        GatedBiModalNetwork=GatedBiModalNetwork,
        AVMNIST=AVMNIST,
        ResNet18=ResNet18,
        ResNet34=ResNet34,
        LossFunctionGroup=LossFunctionGroup,
        MissingPatternConfig=MissingPatternConfig,
        ModalityConfig=ModalityConfig,
        Modality=Modality,
    )
    return ns
