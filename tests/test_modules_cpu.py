"""CPU tests of the host-side mirror: constructors, state_dict surface, RNG-order parity with the oracle."""
import os
import re
import sys

import pytest
import torch

import late_fusion_oracle as O


def _build(seed=0, dropout=0.5):
    from mml_b200.avmnist import AVMNIST
    from mml_b200.resnet import ResNet18, ResNet34

    torch.manual_seed(seed)
    return AVMNIST(ResNet18(in_channels=1, hidden_dim=64), ResNet34(in_channels=1, hidden_dim=128), hidden_dim=128, dropout=dropout, fusion_fn="concat")


def test_state_dict_matches_reference_surface_and_init():
    model = _build(0)
    sd = model.state_dict()
    torch.manual_seed(0)
    ref = O.init_avmnist_state()
    assert list(sd.keys()) == list(ref.keys())
    assert len(sd) == 346
    assert sum(p.numel() for p in model.parameters()) == 32_580_746
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        assert torch.equal(sd[k], ref[k]), f"init differs at {k}"


def test_constructor_contract():
    from mml_b200.avmnist import AVMNIST
    from mml_b200.resnet import ResNet18, ResNet34, ResNetEncoder

    enc = ResNet18(1, 64)
    assert isinstance(enc, ResNetEncoder) and enc.get_embedding_size() == 64
    assert ResNet34(hidden_dim=128).get_embedding_size() == 128
    with pytest.raises(ValueError):
        AVMNIST(enc, ResNet34(), 128, fusion_fn="sum")
    m = AVMNIST(enc, ResNet34(), 128, dropout=0.0)
    assert isinstance(m.net[2], torch.nn.Identity)
    assert m.get_encoder("audio") is enc
    with pytest.raises(ValueError):
        m.get_encoder("text")


def test_no_cpu_fallback():
    model = _build(0)
    with pytest.raises(RuntimeError, match="no CPU"):
        model.forward(A=torch.zeros(2, 112, 112), I=torch.zeros(2, 1, 28, 28))
    with pytest.raises(RuntimeError, match="GPU only"):
        model.audio_encoder(torch.zeros(2, 112, 112))


def test_loss_and_optimizer_guards():
    from mml_b200.avmnist import AVMNIST

    class Term:
        def __init__(self, fn, w=1.0):
            self.loss_fn, self.weight = fn, w

    AVMNIST._check_loss({"ce": Term(torch.nn.CrossEntropyLoss())})
    for bad in ({"ce": Term(torch.nn.CrossEntropyLoss(label_smoothing=0.1))}, {"ce": Term(torch.nn.CrossEntropyLoss(), 0.5)},
                {"mse": Term(torch.nn.MSELoss())}, {"a": Term(torch.nn.CrossEntropyLoss()), "b": Term(torch.nn.CrossEntropyLoss())}):
        with pytest.raises(NotImplementedError):
            AVMNIST._check_loss(bad)


def test_c_abi_library_loads_and_exports_every_header_symbol():
    from mml_b200 import _lib

    lib = _lib.load_library()
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mml_b200.h")).read()
    declared = set(re.findall(r"\b(mml_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mml_version() >= 100


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "task-specific-pretraining-multimodal_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "late_fusion_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_shim_registers_yaml_tags_and_swaps_reference_model():
    import yaml

    from mml_b200 import shim
    from mml_b200.avmnist import AVMNIST
    from mml_b200.resnet import ResNetEncoder

    got = shim.install()
    enc = yaml.safe_load("enc: !ResNet18\n  in_channels: 1\n  hidden_dim: 64\n")["enc"]
    assert isinstance(enc, ResNetEncoder) and enc.get_embedding_size() == 64
    enc34 = yaml.safe_load("enc: !ResNet34\n  hidden_dim: 128\n")["enc"]
    assert len(enc34.blocks()) == 16
    assert got["AVMNIST"] is AVMNIST
    # with the reference importable (build container only) the resolver must hand out the B200 class
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_import

    if ref_import.reference_available():
        ref_import.import_reference()
        shim.install()
        from config.resolvers import resolve_model_name

        assert resolve_model_name("avmnist") is AVMNIST
        cfg = yaml.safe_load("a: !ResNet18\n  in_channels: 1\n  hidden_dim: 64\n")
        assert isinstance(cfg["a"], ResNetEncoder)
        # opt-in dataset swap: ``dataset: "AVMNIST"`` / "mosi" of a YAML resolve to the pinned in-memory classes, and the reference's own
        # MissingPatternConfig output (Modality-keyed probabilities) is a valid ``missing_patterns`` argument for them
        from collections import OrderedDict

        from config.resolvers import resolve_dataset_name
        from mml_b200 import datasets as D

        ref_cls = resolve_dataset_name("avmnist")
        got = shim.install(datasets=True)
        assert resolve_dataset_name("avmnist") is D.AVMNIST and resolve_dataset_name("MOSI") is D.MOSI and resolve_dataset_name("mosei") is D.MOSEI
        assert resolve_dataset_name("mm_imdb") is D.MMIMDb and resolve_model_name("mmimdb") is not D.MMIMDb  # the MODEL of that name is resolved separately
        assert got["reference.config.resolvers.AVMNIST"] is ref_cls and ref_cls is not D.AVMNIST
        ns = ref_import.import_reference()
        M = ns.Modality
        pats = ns.MissingPatternConfig(modalities=OrderedDict([(M.AUDIO, ns.ModalityConfig(missing_rate=0.2, apply_to=None)),
                                                               (M.IMAGE, ns.ModalityConfig(missing_rate=0.0, apply_to=None))]),
                                       selected_patterns=["ai"]).generate_patterns()
        import numpy as np
        import torch

        ds = D.AVMNIST.from_arrays(torch.arange(6) % 10, torch.rand(6, 4, 5), torch.zeros(6, 3, 3, dtype=torch.uint8), "train",
                                   missing_patterns=pats, selected_patterns=["ai"], cmap=np.zeros((256, 3)), pin=False)
        assert ds.missing_patterns == {"ai": {"audio": 0.8, "image": 1.0}} and ds.masks["ai"]["image"].all()
        assert M.AUDIO in ds[0] and ds[0]["pattern_name"] == "ai"  # Modality-keyed items once the ``modalities`` package is importable
        import config.resolvers as R
        import data as RDATA

        for n in ("AVMNIST", "MOSI", "MOSEI", "MMIMDb"):  # leave the reference modules as they were for the other tests
            setattr(R, n, got[f"reference.config.resolvers.{n}"])
            setattr(RDATA, n, got[f"reference.data.{n}"])


def test_ctypes_signatures_match_header_arity():
    """Every ctypes argtypes list has exactly as many entries as the C prototype has parameters."""
    from mml_b200 import _lib

    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mml_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for m in re.finditer(r"\b(mml_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert name in _lib.SIGNATURES, name
        assert len(_lib.SIGNATURES[name][1]) == n, f"{name}: header has {n} parameters, binding declares {len(_lib.SIGNATURES[name][1])}"


def test_pattern_generation_matches_reference_fixture_and_masks():
    """mml_b200.data.generate_patterns against tests/golden/patterns.txt (written from the reference's MissingPatternConfig)."""
    import os
    from collections import OrderedDict

    import torch

    from mml_b200 import data

    cases = {
        "avmnist_audio02": (OrderedDict([("audio", (0.2, None)), ("image", (0.0, None))]), ["ai"]),
        "avmnist_all": (OrderedDict([("audio", (0.2, None)), ("image", (0.3, None))]), None),
        "avmnist_apply": (OrderedDict([("audio", (0.25, ["a"])), ("image", (0.5, ["ai", "i"]))]), None),
        "mosi_3": (OrderedDict([("audio", (0.2, None)), ("video", (0.0, None)), ("text", (0.9, None))]), None),
    }
    want = {}
    with open(os.path.join(os.path.dirname(__file__), "golden", "patterns.txt")) as f:
        for line in f:
            c, pat, mod, v = line.strip().split("|")
            want.setdefault(c, {}).setdefault(pat, {})[mod] = float(v)
    for cname, (mods, sel) in cases.items():
        got = data.generate_patterns(mods, sel)
        assert {k: dict(v) for k, v in got.items()} == want[cname], cname
    pats = data.generate_patterns(cases["avmnist_audio02"][0], ["ai"])
    masks = data.draw_missing_masks(pats, 20000, torch.Generator().manual_seed(1))
    assert set(masks["ai"]["audio"].unique().tolist()) == {0.0, 1.0}
    assert abs(float(masks["ai"]["audio"].mean()) - 0.8) < 0.01 and float(masks["ai"]["image"].min()) == 1.0
    b = data.attach_masks({"audio": torch.zeros(2, 4), "image": torch.zeros(2, 3), "labels": torch.zeros(2)},
                          {"audio": torch.ones(2), "image": torch.ones(2)}, ["audio", "image"])
    assert set(b) == {"audio_original", "audio_missing_index", "image_original", "image_missing_index", "labels"}


def test_shim_builds_the_other_configs_from_the_reference_yaml_tags():
    """The YAML tags of configs 3 / 4 build the B200 classes with the reference's state_dict layout (CPU: construction only)."""
    import yaml

    import gated_fusion_oracle as G
    import utt_fusion_oracle as U
    from mml_b200 import mmimdb, mono, shim, utt_fusion
    from mml_b200.resnet import ResNet18

    got = shim.install()
    assert got["MMIMDb"] is mmimdb.MMIMDb and got["UttFusionModel"] is utt_fusion.UttFusionModel and got["MonomodalEncoder"] is mono.MonomodalEncoder
    # mmimdb_baseline.yaml:10-31
    torch.manual_seed(0)
    cfg = yaml.safe_load(
        "img: !MMIMDbModalityEncoder {input_dim: 4096, output_dim: 512}\n"
        "txt: !MMIMDbModalityEncoder {input_dim: 300, output_dim: 512}\n"
        "gmu: !GatedBiModalNetwork {input_one_dim: 512, output_one_dim: 512, input_two_dim: 512, output_two_dim: 512}\n"
        "clf: !MLPGenreClassifier {input_size: 512, hidden_size: 512, output_size: 23}\n")
    model = mmimdb.MMIMDb(cfg["img"], cfg["txt"], gated_bimodal_network=cfg["gmu"], classifier=cfg["clf"])
    torch.manual_seed(0)
    ref = G.init_mmimdb_state()
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys()) and all(torch.equal(sd[k], ref[k]) for k in ref)
    # mmimdb_pooling.yaml: pooling built inside the model constructor, after the classifier
    for pt in ("max", "avg", "sum", "attention", "gated"):
        torch.manual_seed(0)
        m2 = mmimdb.MMIMDb(mmimdb.MMIMDbModalityEncoder(4096, 512), mmimdb.MMIMDbModalityEncoder(300, 512),
                           multimodal_pooling={"pooling_type": pt, "hidden_dim": 512, "dropout": 0.1}, classifier=mmimdb.MLPGenreClassifier(512, 23, 512))
        torch.manual_seed(0)
        r2 = G.init_mmimdb_pooling_state(pt)
        s2 = m2.state_dict()
        assert list(s2.keys()) == list(r2.keys()) and all(torch.equal(s2[k], r2[k]) for k in r2), pt
    # utt_fusion_base_training.yaml:14-46
    torch.manual_seed(0)
    cfg = yaml.safe_load(
        "a: !LSTMEncoder {input_size: 5, hidden_size: 64, embd_method: last}\n"
        "v: !LSTMEncoder {input_size: 20, hidden_size: 64, embd_method: last}\n"
        "t: !TextCNN {input_size: 768, embd_size: 64, dropout: 0.5, in_channels: 1, out_channels: 128, kernel_heights: [3, 4, 5]}\n"
        "c: !FcClassifier {input_dim: 192, layers: [192, 64, 32], output_dim: 3, dropout: 0.5}\n")
    um = utt_fusion.UttFusionModel(cfg["a"], cfg["v"], cfg["t"], cfg["c"], clip=1.0)
    torch.manual_seed(0)
    ru = U.init_utt_state()
    su = um.state_dict()
    assert list(su.keys()) == list(ru.keys()) and all(torch.equal(su[k], ru[k]) for k in ru)
    # monomodal wrapper (train_monomodal.py:68-71)
    torch.manual_seed(0)
    mm = mono.MonomodalEncoder(ResNet18(1, 64), 64, 10)
    torch.manual_seed(0)
    rm = O.init_monomodal_state("resnet18", 1, 64, 10)
    sm = mm.state_dict()
    assert list(sm.keys()) == list(rm.keys()) and all(torch.equal(sm[k], rm[k]) for k in rm)
    # ... and around an MMIMDb encoder (configs/mmimdb/mono/*.yaml)
    import gated_fusion_oracle as G
    torch.manual_seed(0)
    mv = mono.MonomodalEncoder(mmimdb.MMIMDbModalityEncoder(300, 512), 512, 23)
    torch.manual_seed(0)
    rv = G.init_mono_vector_state(300, 512, 23)
    sv = mv.state_dict()
    assert list(sv.keys()) == list(rv.keys()) and all(torch.equal(sv[k], rv[k]) for k in rv)
    # ... and around the MOSI encoders (configs/mosi/mono/*.yaml)
    for kind, make, d in (("lstm", lambda: utt_fusion.LSTMEncoder(20, 64, "last"), 20), ("textcnn", lambda: utt_fusion.TextCNN(768, 64, 1, 128, [3, 4, 5], 0.5), 768)):
        torch.manual_seed(0)
        ms = mono.MonomodalEncoder(make(), 64, 3)
        torch.manual_seed(0)
        rs = U.init_mono_seq_state(kind, d)
        ss = ms.state_dict()
        assert list(ss.keys()) == list(rs.keys()) and all(torch.equal(ss[k], rs[k]) for k in rs)
    with pytest.raises(NotImplementedError):
        mono.MonomodalEncoder(torch.nn.Linear(4, 4), 4, 3)
    with pytest.raises(RuntimeError):
        mv.train_step({"text": torch.zeros(2, 300), "label": torch.zeros(2, 23)}, torch.optim.Adam(mv.parameters()), None, torch.device("cpu"), None)
    # no CPU execution path anywhere
    with pytest.raises(RuntimeError):
        model.train_step({"image": torch.zeros(2, 4096), "text": torch.zeros(2, 300), "label": torch.zeros(2, 23), "pattern_name": ["it"] * 2},
                         torch.optim.Adam(model.parameters()), None, torch.device("cpu"), None)
    with pytest.raises(RuntimeError):
        um.train_step({"audio": torch.zeros(2, 50, 5), "video": torch.zeros(2, 50, 20), "text": torch.zeros(2, 50, 768), "label": torch.zeros(2, dtype=torch.long),
                       "pattern_name": ["atv"] * 2}, torch.optim.Adam(um.parameters()), None, torch.device("cpu"), None)


def test_flat_state_layout_and_param_group_ranges(monkeypatch):
    """Host logic of FlatState on CPU tensors (the bf16 shadow cast is the only device call; it is patched out here):
    augmented weight|bias blocks, strided views, load_state_dict through them, optimizer param groups -> flat ranges."""
    from mml_b200 import engine, ops

    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = torch.nn.Sequential(torch.nn.BatchNorm1d(300), torch.nn.Linear(300, 64))
            self.head = torch.nn.Linear(64, 10)

    torch.manual_seed(0)
    net = Net()
    ref = {k: v.clone() for k, v in net.state_dict().items()}
    fs = engine.FlatState(net, torch.device("cpu"), augment={"enc.1.weight": "enc.1.bias"})
    o, n_out, n_in, ld = fs.aug["enc.1.weight"]
    assert (n_out, n_in, ld) == (64, 300, 320) and fs.aug["enc.1.bias"] == fs.aug["enc.1.weight"]
    m = fs.aug_matrix(fs.P, "enc.1.weight")
    assert torch.equal(m[:, :300], ref["enc.1.weight"]) and torch.equal(m[:, 300], ref["enc.1.bias"]) and float(m[:, 301:].abs().max()) == 0.0
    assert net.enc[1].weight.data_ptr() == m.data_ptr() and net.enc[1].bias.data_ptr() == m[:, 300].data_ptr()
    assert net.enc[1].weight.grad.data_ptr() == fs.aug_matrix(fs.G, "enc.1.weight").data_ptr()
    for k, v in net.state_dict().items():  # names, shapes and values survive the re-homing
        assert v.shape == ref[k].shape and torch.equal(v, ref[k]), k
    # load_state_dict writes through the strided views into the flat buffer
    new = {k: (torch.full_like(v, 0.5) if v.is_floating_point() else v) for k, v in ref.items()}
    net.load_state_dict(new)
    assert float(m[:, :301].min()) == 0.5 and float(m[:, 301:].abs().max()) == 0.0 and fs.is_bound()
    # param groups: encoder (BN + augmented Linear) and head with different hyper-parameters -> two contiguous ranges
    opt = torch.optim.Adam([{"params": list(net.enc.parameters()), "lr": 1e-4, "weight_decay": 2e-4}, {"params": list(net.head.parameters()), "lr": 5e-4}])
    fs.adopt_optimizer(opt)
    fs.sync_hyper(opt, 0.5)
    assert len(fs.ranges) == 2 and fs.ranges[0][0] == 0 and fs.ranges[0][1] == fs.ranges[1][0] == fs.offsets["head.weight"] and fs.ranges[1][1] == fs.total
    assert fs.adam_ranges(0, fs.total) == [(0, fs.ranges[0][1], 0), (fs.ranges[0][1], fs.total, 1)]
    assert abs(float(fs.hyper[0, 0]) - 1e-4) < 1e-10 and abs(float(fs.hyper[1, 0]) - 5e-4) < 1e-10 and float(fs.hyper[1, 5]) == 0.5
    assert opt.state[net.head.weight]["exp_avg"].data_ptr() == fs.M.data_ptr() + 4 * fs.offsets["head.weight"]
    v0 = fs.range_version
    fs.adopt_optimizer(opt)  # same grouping: nothing changes
    assert fs.range_version == v0
    fs.adopt_optimizer(torch.optim.Adam(net.parameters(), lr=1e-3))
    assert fs.ranges == [(0, fs.total, 0)] and fs.range_version == v0 + 1
    with pytest.raises(NotImplementedError):  # weight and bias of an augmented Linear cannot be in different groups
        fs.adopt_optimizer(torch.optim.Adam([{"params": [net.enc[1].weight]}, {"params": [p for p in net.parameters() if p is not net.enc[1].weight]}]))
    with pytest.raises(NotImplementedError):
        fs.adopt_optimizer(torch.optim.SGD(net.parameters(), lr=0.1))


def test_optimizer_state_is_revalidated_after_load_state_dict(monkeypatch):
    """CheckpointManager resume path: ``optimizer.load_state_dict`` replaces the state tensors; the next step must pick the restored
    moments / step up (and alias them again) instead of continuing on the old buffers.  A brand-new Adam starts from zero moments."""
    import copy

    from mml_b200 import engine, ops

    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
    fs = engine.FlatState(net, torch.device("cpu"))
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    fs.adopt_optimizer(opt)
    assert float(fs.M.abs().max()) == 0.0 and int(fs.step) == 0
    # pretend three fused steps happened
    fs.M.copy_(torch.arange(fs.total, dtype=torch.float32) * 0.01)
    fs.V.copy_(torch.arange(fs.total, dtype=torch.float32) * 0.02)
    fs.step.fill_(3)
    fs._host_step.fill_(3.0)
    saved = copy.deepcopy(opt.state_dict())
    assert float(saved["state"][0]["step"]) == 3.0 and torch.equal(saved["state"][0]["exp_avg"], fs.M[:32].view(4, 8))
    # ... two more steps, then resume from the checkpoint
    fs.M.add_(1.0)
    fs.V.add_(1.0)
    fs.step.fill_(5)
    fs._host_step.fill_(5.0)
    fs.adopt_optimizer(opt)  # unchanged optimizer: nothing to do
    assert int(fs.step) == 5
    opt.load_state_dict(saved)
    assert opt.state[net[0].weight]["exp_avg"].data_ptr() != fs.M.data_ptr()  # torch replaced the tensors
    fs.adopt_optimizer(opt)
    assert int(fs.step) == 3 and float(fs._host_step) == 3.0
    assert torch.equal(fs.M[:32], torch.arange(32, dtype=torch.float32) * 0.01) and torch.equal(fs.V[:32], torch.arange(32, dtype=torch.float32) * 0.02)
    assert opt.state[net[0].weight]["exp_avg"].data_ptr() == fs.M.data_ptr() and opt.state[net[1].bias]["step"] is fs._host_step
    # a brand-new optimizer starts like torch does: zero moments, step 0
    opt2 = torch.optim.Adam(net.parameters(), lr=1e-3)
    fs.adopt_optimizer(opt2)
    assert float(fs.M.abs().max()) == 0.0 and float(fs.V.abs().max()) == 0.0 and int(fs.step) == 0
    assert opt2.state[net[0].weight]["exp_avg"].data_ptr() == fs.M.data_ptr()


def test_convblock_modules_match_reference_surface_and_padded_storage(monkeypatch):
    """MNISTAudio / MNISTImage / ConvBlock (avmnist.py:34-185, conv.py:16-59): same constructor keywords, state_dict names, shapes and
    initial values as the oracle's restatement of the reference; FlatState(pad=...) stores narrow layers as 64-channel tensors with a
    zero tail and exposes the logical slice."""
    from mml_b200 import engine, ops
    from mml_b200.avmnist import AVMNIST
    from mml_b200.convblock import CP, ConvBlock, ConvBlockArgs, MNISTAudio, MNISTImage, pad_map

    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))
    A = ConvBlockArgs
    torch.manual_seed(0)
    au = MNISTAudio(conv_block_one_one_args=A(1, 32), conv_block_one_two_args=A(32, 32), conv_block_two_one_args=A(32, 64),
                    conv_block_two_two_args=A(64, 64), hidden_dim=64, conv_batch_norm=True)
    im = MNISTImage(A(1, 32), A(32, 64), A(64, 64), A(64, 64), 128, max_pool_kernel_size=(2, 2))
    model = AVMNIST(au, im, 128, dropout=0.5)
    torch.manual_seed(0)
    ref = O.init_convblock_avmnist_state()
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]) for k in ref)
    assert au.get_embedding_size() == 64 and im.get_embedding_size() == 128 and au.pool_k == [2, 3] and im.pool_k == [2, 2]
    with pytest.raises(NotImplementedError):
        ConvBlock(A(1, 32, conv_one_kernel_size=5), A(32, 32))
    with pytest.raises(NotImplementedError):
        MNISTAudio(A(1, 32), A(32, 32), A(32, 64), A(64, 64), 64, conv_batch_norm=False)
    with pytest.raises(RuntimeError):  # no CPU path
        au(torch.zeros(2, 32, 94))
    pad = {**pad_map(au, "audio_encoder."), **pad_map(im, "image_encoder.")}
    assert "audio_encoder.net.0.conv_one.weight" not in pad and pad["audio_encoder.net.0.conv_two.weight"] == CP
    fs = engine.FlatState(model, torch.device("cpu"), pad=pad)
    for k, v in model.state_dict().items():
        assert v.shape == ref[k].shape and torch.equal(v, ref[k]), k
    w = fs.flat_slice(fs.P, "audio_encoder.net.0.conv_two.weight").view(CP, 3, 3, CP)
    assert float(w[32:].abs().max()) == 0.0 and float(w[..., 32:].abs().max()) == 0.0
    assert torch.equal(w[:32, :, :, :32].permute(0, 3, 1, 2), ref["audio_encoder.net.0.conv_two.weight"])
    gm = fs.flat_slice(fs.P, "audio_encoder.net.0.batch_norm_one.weight")
    assert gm.numel() == CP and float(gm[:32].min()) == 1.0 and float(gm[32:].abs().max()) == 0.0
    o = fs.buf_offsets["audio_encoder.net.0.batch_norm_one.running_var"]
    assert float(fs.S[o:o + 32].min()) == 1.0 and float(fs.S[o + 32:o + CP].abs().max()) == 0.0  # padded variance 0 -> eval scale 0
    # load_state_dict goes through the strided views; the padding stays zero
    model.load_state_dict({k: (torch.full_like(v, 0.25) if v.is_floating_point() else v) for k, v in ref.items()})
    assert float(w[:32, :, :, :32].min()) == 0.25 and float(w[32:].abs().max()) == 0.0 and float(w[..., 32:].abs().max()) == 0.0
    assert model.audio_encoder.net[0].conv_two.weight.grad.shape == (32, 32, 3, 3)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, weight_decay=1e-4)
    fs.adopt_optimizer(opt)
    assert fs.ranges == [(0, fs.total, 0)]
    assert opt.state[model.audio_encoder.net[0].conv_two.weight]["exp_avg"].shape == (32, 32, 3, 3)


def test_batch_unpacking_follows_the_reference_batch_contracts():
    """Batch dict handling of the four step methods (keys are ``modalities.Modality`` members in the reference, str() lower-case)."""
    import enum

    from mml_b200.avmnist import AVMNIST
    from mml_b200.mmimdb import MMIMDb
    from mml_b200.mono import MonomodalEncoder
    from mml_b200.utt_fusion import UttFusionModel

    class Modality(enum.Enum):
        AUDIO, IMAGE, TEXT, VIDEO = "audio", "image", "text", "video"

        def __str__(self):
            return self.value

    t = lambda *s: torch.zeros(*s)
    # AVMNIST: reference contract (already masked tensors) and the device-mask extension
    A, I, ma, mi, y, pat = AVMNIST._unpack(None, {Modality.AUDIO: t(2, 4, 4), Modality.IMAGE: t(2, 1, 3, 3), "labels": t(2), "pattern_name": ["ai", "a"]})
    assert A.shape == (2, 4, 4) and ma is None and mi is None and pat == ["ai", "a"]
    A, I, ma, mi, y, pat = AVMNIST._unpack(None, {"audio_original": t(2, 4, 4) + 1, "audio_missing_index": t(2), "image": t(2, 1, 3, 3), "labels": t(2)})
    assert float(A.min()) == 1.0 and ma is not None and mi is None
    with pytest.raises(KeyError):
        AVMNIST._unpack(None, {"audio": t(2, 4, 4), "labels": t(2)})
    # MMIMDb: "label" (singular), image / text
    Ib, Tb, m_i, m_t, yb, pb = MMIMDb._unpack(None, {Modality.IMAGE: t(2, 8), Modality.TEXT: t(2, 5), "label": t(2, 23), "pattern_name": ["it", "t"]})
    assert Ib.shape == (2, 8) and Tb.shape == (2, 5) and yb.shape == (2, 23) and m_i is None
    # UttFusion: three modalities, any subset given as original + mask
    out = UttFusionModel._unpack(None, {Modality.AUDIO: t(2, 6, 5), Modality.VIDEO: t(2, 6, 20), "text_original": t(2, 6, 768) + 2, "text_missing_index": t(2),
                                        "label": t(2), "pattern_name": ["atv", "av"]})
    assert out[3][0] is None and out[3][1] is None and out[3][2] is not None and float(out[2].min()) == 2.0
    # monomodal: prefers <modality>_original (the un-masked tensor), config name selects the modality
    cfg = type("C", (), {"experiment": type("E", (), {"name": "AVMNIST_Image_Encoder_Resnet_Pretrain"})()})()
    key, x, lab = MonomodalEncoder._unpack({"AUDIO": t(2, 4, 4), "IMAGE": t(2, 3, 3), "IMAGE_original": t(2, 3, 3) + 3, "IMAGE_missing_index": t(2),
                                            "labels": torch.zeros(2, dtype=torch.long), "pattern_name": ["ai"] * 2}, cfg)
    assert key == "IMAGE" and float(x.min()) == 3.0
    with pytest.raises(NotImplementedError):
        MonomodalEncoder._unpack({"AUDIO": ["a.pt", "b.pt"], "labels": torch.zeros(2, dtype=torch.long)}, None)
