"""Host-side logic of the drop-in shell on CPU: sizes, schedules, CLI defaults, import aliasing, pickling,
chunk partitioning and the batched streaming loader's data semantics (checked against a literal restatement of
the reference's per-sample loop)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_src_alias_is_the_same_module_object():
    import src.pretraining.models.weatherbert as a
    import weathermodel_b200.pretraining.models.weatherbert as b
    import src.base_trainer.base_trainer as c
    import weathermodel_b200.base_trainer.base_trainer as d

    assert a is b and c is d
    from src.utils.utils import get_model_params

    assert get_model_params("large") == {"num_heads": 16, "num_layers": 8, "hidden_dim_factor": 36}
    with pytest.raises(ValueError):
        get_model_params("xl")


def test_cli_flags_and_defaults_match_reference():
    from src.pretraining.pretraining_main import parser

    d = vars(parser.parse_args([]))
    assert d == {"model": "weatherformer", "resume_from_checkpoint": None, "pretrained_model_path": None,
                 "batch_size": 256, "n_masked_features": 10, "n_epochs": 100, "init_lr": 0.0005, "use_optimal_lr": False,
                 "n_warmup_epochs": 10, "decay_factor": 0.99, "model_size": "small", "masking_prob": 0.30,
                 "n_mixture_components": 1, "beta": 0.5}
    assert isinstance(parser.parse_args(["--n-warmup-epochs", "3"]).n_warmup_epochs, float)  # parsed as float


def test_schedules_match_reference_golden():
    from src.base_trainer.base_trainer import BaseTrainer
    from src.utils.utils import get_scheduler

    g = dict(np.load(os.path.join(GOLD, "schedules.npz")))
    for nm, warm, total, decay in [("exp", 10.0, 100, 0.99), ("cos", 5, 50, None), ("nowarm", 0, 20, 0.9)]:
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=5e-4)
        s = get_scheduler(opt, warm, total, decay)
        lrs = []
        for _ in range(total):
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            s.step()
        assert np.allclose(lrs, g[nm], rtol=1e-12, atol=0)
    f = BaseTrainer._get_n_masked_features
    assert [f(None, e, 10) for e in (None, 0, 4, 5, 9, 10, 37, 99)] == [10, 10, 10, 12, 12, 14, 24, 25]


def test_models_initialise_like_the_reference_and_refuse_cpu_compute():
    from src.pretraining.models.weatherbert import WeatherBERT
    from src.pretraining.models.weatherformer import WeatherFormer

    for fname, cls in [("weatherbert_mini_b8.npz", WeatherBERT), ("weatherformer_mini_b8.npz", WeatherFormer)]:
        g = dict(np.load(os.path.join(GOLD, fname)))
        torch.manual_seed(1234)
        m = cls(weather_dim=31, output_dim=31, device=torch.device("cpu"), num_heads=4, num_layers=2, hidden_dim_factor=12)
        sd = m.state_dict()
        ref_keys = sorted(k[len("param/"):] for k in g if k.startswith("param/"))
        assert sorted(sd) == ref_keys  # identical state_dict keys (checkpoint compatibility)
        for k in ref_keys:
            assert np.array_equal(sd[k].numpy(), g["param/" + k]), k  # identical init for the same seed
        assert [n for n, _ in m.named_parameters()] == [k[len("grad/"):] for k in g if k.startswith("grad/")]
    assert m.name == "weatherformer" and m.total_params() == 61262 and m.total_params_formatted() == "61.3k"
    assert m.input_dim == 34 and m.max_len == 365 and m.out_proj.out_features == 62
    w = torch.zeros(2, 365, 31)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(w, torch.zeros(2, 2), torch.zeros(2, 365), torch.zeros(2, 1), weather_feature_mask=torch.zeros(2, 365, 31, dtype=torch.bool))
    with pytest.raises(NotImplementedError):
        m(w, torch.zeros(2, 2), torch.zeros(2, 365), torch.zeros(2, 1), torch.zeros(2, 365, 31, dtype=torch.bool),
          src_key_padding_mask=torch.zeros(2, 365, dtype=torch.bool))


def test_flat_parameter_views_and_pickling_on_cpu(tmp_path):
    from src.pretraining.models.weatherbert import WeatherBERT
    from weathermodel_b200.engine import encoder_param_names

    torch.manual_seed(0)
    m = WeatherBERT(31, 31, torch.device("cpu"), num_heads=4, num_layers=2, hidden_dim_factor=12)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    assert [n for n, _ in m.named_parameters()] == encoder_param_names(2)  # reference named_parameters() order
    rt = m.runtime
    rt.ensure_flat(torch.device("cpu"))
    assert rt.flat_params.numel() % 64 == 0
    for (name, p), off in zip(rt._named, rt.offsets):
        assert p.data_ptr() == rt.flat_params.data_ptr() + 4 * off
        assert torch.equal(p.detach(), before[name])
    # gaps between tensors are zero (the padded head reads its bias pad from them)
    used = torch.zeros_like(rt.flat_params, dtype=torch.bool)
    for (_, p), off in zip(rt._named, rt.offsets):
        used[off:off + p.numel()] = True
    assert (rt.flat_params[~used] == 0).all()
    path = tmp_path / "m.pth"
    torch.save(m, path)  # the runtime (device buffers, C handle) must not be pickled
    m2 = torch.load(path, weights_only=False)
    assert "_wm_runtime" not in m2.__dict__
    for k, v in m2.state_dict().items():
        assert torch.equal(v, before[k])
    m3 = WeatherBERT(31, 31, torch.device("cpu"), num_heads=4, num_layers=2, hidden_dim_factor=12)
    m3.load_pretrained(m2)
    assert torch.equal(m3.in_proj.weight, m2.in_proj.weight) and m3.in_proj.weight is not m2.in_proj.weight


def test_chunk_partition_matches_reference_rule():
    from src.pretraining.dataloader.pretraining_dataloader import chunk_ids_for
    from src.utils.constants import NUM_DATASET_PARTS, VALIDATION_CHUNK_IDS

    train = list(set(range(NUM_DATASET_PARTS)).difference(VALIDATION_CHUNK_IDS))
    assert chunk_ids_for("train") == train and len(train) == 111
    assert chunk_ids_for("validation") == VALIDATION_CHUNK_IDS
    for world in (2, 4, 8):
        per = len(train) // world
        got = [chunk_ids_for("train", world, r) for r in range(world)]
        assert all(len(x) == per for x in got)  # equal chunk counts per rank (no DDP-style hang)
        assert sum(got, []) == train[: per * world]
        assert chunk_ids_for("validation", world, world - 1) == VALIDATION_CHUNK_IDS[(world - 1) * (8 // world): 8]


def _write_chunks(base, ids, n, seed=0, late_every=0):
    g = torch.Generator().manual_seed(seed)
    os.makedirs(base, exist_ok=True)
    for cid in ids:
        w = torch.randn(n, 365, 31, generator=g)
        coords = torch.stack([torch.rand(n, generator=g) * 120 - 60, torch.rand(n, generator=g) * 360 - 180], 1)
        idx = torch.randint(0, 2, (n,), generator=g).float()
        if late_every:
            idx[::late_every] = 2.0  # 1984 + ((2*365+364)*7)/365 > 2002: dropped by the cutoff filter
        index = torch.stack([idx, torch.full((n,), 7.0)], 1)
        torch.save(torch.utils.data.TensorDataset(w, coords, index), os.path.join(base, f"weather_dataset_weekly_{cid}.pt"))


def _reference_style_stream(paths, kind, p, n_masked, shuffle, cutoff=2002.0):
    """Literal restatement of the reference StreamingDataset.__iter__ (per-sample), CPU generator."""
    out = []
    for i in range(0, len(paths), 3):
        data = list(torch.load(paths[i + 1], weights_only=False))
        n = len(data)
        weather = torch.stack([d[0] for d in data])
        coords = torch.stack([d[1] for d in data])
        years = torch.zeros(n, 365)
        interval = torch.zeros(n, 1)
        for j, (_, _, index) in enumerate(data):
            interval[j, 0] = index[1]
            years[j] = 1984.0 + ((index[0] * 365 + torch.arange(365, dtype=torch.float32)) * index[1]) / 365
        if kind == "weatherbert":
            mask = torch.rand(n, 365, 31) < p
        else:
            mask = (torch.argsort(torch.rand(n, 31), dim=-1) < n_masked).unsqueeze(1).expand(-1, 365, -1)
        if shuffle and n > 1:
            perm = torch.randperm(n)
            weather, coords, years, interval, mask = weather[perm], coords[perm], years[perm], interval[perm], mask[perm]
        for j in range(n):
            if torch.max(years[j]) >= cutoff:
                continue
            out.append((weather[j], coords[j], years[j], interval[j], mask[j]))
    return out


@pytest.mark.parametrize("kind,shuffle", [("weatherbert", True), ("weatherformer", True), ("weatherformer", False)])
def test_streaming_loader_reproduces_reference_sample_stream(tmp_path, monkeypatch, kind, shuffle):
    import src.pretraining.dataloader.pretraining_dataloader as dl

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(dl, "DRY_RUN", True)
    base = "data/nasa_power/processed/"
    ids = [1, 34, 53, 72, 81]
    _write_chunks(base, ids, n=21, late_every=5)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    random.seed(99)
    torch.manual_seed(1234)
    loader = dl.streaming_dataloader(8, split="train", shuffle=shuffle, masking_function=kind, masking_prob=0.3,
                                     n_masked_features=7)
    paths = loader.dataset.file_paths
    batches = list(loader)
    torch.manual_seed(1234)
    torch.empty((), dtype=torch.int64).random_()  # DataLoader.__iter__ draws its base seed first (both sides)
    ref = _reference_style_stream(paths, kind, 0.3, 7, shuffle)
    assert sum(b[0].shape[0] for b in batches) == len(ref) == 5 * (21 - 5)
    assert [b[0].shape[0] for b in batches] == [8] * (len(ref) // 8) + ([len(ref) % 8] if len(ref) % 8 else [])
    flat = [tuple(t[i] for t in b) for b in batches for i in range(b[0].shape[0])]
    for got, want in zip(flat, ref):
        for a, b_ in zip(got, want):
            assert torch.equal(a, b_)
    assert batches[0][4].dtype == torch.bool and batches[0][3].shape == (8, 1)


@pytest.mark.parametrize("kind,shuffle", [("weatherbert", True), ("weatherformer", True), ("weatherformer", False)])
def test_streaming_loader_matches_the_reference_loaders_own_stream(tmp_path, monkeypatch, kind, shuffle):
    """tests/golden/loader_stream_*.npz was dumped from the UNMODIFIED reference `streaming_dataloader` iterating the
    same synthetic chunk files (oracle/make_golden_loader.py): sample order, cutoff filtering, batch boundaries across
    chunks, years, coords, intervals and every mask bit must be identical."""
    import src.pretraining.dataloader.pretraining_dataloader as dl

    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden",
                                     f"loader_stream_{kind}_{'shuffle' if shuffle else 'ordered'}.npz")))
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(dl, "DRY_RUN", True)
    _write_chunks("data/nasa_power/processed/", [1, 34, 53, 72, 81], n=21, late_every=5)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    random.seed(99)
    torch.manual_seed(1234)
    loader = dl.streaming_dataloader(8, split="train", shuffle=shuffle, masking_function=kind, masking_prob=0.3,
                                     n_masked_features=7)
    assert [int(p_.split("_")[-1].split(".")[0]) for p_ in loader.dataset.file_paths[1::3]] == gold["chunk_order"].tolist()
    batches = list(loader)
    assert [b[0].shape[0] for b in batches] == gold["batch_sizes"].tolist()
    cat = [torch.cat([b[i] for b in batches]) for i in range(5)]
    n = cat[0].shape[0]
    assert np.array_equal(cat[0][:, 0, 0].numpy(), gold["key"])
    assert np.array_equal(cat[0].double().sum((1, 2)).numpy(), gold["weather_sum"])
    assert np.array_equal(cat[1].numpy(), gold["coords"])
    assert np.array_equal(cat[2].numpy(), gold["year"])  # float32, same op order
    assert np.array_equal(cat[3].numpy(), gold["interval"])
    assert np.array_equal(np.packbits(cat[4].numpy().reshape(n, -1), axis=1), gold["mask_bits"])


def test_chunk_payload_offset_inside_the_zip_file(tmp_path):
    """The CUDA loader reads the weather tensor of a chunk file straight into a pinned buffer: the byte offset it derives
    from the zip directory must point at the tensor's storage (torch.save layout), also when the tensor is a view."""
    import numpy as np
    from weathermodel_b200.pretraining.dataloader.pretraining_dataloader import _stored_payload_span

    w = torch.randn(40, 365, 31)
    coords, index = torch.rand(40, 2), torch.rand(40, 2)
    path = str(tmp_path / "weather_dataset_weekly_1.pt")
    torch.save(torch.utils.data.TensorDataset(w, coords, index), path)
    off = _stored_payload_span(path, w.numel() * 4)
    assert off is not None and off % 64 == 0  # torch aligns records to 64 bytes
    raw = np.fromfile(path, dtype=np.float32, count=w.numel(), offset=off)
    assert np.array_equal(raw, w.numpy().ravel())
    assert _stored_payload_span(path, w.numel() * 4 + 4) is None          # no record of that size
    assert _stored_payload_span(str(tmp_path / "missing.pt"), 16) is None  # unreadable file: caller falls back


def test_cutoff_decision_on_the_host_equals_the_loader_expression():
    """_load_chunk decides "does the cutoff-year filter drop anything" from the index columns on the host; the value
    must be the one the per-chunk expression of _chunk_samples (reference pretraining_dataloader.py:251-256, 276) gives."""
    g = torch.Generator().manual_seed(3)
    for late in (0, 1, 7):
        index = torch.stack([torch.randint(0, 2, (64,), generator=g).float(), torch.full((64,), 7.0)], 1)
        if late:
            index[::late, 0] = 2.0  # 1984 + ((2 * 365 + 364) * 7) / 365 > 2002
        t = torch.arange(365, dtype=torch.float32)
        years = 1984.0 + ((index[:, 0:1] * 365 + t) * index[:, 1:2].contiguous()) / 365
        want = bool((years.max(dim=1).values < 2002.0).all())
        idx = index.float()
        got = bool(((1984.0 + ((idx[:, 0:1] * 365 + t) * idx[:, 1:2]) / 365).max(dim=1).values < 2002.0).all())
        assert got == want == (late == 0)


def test_fused_adam_host_path_matches_torch_adam():
    from weathermodel_b200.optim import FusedAdam

    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(50))
    b = torch.nn.Parameter(a.detach().clone())
    o1, o2 = FusedAdam([a], lr=1e-2, allow_host_params=True), torch.optim.Adam([b], lr=1e-2)
    for _ in range(5):
        gr = torch.randn(50)
        a.grad, b.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    sd = o1.state_dict()
    o3 = FusedAdam([a], lr=1e-2, allow_host_params=True)
    o3.load_state_dict(sd)
    assert float(o3.state[a]["step"]) == 5.0


def test_cli_module_is_runnable_with_dash_m(tmp_path):
    """`python -m src.pretraining.pretraining_main` is how the reference is launched (pretraining.sh:45-51)."""
    import subprocess

    res = subprocess.run([sys.executable, "-m", "src.pretraining.pretraining_main", "--help"], cwd=tmp_path,
                         env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "--n-masked-features" in res.stdout and "--beta" in res.stdout


def test_lr_finder_matches_the_reference_rule():
    """weathermodel_b200's find_optimal_lr against the learning rates the UNMODIFIED reference function returned for
    the same loss curves (tests/golden/lr_finder.npz, written by oracle/make_golden_lr.py): steepest decline / 10,
    floored at 10 * start_lr, iterator restarted when the loader runs out, optimiser LR restored."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden_lr import StubTrainer

    from weathermodel_b200.base_trainer.find_optimal_lr import find_optimal_lr, select_lr

    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "lr_finder.npz")))
    keys = sorted({k.rsplit("/", 1)[0] for k in gold})
    assert len(keys) >= 10
    loader = [(torch.zeros(2, 1),)] * 7
    for key in keys:
        start_lr = float(key.split("/")[1])
        tr = StubTrainer(gold[key + "/curve"])
        lr = find_optimal_lr(tr, loader, start_lr=start_lr)
        assert abs(lr - float(gold[key + "/lr"])) <= 1e-12 * max(1.0, abs(lr)), (key, lr, float(gold[key + "/lr"]))
        assert tr.calls == int(gold[key + "/calls"]), (key, tr.calls)
        assert tr.optimizer.param_groups[0]["lr"] == float(gold[key + "/restored_lr"]) == 123.0
    assert select_lr([], [], 1e-4) == 1e-3  # nothing recorded: conservative default
