"""Host half of the ingest path: the bundled minimal HDF5 reader / writer, path resolution, left-pad collate.
No GPU needed.  The fixture `tests/golden/ingest_fixture.h5` has gen.py's layout (gen.py:406-413); `ingest.npz` holds what the
unmodified reference data loader returned for it (make_golden.py:gen_ingest)."""
import os

import numpy as np
import pytest
import torch


def test_hdf5_reader_on_fixture_matches_reference_raw_batch(golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _hdf5
    f = _hdf5.File(os.path.join(golden_dir, "ingest_fixture.h5"))
    g = np.load(os.path.join(golden_dir, "ingest.npz"))
    assert {"signal", "noisy", "noise", "lengths", "mass1", "mass2", "spin1z", "spin2z", "psd_model"} <= set(f.keys())
    assert float(f.attrs["sampling_rate"]) == 4096.0 and f.attrs["time_axis"] == "seconds-rel-peak"
    lens = f["lengths"][...]
    assert lens.tolist() == [1024, 768, 1024, 512, 1000, 768] and f["noisy"].shape == (6,)
    noisy_ref, mask_ref = g["raw_std/noisy"], g["raw_std/mask"]          # reference: no whitening -> raw rows, left-padded
    for i, L in enumerate(lens):
        row = f["noisy"][i]
        assert row.dtype == np.float32 and row.shape == (L,)
        want = noisy_ref[i, 0, 1024 - L:]
        ok = np.isfinite(row)
        assert np.array_equal(row[ok], want[ok]) and np.all(want[~ok] == 0.0)       # the NaN was scrubbed by the reference
        assert mask_ref[i, 0, :1024 - L].sum() == 0 and mask_ref[i, 0, 1024 - L:].all()
    assert f["psd_model"].shape == (6, 513) and f["psd_model"][2].dtype == np.float64
    assert f.get("psd_welch") is None and "psd" not in f
    f.close()


def test_hdf5_write_read_round_trip(tmp_path):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import _hdf5
    rng = np.random.default_rng(0)
    rows = [rng.standard_normal(n).astype(np.float32) for n in (100, 257, 4096, 1, 0, 33)]
    times = [np.arange(len(r), dtype=np.float64) * 0.25 for r in rows]
    data = {"signal": rows, "times": times, "lengths": np.array([len(r) for r in rows], dtype=np.int64), "b": np.float32([[1, 2], [3, 4]])}
    data.update({f"extra{i:02d}": np.full(3, i, dtype=np.float64) for i in range(20)})      # > 8 links: several symbol-table nodes
    p = str(tmp_path / "t.h5")
    _hdf5.write_file(p, data, {"sampling_rate": 4096.0, "note": "hello", "n": np.int64(6)})
    with _hdf5.File(p) as f:
        assert len(f.keys()) == 24 and f.attrs["note"] == "hello" and int(f.attrs["n"]) == 6
        for i, r in enumerate(rows):
            assert np.array_equal(f["signal"][i], r) and np.array_equal(f["times"][i], times[i])
        assert f["signal"][-1].shape == (33,) and len(f["signal"][1:3]) == 2
        assert np.array_equal(f["b"][...], data["b"]) and float(f["extra07"][1]) == 7.0
        with pytest.raises(KeyError):
            f["missing"]
        with pytest.raises(IndexError):
            f["signal"][6]
    with pytest.raises(OSError):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not hdf5" * 20)
        _hdf5.File(str(bad))


def test_resolve_path_and_pad_collate(tmp_path, golden_dir):
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import dataloader as D
    with pytest.raises(FileNotFoundError):
        D.resolve_h5_path(str(tmp_path / "nope.h5"))
    with pytest.raises(FileNotFoundError):
        D.resolve_h5_path(str(tmp_path))
    (tmp_path / "a.h5").write_bytes(b"x")
    (tmp_path / "b.hdf5").write_bytes(b"y")
    os.utime(tmp_path / "a.h5", (1, 1))
    assert D.resolve_h5_path(str(tmp_path)).endswith("b.hdf5")
    items = []
    for L in (5, 3, 4):
        x = torch.arange(L, dtype=torch.float32).view(1, L)
        items.append((x, x + 10, torch.tensor(float(L)), torch.ones(1, L), torch.full((4, L), 0.5)))
    clean, noisy, sigma, mask, meta = D.pad_collate(items)
    assert clean.shape == (3, 1, 5) and meta.shape == (3, 4, 5) and sigma.tolist() == [5.0, 3.0, 4.0]
    assert clean[1, 0].tolist() == [0, 0, 0, 1, 2] and mask[1, 0].tolist() == [0, 0, 1, 1, 1] and meta[2, :, 0].sum() == 0
    with pytest.raises(ValueError):
        D.NoisyWaveDataset(os.path.join(golden_dir, "ingest_fixture.h5"), sigma_mode="bogus")


def test_measurement_loaders_match_reference(golden_dir):
    """`_load_measurement_from_h5` / `_meta_to_stack` (inference.py:59-122) against the outputs of the reference's own functions on
    the ingest fixture (tests/golden/make_golden.py::gen_checkpoint)."""
    import os
    import numpy as np
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    g = np.load(os.path.join(golden_dir, "checkpoint.npz"))
    y, clean, fs, P_model, (fw, Pw), meta = inf._load_measurement_from_h5(os.path.join(golden_dir, "ingest_fixture.h5"), 4)
    assert y.dtype == np.float32 and np.array_equal(y, g["meas/y"]) and np.array_equal(clean, g["meas/clean"])
    assert fs == float(g["meas/fs"]) and np.array_equal(P_model, g["meas/P_model"]) and fw is None and Pw is None
    assert sorted(meta) == list(g["meas/meta_keys"])
    assert np.array_equal(np.array([meta[k] for k in sorted(meta)]), g["meas/meta_vals"])
    meta["q"] = meta["mass1"] / meta["mass2"]
    meta["chirp_mass"] = 21.5
    for need in (1, 3, 5, 7, 9):
        st = inf._meta_to_stack(meta, 64, need, 65.0, 10.0)
        ref = g[f"meta_stack/{need}"]
        if need == 1:
            assert st is None and ref.shape[0] == 0
        else:
            assert st.dtype == np.float32 and st.shape == ref.shape and np.array_equal(st, ref), need
    assert inf._meta_to_stack({"q": float("nan")}, 8, 6, 80.0, 10.0)[4].max() == 0.0      # inference.py:113


def test_load_checkpoint_reads_reference_format(golden_dir):
    """A checkpoint written by the reference classes (payload keys of train.py:606-630) loads with strict=True; the architecture
    comes from ckpt['args'] with the reference's fallbacks and EMA weights win when present (inference.py:614-650)."""
    import os
    import torch
    from diffusion_models_for_gravitational_waveform_reconstruction_b200 import inference as inf
    path = os.path.join(golden_dir, "ref_checkpoint.pth")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"model_state", "optimizer_state", "args", "epoch", "model_ema_state"}
    model, diff, ck_args = inf.load_checkpoint(path, device="cpu", use_ema=True)
    assert (model.in_ch, model.cond_in_ch, model.use_selfcond) == (7, 5, True)
    assert model.spec.base_ch == 16 and model.spec.depth == 2 and model.spec.time_dim == 32 and model.spec.max_time == 999.0
    assert diff.T == 1000 and ck_args["meta_scale"] == {"M": 65.0, "q": 10.0}
    sd = model.state_dict()
    assert list(sd) == list(ck["model_ema_state"])
    assert all(torch.equal(sd[k], ck["model_ema_state"][k]) for k in sd)
    raw, _, _ = inf.load_checkpoint(path, device="cpu", use_ema=False)
    assert all(torch.equal(raw.state_dict()[k], ck["model_state"][k]) for k in sd)
    assert not torch.equal(raw.state_dict()["final.weight"], sd["final.weight"])
