"""Drop-in mirror of the reference ingest path (src/snr_denoising/dataloader.py), SURVEY.md section 8f.4.

Same names and return tuples as the reference -- `resolve_h5_path`, `NoisyWaveDataset`, `pad_collate`, `make_dataloader` -- so
`train.train_diffusion(args)` and `inference`-style scripts can open a `gen.py` HDF5 file unchanged.  What differs is WHERE the
work happens:

  * the file is read with h5py when it is importable, else with the bundled minimal reader (`_hdf5.File`: exactly the
    structures libhdf5 writes for gen.py's variable-length float32 rows, gen.py:406-413);
  * `make_dataloader` returns a `BatchLoader` instead of a torch DataLoader with worker processes: it reads the raw rows of a
    batch, LEFT-pads them (dataloader.py:248-268) straight into one of two pinned host buffers, copies it to the device on a
    side stream while the previous batch trains, and only then whitens (train-like / model PSD / saved Welch PSD,
    dataloader.py:110-143, 166-188) and estimates sigma (dataloader.py:190-200) -- on the GPU, for the whole batch, with the
    kernels of `whitening.py` (the reference does this per sample in numpy inside the workers);
  * it yields `(clean, noisy, sigma, mask, meta)` CUDA tensors with the reference's shapes ([B,1,L], [B,1,L], [B], [B,1,L],
    [B,4,L] or [B,0,L]); the training loop's `.to(device)` calls are then no-ops.

`NoisyWaveDataset.__getitem__` keeps the reference's per-sample contract (CPU tensors) for code that indexes the dataset
directly; it runs the same GPU kernels on a batch of one.  The shuffled order is what `torch.utils.data.RandomSampler` would
produce from the global torch RNG.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import whitening

__all__ = ["_mad_std", "resolve_h5_path", "open_h5", "NoisyWaveDataset", "pad_collate", "make_dataloader", "BatchLoader"]


def _mad_std(x: np.ndarray) -> float:
    """dataloader.py:10-12 (host version: a scalar helper, not on the hot path)."""
    x64 = np.asarray(x, dtype=np.float64)
    return 1.4826 * np.median(np.abs(x64 - np.median(x64))) + 1e-24


def resolve_h5_path(path: str) -> str:
    """dataloader.py:14-24: a directory resolves to its most recently modified .h5 / .hdf5 file."""
    if os.path.isdir(path):
        cands = [f for f in os.listdir(path) if f.lower().endswith((".h5", ".hdf5"))]
        if not cands:
            raise FileNotFoundError(f"No .h5/.hdf5 files found in directory: {path}")
        cands_full = [os.path.join(path, f) for f in cands]
        cands_full.sort(key=os.path.getmtime, reverse=True)
        return cands_full[0]
    if not os.path.exists(path):
        raise FileNotFoundError(f"HDF5 path not found: {path}")
    return path


def open_h5(path: str):
    """h5py.File(path, 'r') when h5py is installed, else the bundled reader (same indexing / attrs surface)."""
    try:
        import h5py                                               # noqa: F401
    except ImportError:
        from ._hdf5 import File
        return File(path, "r")
    try:
        return h5py.File(path, "r", swmr=True)
    except Exception:
        return h5py.File(path, "r")


def _finite(a: np.ndarray) -> np.ndarray:
    return a if np.isfinite(a).all() else np.nan_to_num(a, nan=0.0, posinf=0.0, neginf=0.0)     # dataloader.py:160-164


class NoisyWaveDataset(torch.utils.data.Dataset):
    """dataloader.py:26-246.  Returns (clean, noisy, sigma, mask, meta_bc) per index."""

    def __init__(self, h5_path: str, whiten: bool = False, whiten_mode: str = "auto", sigma_mode: str = "std",
                 sigma_fixed: float = 1.0, allow_no_signal: bool = False, include_metadata: bool = True,
                 mass_scale: float = 80.0):
        self.h5_path = resolve_h5_path(h5_path)
        self.h5 = None
        self._signal = self._noisy = self._psd_model = self._psd_welch = self._psd_welch_freqs = None
        self.whiten = bool(whiten)
        self.whiten_mode = str(whiten_mode).lower()
        self.sigma_mode = sigma_mode
        self.sigma_fixed = float(sigma_fixed)
        self.fs = None
        self.N = None
        self.allow_no_signal = allow_no_signal
        self.include_metadata = bool(include_metadata)
        self.mass_scale = float(mass_scale)
        self._m1 = self._m2 = self._s1 = self._s2 = None
        if sigma_mode not in ("std", "mad", "fixed"):
            raise ValueError(f"Unknown sigma_mode: {sigma_mode}")

    def _ensure_open(self) -> None:
        if self.h5 is not None:
            return
        self.h5 = open_h5(self.h5_path)
        if "noisy" not in self.h5:
            raise KeyError("HDF5 must have 'noisy' dataset.")
        self._noisy = self.h5["noisy"]
        if "signal" in self.h5:
            self._signal = self.h5["signal"]
        elif not self.allow_no_signal:
            raise KeyError("Missing 'signal' dataset. Set allow_no_signal=True for inference.")
        self._psd_model = self.h5.get("psd_model", self.h5.get("psd", None))
        self._psd_welch = self.h5.get("psd_welch", None)
        self._psd_welch_freqs = self.h5.get("psd_welch_freqs", None)
        self._m1, self._m2 = self.h5.get("mass1", None), self.h5.get("mass2", None)
        self._s1, self._s2 = self.h5.get("spin1z", None), self.h5.get("spin2z", None)
        self.N = self._noisy.shape[0]
        if self._signal is not None and self._signal.shape[0] != self.N:
            raise ValueError("Mismatched leading dimension between 'signal' and 'noisy'.")
        attrs = self.h5.attrs
        fs_attr = attrs.get("sampling_rate", 0.0)
        dt_attr = attrs.get("delta_t", 1.0 / 4096.0)
        self.fs = float(fs_attr) if float(fs_attr) > 0 else float(1.0 / float(dt_attr))

    def __len__(self) -> int:
        if self.N is None:
            self._ensure_open()
        return int(self.N)

    # ---- raw access (host): one sample's arrays exactly as stored
    def whiten_kind(self) -> str:
        """Which whitening `__getitem__` applies (dataloader.py:166-188): 'none' | 'model' | 'welch' | 'train'."""
        self._ensure_open()
        if not self.whiten:
            return "none"
        has_model = self._psd_model is not None
        has_welch = self._psd_welch is not None and self._psd_welch_freqs is not None
        if self.whiten_mode == "auto":
            return "model" if has_model else ("welch" if has_welch else "train")
        if self.whiten_mode == "model" and has_model:
            return "model"
        if self.whiten_mode == "welch" and has_welch:
            return "welch"
        return "train"

    def read_raw(self, idx: int):
        """(noisy fp32 [L], clean fp32 [L], psd rows or None, meta fp32 [4] or [0]) of sample idx, NaN/Inf scrubbed."""
        self._ensure_open()
        noisy = _finite(np.array(self._noisy[idx], dtype=np.float32))
        clean = (_finite(np.array(self._signal[idx], dtype=np.float32)) if self._signal is not None
                 else np.zeros_like(noisy, dtype=np.float32))
        kind = self.whiten_kind()
        psd = None
        if kind == "model":
            psd = (np.array(self._psd_model[idx], dtype=np.float64),)
        elif kind == "welch":
            psd = (np.array(self._psd_welch_freqs[idx], dtype=np.float64), np.array(self._psd_welch[idx], dtype=np.float64))
        if self.include_metadata:
            def _get(dset):
                try:
                    return float(np.array(dset[idx]).item())
                except Exception:
                    return 0.0
            ms = max(self.mass_scale, 1e-9)
            meta = np.array([_get(self._m1) / ms, _get(self._m2) / ms, _get(self._s1), _get(self._s2)], dtype=np.float32)
        else:
            meta = np.zeros(0, dtype=np.float32)
        return noisy, clean, psd, meta

    # ---- device side: whitening + sigma of a batch of EQUAL-length rows (dataloader.py:166-200)
    def condition_batch(self, noisy: torch.Tensor, clean: torch.Tensor, psd: Optional[Sequence[torch.Tensor]]):
        """noisy, clean: [n, L] CUDA fp32 (raw).  -> (noisy', clean' fp32 [n, L], sigma fp32 [n])."""
        kind = self.whiten_kind()
        L = noisy.shape[-1]
        if kind == "train":
            noisy, clean, _ = whitening.whiten_train_like(noisy, clean)
        elif kind == "model":
            P = whitening.interp_psd_batch(psd[0], L, self.fs)
            noisy = whitening.apply_psd(noisy, P, False, torch.float32, loader_floor=True)
            clean = whitening.apply_psd(clean, P, False, torch.float32, loader_floor=True)
        elif kind == "welch":
            P = whitening.interp_grid(psd[0], psd[1], L, self.fs)
            noisy = whitening.apply_psd(noisy, P, False, torch.float32, loader_floor=True)
            clean = whitening.apply_psd(clean, P, False, torch.float32, loader_floor=True)
        s = whitening.sigma(noisy, self.sigma_mode, self.sigma_fixed)
        s = torch.where(torch.isfinite(s) & (s > 0), s, torch.ones_like(s))          # dataloader.py:199-200
        return noisy, clean, s.float()

    def __getitem__(self, idx: int):
        noisy, clean, psd, meta = self.read_raw(int(idx))
        dev = torch.device("cuda")
        n_d, c_d = torch.from_numpy(noisy).to(dev)[None], torch.from_numpy(clean).to(dev)[None]
        psd_d = [torch.from_numpy(p).to(dev)[None] for p in psd] if psd is not None else None
        n_w, c_w, s = self.condition_batch(n_d, c_d, psd_d)
        L = noisy.shape[-1]
        noisy_t, clean_t = n_w[0].cpu().float().unsqueeze(0), c_w[0].cpu().float().unsqueeze(0)
        meta_bc = torch.from_numpy(np.tile(meta[:, None], (1, L))).float() if meta.size else torch.zeros(0, L)
        return clean_t, noisy_t, s[0].cpu(), torch.ones_like(noisy_t), meta_bc

    def close(self) -> None:
        try:
            if self.h5 is not None:
                self.h5.close()
        except Exception:
            pass
        self.h5 = None
        self._signal = self._noisy = self._psd_model = self._psd_welch = self._psd_welch_freqs = None
        self._m1 = self._m2 = self._s1 = self._s2 = None

    def __del__(self):
        self.close()


def pad_collate(batch: List[Tuple[torch.Tensor, ...]]):
    """dataloader.py:248-268: LEFT-pad every tensor of the batch to the longest sample."""
    clean_list, noisy_list, sigma_list, mask_list, meta_list = zip(*batch)
    Lmax = int(max(x.shape[-1] for x in noisy_list))

    def _pad_left(x: torch.Tensor, target: int) -> torch.Tensor:
        pad = target - x.shape[-1]
        return F.pad(x, (pad, 0)) if pad > 0 else x

    clean_pad = torch.stack([_pad_left(x, Lmax) for x in clean_list], dim=0)
    noisy_pad = torch.stack([_pad_left(x, Lmax) for x in noisy_list], dim=0)
    mask_pad = torch.stack([_pad_left(x, Lmax) for x in mask_list], dim=0)
    sigma = torch.stack([s if isinstance(s, torch.Tensor) else torch.tensor(s, dtype=torch.float32) for s in sigma_list], dim=0)
    meta_pad = torch.stack([_pad_left(x, Lmax) for x in meta_list], dim=0)
    return clean_pad, noisy_pad, sigma, mask_pad, meta_pad


class BatchLoader:
    """Batched replacement of `DataLoader(ds, collate_fn=pad_collate, pin_memory=True)`: see the module docstring."""

    def __init__(self, dataset: NoisyWaveDataset, batch_size: int = 16, shuffle: bool = True, device=None, drop_last: bool = False):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, int(batch_size), bool(shuffle), bool(drop_last)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if self.device.type != "cuda":
            raise RuntimeError("gwb200 BatchLoader stages batches for the GPU: no CPU path")
        self._pinned: List[Optional[dict]] = [None, None]
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._free = [None, None]                                 # events: the device has consumed the staging slot

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _order(self) -> List[int]:
        n = len(self.dataset)
        if not self.shuffle:
            return list(range(n))
        # torch.utils.data.RandomSampler.__iter__: a generator seeded from the global RNG, then randperm
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g).tolist()

    def _stage(self, slot: int, idxs: List[int]):
        """Host half: read rows, left-pad into pinned buffers; then async H2D on the copy stream.  Returns device tensors + meta."""
        ds = self.dataset
        rows = [ds.read_raw(i) for i in idxs]
        B, Lmax = len(rows), max(r[0].shape[-1] for r in rows)
        Cm = rows[0][3].shape[0]
        pin = self._pinned[slot]
        if pin is None or pin["cap"][0] < B or pin["cap"][1] < Lmax:
            cap_B = max(B, self.batch_size)
            cap_L = max(Lmax, pin["cap"][1] if pin else 0)
            pin = {"cap": (cap_B, cap_L), "noisy": torch.empty(cap_B, cap_L).pin_memory(), "clean": torch.empty(cap_B, cap_L).pin_memory(),
                   "dev_noisy": torch.empty(cap_B, cap_L, device=self.device), "dev_clean": torch.empty(cap_B, cap_L, device=self.device)}
            self._pinned[slot] = pin
            self._free[slot] = None
        if self._free[slot] is not None:
            self._free[slot].synchronize()                        # the batch that used this slot two iterations ago is consumed
        noisy_h, clean_h = pin["noisy"][:B, :Lmax], pin["clean"][:B, :Lmax]
        lens = []
        for b, (noisy, clean, _psd, _meta) in enumerate(rows):
            L = noisy.shape[-1]
            lens.append(L)
            noisy_h[b, : Lmax - L] = 0.0
            clean_h[b, : Lmax - L] = 0.0
            noisy_h[b, Lmax - L:] = torch.from_numpy(noisy)
            clean_h[b, Lmax - L:] = torch.from_numpy(clean)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            dn, dc = pin["dev_noisy"][:B, :Lmax], pin["dev_clean"][:B, :Lmax]
            dn.copy_(noisy_h, non_blocking=True)
            dc.copy_(clean_h, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record()
        return {"slot": slot, "B": B, "Lmax": Lmax, "lens": lens, "rows": rows, "Cm": Cm, "dn": dn, "dc": dc, "ready": ready}

    def _finish(self, st):
        """Device half: per length group whiten + sigma on the raw (unpadded) rows, then write them back left-padded."""
        ds, dev = self.dataset, self.device
        B, Lmax, lens, rows = st["B"], st["Lmax"], st["lens"], st["rows"]
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(st["ready"])
        noisy = torch.zeros(B, 1, Lmax, device=dev)
        clean = torch.zeros(B, 1, Lmax, device=dev)
        mask = torch.zeros(B, 1, Lmax, device=dev)
        sigma = torch.empty(B, device=dev)
        kind = ds.whiten_kind()
        for L in sorted(set(lens)):
            sel = [b for b in range(B) if lens[b] == L]
            idx = torch.tensor(sel, device=dev)
            n_raw = st["dn"][idx, Lmax - L:].contiguous()
            c_raw = st["dc"][idx, Lmax - L:].contiguous()
            psd = None
            if kind in ("model", "welch"):
                k = len(rows[sel[0]][2])
                psd = [torch.from_numpy(np.stack([rows[b][2][j] for b in sel])).to(dev) for j in range(k)]
            n_w, c_w, s = ds.condition_batch(n_raw, c_raw, psd)
            noisy[idx, 0, Lmax - L:] = n_w
            clean[idx, 0, Lmax - L:] = c_w
            mask[idx, 0, Lmax - L:] = 1.0
            sigma[idx] = s
        ev = torch.cuda.Event()
        ev.record(cur)
        self._free[st["slot"]] = ev
        Cm = st["Cm"]
        if Cm:
            mv = torch.from_numpy(np.stack([r[3] for r in rows])).to(dev)              # [B, 4]
            meta = mv[:, :, None] * mask                                                  # broadcast over the valid (unpadded) part
        else:
            meta = torch.zeros(B, 0, Lmax, device=dev)
        return clean, noisy, sigma, mask, meta

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        order = self._order()
        bs = self.batch_size
        batches = [order[i:i + bs] for i in range(0, len(order), bs)]
        if self.drop_last and batches and len(batches[-1]) < bs:
            batches.pop()
        staged = self._stage(0, batches[0]) if batches else None
        for k in range(len(batches)):
            nxt = self._stage((k + 1) & 1, batches[k + 1]) if k + 1 < len(batches) else None     # H2D of k+1 overlaps batch k
            yield self._finish(staged)
            staged = nxt


def make_dataloader(h5_path: str, batch_size: int = 16, shuffle: bool = True, num_workers: int = 4, prefetch_factor: int = 2,
                    persistent_workers: bool = True, pin_memory: Optional[bool] = None, whiten: bool = False,
                    whiten_mode: str = "auto", sigma_mode: str = "std", sigma_fixed: float = 1.0, allow_no_signal: bool = False,
                    include_metadata: bool = True, mass_scale: float = 80.0, device=None) -> BatchLoader:
    """dataloader.py:270-310.  `num_workers`, `prefetch_factor`, `persistent_workers`, `pin_memory` are accepted for signature
    compatibility: staging is always pinned + double-buffered, and the per-sample numpy work of the workers runs on the GPU."""
    ds = NoisyWaveDataset(h5_path=h5_path, whiten=whiten, whiten_mode=whiten_mode, sigma_mode=sigma_mode, sigma_fixed=sigma_fixed,
                          allow_no_signal=allow_no_signal, include_metadata=include_metadata, mass_scale=mass_scale)
    return BatchLoader(ds, batch_size=batch_size, shuffle=shuffle, device=device)
