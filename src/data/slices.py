"""Slice data for the drop-in entry points: same normalisation / flattening / masking conventions as the
reference data layer (src/data/nerp_datasets.py, src/undersampling/undersampler.py), vectorised and kept
resident instead of served by a per-sample DataLoader.

fastMRI .h5 files are read when h5py and the file exist; otherwise a synthetic fastMRI-knee-shaped slice
(mri_implicit_neural_representations_b200.synthetic) stands in -- there is no dataset access here."""
from __future__ import annotations

import glob
import math
import os
import warnings
from typing import Optional

import torch

from mri_implicit_neural_representations_b200 import synthetic


def _load_h5_slice(data, data_root, set_name, sample, slice_no):
    try:
        import h5py  # noqa: F401
    except ImportError:
        return None
    files = sorted(glob.glob(os.path.join(data_root, f"{data}_multicoil_{set_name}", "*.h5")))
    if not files:
        return None
    import h5py
    with h5py.File(files[sample], "r") as hf:
        k = torch.from_numpy(hf["kspace"][slice_no])          # [C,H,W] complex64
    return k, files[sample]


def grid_mask(H: int, W: int, gx: int, gy: int) -> torch.Tensor:
    """mask[::gx, ::gy] = True  (reference undersampler.py:79-91)."""
    m = torch.zeros(H, W, dtype=torch.bool)
    m[::gx, ::gy] = True
    return m


class SliceDataset:
    """coords [N,3], image [N,2], optional coords_mask [N,3] bool, dist_to_center [N]; img_shape (C,H,W,2)."""

    def __init__(self, data="knee", data_root="data", set="train", transform=True, sample=0, slice=0, full_norm=False,
                 normalization="max", undersampling: Optional[str] = None, use_dists=False, shape=(15, 320, 320), seed=None,
                 device=None):
        loaded = _load_h5_slice(data, data_root, set, sample, slice)
        if loaded is not None:
            kspace, self.file = loaded
            img = synthetic.ifft2c(kspace)
            C, H, W = img.shape
            ch, cw = min(H, 320), min(W, 320)                  # recon-size centre crop (reference :64-76)
            img = img[:, (H - ch) // 2:(H - ch) // 2 + ch, (W - cw) // 2:(W - cw) // 2 + cw]
        else:
            warnings.warn("no fastMRI file available: using a synthetic knee-shaped slice")
            seed = 1234 + 17 * int(sample) + int(slice) if seed is None else seed
            img = synthetic.phantom_slice(seed, *shape)
            self.file = f"synthetic://knee/sample{sample}/slice{slice}"
        if device is not None:
            # the whole preparation below (FFT, normalisation, masks, coordinate grid, distances) is plain tensor
            # arithmetic: on `device` it runs on the GPU and the arrays are born resident
            img = img.to(device)
        C, H, W = img.shape
        if transform:
            t = torch.view_as_real(img) / img.abs().max()         # normalize_image; `normalization` ignored (reference :64-68)
        else:
            k = torch.view_as_real(synthetic.fft2c(img))
            if normalization == "coil":
                mx = torch.view_as_complex(k.contiguous()).abs().reshape(C, -1).max(dim=-1)[0]
                k = k / mx[:, None, None, None]
            elif normalization == "abs_max":
                k = k / torch.view_as_complex(k.contiguous()).abs().max()
            elif normalization == "max":
                k = k / k.abs().max()
            elif normalization == "gaussian_blur":
                # reference nerp_datasets.py:117-123 + data/utils.py:11-28: max-normalise, then a separable Gaussian
                # (sigma 0.1, support +-ceil(10 sigma)) over (H, W) of the real and the imaginary plane of every coil
                k = k / k.abs().max()
                sigma = 0.1
                radius = math.ceil(10.0 * sigma)
                support = torch.arange(-radius, radius + 1, dtype=torch.float, device=k.device)
                kern = torch.distributions.Normal(loc=0, scale=sigma).log_prob(support).exp_()
                kern = kern.mul_(1 / kern.sum())
                planes = k.permute(0, 3, 1, 2).reshape(-1, 1, H, W)            # [C*2, 1, H, W]
                planes = torch.nn.functional.conv2d(planes, kern.view(1, 1, -1, 1), padding=(radius, 0))
                planes = torch.nn.functional.conv2d(planes, kern.view(1, 1, 1, -1), padding=(0, radius))
                k = planes.reshape(C, 2, H, W).permute(0, 2, 3, 1).contiguous()
            elif normalization == "tonemap":
                # reference nerp_datasets.py:129-133
                k = k / (k + 1)
                k = k / k.max()
                k = k - k.mean(dim=(1, 2, 3), keepdim=True)
            elif normalization == "max_std":
                k = k / k.abs().max()
                k = (k - k.mean()) / k.std()
                k = k / k.max()
            elif normalization == "stand":
                k = (k - k.mean()) / (k.std() + 1e-9)
            t = k
        self.img_shape = (C, H, W, 2)
        self.shape = t.shape
        self.coords = synthetic.coords_grid(C, H, W, device=img.device).float().contiguous()
        self.coords_mask = None
        if undersampling is not None and str(undersampling).lower() != "none":
            # reference data/nerp_datasets.py:256-334: 'grid-x*y' | 'random_line-p' | 'radial-acc' -> Undersampler.apply;
            # masked-out samples are zeroed (undersampler.py:59-61), coords_mask = the image mask per coil and column
            from undersampling.undersampler import Undersampler, parse_undersampling_argument
            kind, params = parse_undersampling_argument(undersampling)
            self.undersampler = Undersampler(kind)
            t, _, self.coords_mask = self.undersampler.apply(t, params)
        self.image = t.reshape(C * H * W, 2).float().contiguous()
        self.dist_to_center = None
        if use_dists:
            self.dist_to_center = torch.sqrt(self.coords[:, 1] ** 2 + self.coords[:, 2] ** 2)

    def __len__(self):
        return self.coords.shape[0]


class GridOrderLoader:
    """Yields (coords, gt, dist, mask) batches in grid order like the reference's DataLoader(shuffle=False) +
    collate_inr (src/models/utils.py:47-53,84-99), by slicing -- no per-sample __getitem__.
    `with_mask=False` reproduces MRIDatasetWithDistances.__getitem__ (nerp_datasets.py:391-396), which hands out an
    empty mask even for an undersampled slice: with distances and per-sample batches the reference's loops therefore run
    over ALL points (the masked-out ones hold zeros); the per-coil wrapper (:428-441) does pass the mask."""

    def __init__(self, ds: SliceDataset, batch_size: int, with_mask: bool = True):
        self.ds, self.bs, self.with_mask = ds, int(batch_size), bool(with_mask)

    def __len__(self):
        return (len(self.ds) + self.bs - 1) // self.bs

    def __iter__(self):
        ds = self.ds
        for i in range(0, len(ds), self.bs):
            j = i + self.bs
            dist = ds.dist_to_center[i:j] if ds.dist_to_center is not None else []
            mask = ds.coords_mask[i:j] if (ds.coords_mask is not None and self.with_mask) else []
            yield ds.coords[i:j], ds.image[i:j], dist, mask


def get_data_loader(data, data_root, set, batch_size, transform=True, num_workers=0, sample=0, slice=0,
                    challenge="multicoil", shuffle=True, full_norm=False, normalization="max", use_dists="no",
                    undersampling=None, per_coil=False, shape=(15, 320, 320), device=None):
    """Same signature / return triple as reference src/models/utils.py:57-141 (`shuffle` is ignored there too)."""
    assert data in ["brain", "knee"], "Unsupported parameter is provided in the get_data_loader() function"
    use_d = use_dists in ("yes", True)
    full = SliceDataset(data, data_root, set, transform, sample, slice, full_norm, normalization, None, use_d, shape,
                        device=device)
    train = full
    if undersampling is not None and str(undersampling).lower() != "none":
        train = SliceDataset(data, data_root, set, transform, sample, slice, full_norm, normalization, undersampling, use_d, shape,
                             device=device)
    C, H, W, _ = full.img_shape
    train_bs = H * W if per_coil else batch_size
    return full, GridOrderLoader(train, train_bs, with_mask=per_coil or not use_d), GridOrderLoader(full, batch_size)
