"""Synthetic fastMRI-knee-shaped multi-coil slices (there is no dataset access in this environment).

Shapes and conventions follow the reference data layer so the hot path sees the same inputs it would
get from MRIDataset: complex data as [C,H,W,2] fp32, flattened in (coil, row, col) order next to
torch.linspace(-1,1) coordinates (reference src/data/utils.py:98-108, src/data/nerp_datasets.py:101-105);
image space is normalised by the complex-magnitude maximum (:90-96), k-space by a chosen mode (:107-143)."""
from __future__ import annotations

import math

import torch


def phantom_slice(seed: int = 1234, C: int = 15, H: int = 320, W: int = 320, device="cpu") -> torch.Tensor:
    """Complex coil images [C,H,W] (complex64): ellipse phantom x smooth coil sensitivities + noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    img = torch.zeros(H, W)
    for _ in range(12):
        cx, cy = (torch.rand(2, generator=g) * 1.2 - 0.6).tolist()
        ax, ay = (torch.rand(2, generator=g) * 0.35 + 0.05).tolist()
        th = float(torch.rand(1, generator=g)) * math.pi
        amp = float(torch.rand(1, generator=g)) * 0.8 + 0.2
        xr = (xx - cx) * math.cos(th) + (yy - cy) * math.sin(th)
        yr = -(xx - cx) * math.sin(th) + (yy - cy) * math.cos(th)
        img += amp * torch.sigmoid(40 * (1 - (xr / ax) ** 2 - (yr / ay) ** 2))
    img *= torch.sigmoid(30 * (0.9 - xx ** 2 - yy ** 2 * 0.8))
    img += 0.05 * img * torch.sin(40 * xx + 3) * torch.cos(37 * yy)
    coils = []
    for c in range(C):
        ang = 2 * math.pi * c / C
        sx, sy = 1.1 * math.cos(ang), 1.1 * math.sin(ang)
        mag = torch.exp(-((xx - sx) ** 2 + (yy - sy) ** 2) / 1.5)
        ph = 1.5 * (xx * math.cos(ang + 0.3) + yy * math.sin(ang + 0.3))
        coils.append(torch.polar(mag * img, ph))
    data = torch.stack(coils)
    noise = torch.complex(torch.randn(C, H, W, generator=g), torch.randn(C, H, W, generator=g))
    data = data + 1e-3 * data.abs().max() * noise
    return data.to(torch.complex64).to(device)


def fft2c(x: torch.Tensor) -> torch.Tensor:
    """Centred orthonormal 2-D FFT over the last two dims of a complex tensor (fastmri.fft2c semantics)."""
    return torch.fft.fftshift(torch.fft.fftn(torch.fft.ifftshift(x, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))


def ifft2c(x: torch.Tensor) -> torch.Tensor:
    return torch.fft.fftshift(torch.fft.ifftn(torch.fft.ifftshift(x, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))


def coords_grid(C: int, H: int, W: int, device="cpu") -> torch.Tensor:
    """[C*H*W, 3] (coil, row, col) in [-1,1], flattened coil-major like the reference's create_coords."""
    z, y, x = torch.meshgrid(torch.linspace(-1, 1, C), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    return torch.stack([z.reshape(-1), y.reshape(-1), x.reshape(-1)], dim=1).to(device)


def make_fit_arrays(seed: int = 1234, C: int = 15, H: int = 320, W: int = 320, image_space: bool = True,
                    normalization: str = "max", device="cpu"):
    """(coords [N,3] fp32, target [N,2] fp32, (C,H,W)) for one slice, ready for grid-order batching."""
    data = phantom_slice(seed, C, H, W)
    if image_space:
        t = torch.view_as_real(data)
        t = t / data.abs().max()                                   # normalize_image
    else:
        k = torch.view_as_real(fft2c(data))
        if normalization == "coil":
            mx = torch.view_as_complex(k.contiguous()).abs().reshape(C, -1).max(dim=-1)[0]
            k = k / mx[:, None, None, None]
        elif normalization == "abs_max":
            k = k / torch.view_as_complex(k.contiguous()).abs().max()
        elif normalization == "max":
            k = k / k.abs().max()
        t = k
    target = t.reshape(C * H * W, 2).contiguous().float().to(device)
    return coords_grid(C, H, W, device).float().contiguous(), target, (C, H, W)
