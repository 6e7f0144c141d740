"""Quality metrics of the validation path, exactly as the reference defines them
(src/models/utils.py:227-250; fastmri==0.3.0 ifft2c/complex_abs/rss; skimage==0.18.1 structural_similarity)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def ifft2c(data: torch.Tensor) -> torch.Tensor:
    """Centred orthonormal inverse FFT on [..., H, W, 2]."""
    c = torch.view_as_complex(data.contiguous())
    c = torch.fft.fftshift(torch.fft.ifftn(torch.fft.ifftshift(c, dim=(-2, -1)), dim=(-2, -1), norm="ortho"), dim=(-2, -1))
    return torch.view_as_real(c)


def complex_abs(data: torch.Tensor) -> torch.Tensor:
    return (data ** 2).sum(dim=-1).sqrt()


def rss(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    return torch.sqrt((data ** 2).sum(dim))


def reconstruct(flat: torch.Tensor, shape, in_image_space: bool) -> torch.Tensor:
    """[C*H*W, 2] network output -> root-sum-of-squares magnitude image [H, W] (src/train.py:221-229)."""
    C, H, W = shape
    im = flat.reshape(C, H, W, 2)
    if not in_image_space:
        im = ifft2c(im)
    return rss(complex_abs(im), dim=0)


def psnr(x: torch.Tensor, xhat: torch.Tensor, epsilon: float = 1e-10) -> torch.Tensor:
    """10 log10(max(x) / (mse + eps)): the reference's definition, max NOT squared."""
    return 10 * torch.log10(torch.max(x) / (torch.mean((x - xhat) ** 2) + epsilon))


def ssim(x: torch.Tensor, xhat: torch.Tensor, win: int = 7, K1: float = 0.01, K2: float = 0.03) -> torch.Tensor:
    """Mean SSIM with skimage's defaults (uniform 7x7 window, sample covariance, border crop) and the
    reference's data_range = max(both) - min(both).  float64 on whatever device the inputs live on."""
    x, y = x.double()[None, None], xhat.double()[None, None]
    data_range = torch.maximum(x.max(), y.max()) - torch.minimum(x.min(), y.min())
    NP = win * win
    cov_norm = NP / (NP - 1)
    pool = lambda t: F.avg_pool2d(t, win, stride=1)           # valid windows only == skimage's cropped interior
    ux, uy = pool(x), pool(y)
    vx = cov_norm * (pool(x * x) - ux * ux)
    vy = cov_norm * (pool(y * y) - uy * uy)
    vxy = cov_norm * (pool(x * y) - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return S.mean()
