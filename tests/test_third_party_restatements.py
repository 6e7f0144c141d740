"""Third-party arithmetic the reference pulls in and this image does not have (fastmri==0.3.0 fft2c / ifft2c /
complex_abs / rss, scikit-image==0.18.1 structural_similarity): restated in oracle/inr_oracle.py (and mirrored on the
device in the package's metrics.py) from the libraries' published definitions.  No vectors from the libraries
themselves exist here -- parity against them is UNPINNED; these tests check the restatements against independent
constructions: numpy.fft with explicit shifts, algebraic identities, and an SSIM worked out window by window."""
import numpy as np
import torch

from oracle import inr_oracle as O


def test_centered_orthonormal_fft_against_numpy_and_identities():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 12, 10, 2, generator=g, dtype=torch.float64)
    z = x[..., 0].numpy() + 1j * x[..., 1].numpy()
    want = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(z, axes=(-2, -1)), norm="ortho"), axes=(-2, -1))
    got = O.fft2c(x)
    assert np.allclose(got[..., 0].numpy() + 1j * got[..., 1].numpy(), want, atol=1e-12)
    assert torch.allclose(O.ifft2c(O.fft2c(x)), x, atol=1e-12)                     # inverse pair
    assert abs(float((got ** 2).sum()) - float((x ** 2).sum())) < 1e-9             # Parseval (orthonormal)
    # a centred delta transforms to a constant: the shifts put the origin at index N/2
    d = torch.zeros(1, 8, 8, 2, dtype=torch.float64)
    d[0, 4, 4, 0] = 1.0
    f = O.fft2c(d)
    assert torch.allclose(f[..., 0], torch.full((1, 8, 8), 1 / 8.0, dtype=torch.float64), atol=1e-12) and float(f[..., 1].abs().max()) < 1e-12


def test_complex_abs_and_rss():
    x = torch.tensor([[[3.0, 4.0], [0.0, -2.0]], [[6.0, 8.0], [1.0, 0.0]]])
    mag = O.complex_abs(x)
    assert torch.equal(mag, torch.tensor([[5.0, 2.0], [10.0, 1.0]]))
    assert torch.allclose(O.rss(mag, 0), torch.tensor([125.0 ** 0.5, 5.0 ** 0.5]))


def test_ssim_against_a_window_by_window_computation():
    rng = np.random.default_rng(3)
    a = rng.random((9, 10))
    b = a + 0.1 * rng.standard_normal((9, 10))
    L = max(a.max(), b.max()) - min(a.min(), b.min())
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    vals = []
    for i in range(3, 9 - 3):                 # centres whose 7x7 window lies inside the image (border crop of 3)
        for j in range(3, 10 - 3):
            wa, wb = a[i - 3:i + 4, j - 3:j + 4].ravel(), b[i - 3:i + 4, j - 3:j + 4].ravel()
            ma, mb = wa.mean(), wb.mean()
            va, vb = wa.var(ddof=1), wb.var(ddof=1)               # sample (co)variance, as skimage's default
            cab = ((wa - ma) * (wb - mb)).sum() / (wa.size - 1)
            vals.append(((2 * ma * mb + C1) * (2 * cab + C2)) / ((ma * ma + mb * mb + C1) * (va + vb + C2)))
    assert abs(O.ssim(a, b) - float(np.mean(vals))) < 1e-12
    assert abs(O.ssim(a, a) - 1.0) < 1e-12
