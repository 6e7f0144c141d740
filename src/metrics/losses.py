"""Drop-in for src/metrics/losses.py: loss objects with the reference's names / call conventions.

When the training script can fuse (model + one of these losses), only ``kind`` / options are read and the
loss runs inside the forward kernel's epilogue; called directly they evaluate with torch ops (autograd),
which keeps the unfused ``model(x) -> loss -> backward`` path and validation working."""
import torch


class MSLELoss(torch.nn.Module):            # reference :18-27
    kind = "MSLE"

    def __init__(self, eps=1e-9):
        super().__init__()
        self.eps = eps

    def forward(self, x, y):
        return torch.nn.functional.mse_loss(torch.log(x + 1 + self.eps), torch.log(y + 1 + self.eps))


class TanhL2Loss(torch.nn.Module):          # reference :121-139 (with_mag off by default)
    kind = "tanh"

    def forward(self, x, y, kcoords=None):
        return torch.mean((torch.tanh(x) - torch.tanh(y)) ** 2), 0


def _cplx(t):
    return torch.view_as_complex(t.contiguous()) if t.dtype == torch.float else t


class LogSpaceLoss(torch.nn.Module):        # reference :204-223
    kind = "LSL"

    def __init__(self, config):
        super().__init__()
        self.sigma, self.eps, self.factor = float(config["hdr_ff_sigma"]), float(config["hdr_eps"]), float(config["hdr_ff_factor"])

    def forward(self, input, target):
        x, y = _cplx(input), _cplx(target)
        return (((x - y).abs() / (x.detach().abs() + self.eps)) ** 2).mean()


class HDRLoss_FF(torch.nn.Module):          # reference :226-264, evaluated in its separable form
    kind = "HDR"

    def __init__(self, config):
        super().__init__()
        self.sigma, self.eps, self.factor = float(config["hdr_ff_sigma"]), float(config["hdr_eps"]), float(config["hdr_ff_factor"])

    def forward(self, input, target, kcoords, weights=None, reduce=True):
        x, y = _cplx(input), _cplx(target)
        kcoords = kcoords.to(x.device)
        d = x.detach().abs() + self.eps
        loss = torch.log((x - y).abs() / d) ** 2
        if weights is not None:
            loss = loss * weights.unsqueeze(-1)
        f = torch.exp(-(kcoords[..., 1] ** 2 + kcoords[..., 2] ** 2) / (2 * self.sigma ** 2))
        if not reduce:
            return loss, self.factor * ((1 - f) ** 2).unsqueeze(-1) * (x.abs() / d) ** 2
        reg = self.factor * ((1 - f) ** 2).mean() * ((x.abs() / d) ** 2).mean()   # == mean of the [bs_k, m] outer product
        return loss.mean() + reg, reg


class ConsistencyLoss(torch.nn.Module):    # reference :292-324
    """Supervises output i+1 with (detached) output i on the points OUTSIDE disc i."""

    def __init__(self, bounds):
        super().__init__()
        self.bounds = bounds

    def forward(self, input, dist):
        loss = 0
        for i in range(len(self.bounds) - 1):
            lo, hi = self.bounds[i]
            ind = torch.where((dist < lo) | (dist > hi))
            if ind[0].numel():
                loss = loss + torch.nn.functional.mse_loss(input[i][ind].detach(), input[i + 1][ind])
        return loss


def tv_loss(img, weight=0.0001):            # reference :326-343
    w_variance = torch.nn.functional.l1_loss(img[:, :-1, :], img[:, 1:, :])
    h_variance = torch.nn.functional.l1_loss(img[:-1, :, :], img[1:, :, :])
    return weight * (h_variance + w_variance)
