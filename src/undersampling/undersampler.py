"""Undersampling masks for accelerated-MRI fits: drop-in for the reference's src/undersampling/undersampler.py
(`Undersampler(method).apply(images[C,H,W,S], params) -> (masked images, grid [C*H*W,3], grid mask [C*H*W,3] bool)`).

Same three mask kinds and the same arithmetic (reference :81-147):
  grid         every grid_x-th row and grid_y-th column                                  (:81-92)
  random_line  rows / columns kept with probability p, two `torch.rand` draws, rows first (:96-111)
  radial       golden-angle points on the perimeters of nested squares                   (:114-147, utils.py:30-62)
What changes: no matplotlib side effects (the reference saves `undersampling_mask.png` and prints the acceleration),
the radial mask is built with one vectorised scatter per square instead of a Python loop over its points, and masks,
grid and masked images live on the device of the input tensor so the whole preparation can run on the GPU."""
from typing import Tuple

import numpy as np
import torch

GOLDEN_RATIO = (1 + np.sqrt(5)) / 2
SUPORTED_UNDERSAMPLING_METHODS = ["grid", "random_line", "radial"]


def center_crop(data, shape):
    """reference undersampling/utils.py:7-26"""
    assert 0 < shape[0] <= data.shape[-2]
    assert 0 < shape[1] <= data.shape[-1]
    w_from = (data.shape[-2] - shape[0]) // 2
    h_from = (data.shape[-1] - shape[1]) // 2
    return data[..., w_from:w_from + shape[0], h_from:h_from + shape[1]]


def verify_acc_factor(mask) -> str:
    """reference undersampling/utils.py:65-67"""
    return str((torch.numel(mask) / torch.count_nonzero(mask)).item())


def square_perimeter_points(side: int, square_id: int, idx: np.ndarray):
    """(row, col) of the idx-th point, clockwise from the top-left corner, on the perimeter of the square_id-th nested
    square of a side x side matrix -- the closed form of reference utils.py:30-62 (`get_square_ordered_idxs`)."""
    J = side - 2 * square_id                     # points along one side
    lo, hi = square_id, side - square_id - 1
    idx = np.asarray(idx)
    r = np.empty_like(idx)
    c = np.empty_like(idx)
    s1 = idx < J                                                   # top row, left -> right
    s2 = (idx >= J) & (idx < 2 * J - 2)                            # right column, downwards (corners excluded)
    s3 = (idx >= 2 * J - 2) & (idx < 3 * J - 3)                    # bottom row, right -> left
    s4 = idx >= 3 * J - 3                                          # left column, upwards
    r[s1], c[s1] = lo, lo + idx[s1]
    r[s2], c[s2] = lo + 1 + (idx[s2] - J), hi
    r[s3], c[s3] = hi, hi - (idx[s3] - (2 * J - 2))
    r[s4], c[s4] = hi - (idx[s4] - (3 * J - 3)), lo
    return r, c


def grid_mask(image_h: int, image_w: int, grid_x: int = 3, grid_y: int = 3) -> torch.Tensor:
    mask = torch.zeros((image_h, image_w), dtype=torch.bool)
    mask[::grid_x, ::grid_y] = True
    return mask


def random_line_mask(image_h: int, image_w: int, p: float) -> torch.Tensor:
    mask = torch.zeros((image_h, image_w), dtype=torch.bool)
    mask_x = torch.rand(image_h) <= p
    mask_y = torch.rand(image_w) <= p
    mask[mask_x, :] = True
    mask[:, mask_y] = True
    return mask


def radial_mask(image_shape, acceleration, rng=None) -> torch.Tensor:
    """image_shape = (C, H, W, ...).  `rng`: numpy RandomState (the reference draws from an unseeded one, :115)."""
    rng = np.random.RandomState() if rng is None else rng
    assert acceleration != 0, "Acceleration cannot be zero"
    max_dim = max(image_shape[1:3]) - max(image_shape[1:3]) % 2
    min_dim = min(image_shape[1:3]) - min(image_shape[1:3]) % 2
    num_nested_squares = max_dim // 2
    M = int(np.prod(image_shape[1:3]) / (acceleration * (max_dim / 2 - (max_dim - min_dim) * (1 + min_dim / max_dim) / 4)))
    mask = np.zeros((max_dim, max_dim), dtype=np.float32)
    t = rng.randint(low=0, high=1e4, size=1, dtype=int).item()
    m = np.arange(M)
    frac = np.mod((m + t * M) / GOLDEN_RATIO, 1)
    for square_id in range(num_nested_squares):
        K = 4 * (2 * (num_nested_squares - square_id) - 1)
        r, c = square_perimeter_points(max_dim, square_id, np.floor(frac * K).astype(np.int64))
        mask[r, c] = 1.0
    pad = ((image_shape[1] % 2, 0), (image_shape[2] % 2, 0))
    mask = np.pad(mask, pad, constant_values=0)
    return center_crop(torch.from_numpy(mask.astype(bool)), image_shape[1:3])


def coordinate_grid(channel: int, image_h: int, image_w: int, mask_image: torch.Tensor, device=None):
    """grid [C*H*W, 3] in [-1, 1] (coil, row, col) and its bool mask (the image mask repeated per coil and per
    coordinate column) -- reference :151-181."""
    Z, Y, X = torch.meshgrid(torch.linspace(-1, 1, channel, device=device), torch.linspace(-1, 1, image_h, device=device),
                             torch.linspace(-1, 1, image_w, device=device), indexing="ij")
    grid = torch.hstack((Z.reshape(-1, 1), Y.reshape(-1, 1), X.reshape(-1, 1)))
    m = mask_image.to(device)[None].expand(channel, image_h, image_w).reshape(-1, 1)
    return grid, m.expand(-1, 3).contiguous()


class Undersampler:
    def __init__(self, undersamping_method: str) -> None:
        assert undersamping_method in SUPORTED_UNDERSAMPLING_METHODS, f"Undersamping method: {undersamping_method} not supported"
        self.undersampling_method = undersamping_method
        self._mask_image = None
        self._grid = None
        self._grid_mask = None
        self.rng = None                       # optional numpy RandomState for the radial mask (None: unseeded, as the reference)

    def apply(self, images_tensor: torch.Tensor, params: list) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        assert images_tensor.dim() == 4, "For processing, please provide a 4-dimensional tensor as [batch_size, image_x, image_y, channel_n]"
        C, H, W, S = images_tensor.size()
        if self.undersampling_method == "grid":
            assert len(params) == 2, "Grid undersampling method's paramaters are not correct, it should have two parameters"
            self.create_mask_for_grid_based_undersampling(H, W, params[0], params[1])
        elif self.undersampling_method == "random_line":
            assert len(params) == 1, "Random line undersampling method's paramaters are not correct, it should have one parameters"
            self.create_mask_for_random_line_based_undersampling(H, W, params[0])
        elif self.undersampling_method == "radial":
            assert len(params) == 1, "Radial undersampling method's paramaters are not correct, it should have one parameters"
            self.create_mask_for_radial_based_undersampling(images_tensor.shape, params[0])
        else:
            raise NotImplementedError()
        dev = images_tensor.device
        self._mask_image = self._mask_image.to(dev)
        masked_tensor = images_tensor * self._mask_image.unsqueeze(0).unsqueeze(-1)
        self._grid, self._grid_mask = coordinate_grid(C, H, W, self._mask_image, dev)
        return masked_tensor, self._grid, self._grid_mask

    def __call__(self, images_tensor: torch.Tensor, params: list):
        self.apply(images_tensor, params)

    def create_mask_for_grid_based_undersampling(self, image_h, image_w, grid_x=3, grid_y=3, save_mask=False):
        self._mask_image = grid_mask(image_h, image_w, grid_x, grid_y)

    def create_mask_for_random_line_based_undersampling(self, image_h, image_w, p, save_mask=False):
        self._mask_image = random_line_mask(image_h, image_w, p)

    def create_mask_for_radial_based_undersampling(self, image_shape, acceleration, save_mask=False):
        self._mask_image = radial_mask(image_shape, acceleration, self.rng)

    @property
    def mask_image(self):
        return self._mask_image

    def get_grid_and_mask(self) -> Tuple[torch.Tensor, torch.Tensor]:
        assert self._grid is not None or self._grid_mask is not None, "Call apply() function first"
        return self._grid, self._grid_mask


def parse_undersampling_argument(arg):
    """'grid-3*3' / 'radial-2' / 'random_line-0.5' -> (method, params)  (reference data/nerp_datasets.py:256-311)."""
    if arg is None or str(arg).lower() == "none":
        return arg, []
    parts = arg.split("-")
    assert len(parts) == 2, f"Argument {arg} is incorrect"
    kind, param = parts
    if kind == "grid":
        assert "*" in param, "Please use * symbol for stating grid size"
        dims = param.split("*")
        assert len(dims) == 2, f"Grid dimensions provided ({param}) for undersampling is wrong please provide x*y format"
        return kind, [int(dims[0]), int(dims[1])]
    if kind == "random_line":
        p = float(param)
        assert 0 <= p <= 1.0, "P value is not in range [0,1]"
        return kind, [p]
    if kind == "radial":
        return kind, [float(param)]
    raise ValueError(f"Argument {kind} is not supported")
