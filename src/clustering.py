"""Ring partition of k-space by k-means over radial statistics (reference src/clustering.py:19-134).

`no_steps` equal-width rings over [0, sqrt(2)] (closed on both sides, so a point on an edge counts for both rings,
reference :48-59); per ring the MAXIMUM of log|k| (:60-61); k-means of those `no_steps` numbers into `no_parts`
clusters (sklearn, init="random", n_init=10, max_iter=200, random_state=42, :63-70); the share of rings of every
cluster, in order of first appearance going outwards, accumulates into the partition radii (:71-84); the last radius
is 5 so that the last partition covers everything (:87).

The ring statistic is the only part that touches all N samples: it runs on whatever device the arrays live on (one
`scatter_reduce(amax)` per edge side instead of `no_steps` boolean gathers); the k-means of 40 numbers stays on the host
exactly as in the reference."""
from math import sqrt

import numpy as np
import torch


def ring_edges(no_steps: int):
    """[(r0, r1)] of the initial rings exactly as reference :48-57 computes them (Python doubles)."""
    out = []
    for i in range(no_steps):
        r0 = 0 if i == 0 else sqrt(2) * i / no_steps
        r1 = sqrt(2) if i == no_steps - 1 else sqrt(2) * (i + 1) / no_steps
        out.append((r0, r1))
    return out


def ring_log_max(img: torch.Tensor, dist_to_center: torch.Tensor, no_steps: int) -> torch.Tensor:
    """max over each ring of log(complex_abs(img)) (reference :58-61).  img [...,2], dist_to_center [...] (same leading
    shape).  A point whose distance equals an edge belongs to both neighbouring rings (>= r0 & <= r1)."""
    mag = torch.log(torch.sqrt((img.reshape(-1, 2) ** 2).sum(-1)))            # fastmri.complex_abs then log
    d = dist_to_center.reshape(-1)
    edges = ring_edges(no_steps)
    # torch compares a float32 tensor with a Python scalar in float32: use the same rounded edges
    lo = torch.tensor([e[0] for e in edges], dtype=d.dtype, device=d.device)
    hi = torch.tensor([e[1] for e in edges], dtype=d.dtype, device=d.device)
    # ring by its lower edge: last i with lo[i] <= d; the ring below also holds the point when d == hi[i-1]
    idx = torch.clamp(torch.searchsorted(lo, d, right=True) - 1, 0, no_steps - 1)
    inside = d <= hi[idx]                                  # beyond sqrt(2): in no ring (cannot happen on the [-1,1] grid)
    out = torch.full((no_steps,), float("-inf"), dtype=mag.dtype, device=mag.device)
    out = out.scatter_reduce(0, idx[inside], mag[inside], reduce="amax", include_self=True)
    below = torch.clamp(idx - 1, min=0)
    on_edge = (idx > 0) & (d <= hi[below])
    out = out.scatter_reduce(0, below[on_edge], mag[on_edge], reduce="amax", include_self=True)
    return out


def radii_from_labels(labels, no_parts: int):
    """Reference :73-87: share of rings per cluster in order of first appearance -> cumulative radii, last one = 5."""
    labels = np.asarray(labels)
    unique_elements, indices, counts = np.unique(labels, return_counts=True, return_index=True)
    order = np.argsort(indices)
    counts = counts[order]
    radii = np.array([0] + list(sqrt(2) * np.cumsum(counts / len(labels))))
    radii[no_parts] = 5
    return radii


def _arrays(dataset, img, kcoords):
    if dataset is None and (img is None or kcoords is None):
        raise ValueError("Dataset or image must be provided")
    if dataset is not None:
        C, H, W, S = dataset.shape
        img = dataset.image.reshape(C, H, W, S)
        kcoords = dataset.coords[:, 0:3].reshape(C, H, W, 3)
    return img, kcoords


def partition_kspace(dataset=None, img=None, kcoords=None, show=True, no_steps=40, no_parts=4):
    """Returns (k-means label of every initial ring, radii separating the partitions) -- reference :19-92.
    `show` is accepted and ignored (plotting is outside this engine's scope)."""
    from sklearn.cluster import KMeans
    img, kcoords = _arrays(dataset, img, kcoords)
    dist_to_center = torch.sqrt(kcoords[..., 1] ** 2 + kcoords[..., 2] ** 2)
    means = ring_log_max(img, dist_to_center, no_steps).cpu().numpy().astype(np.float64).reshape(-1, 1)
    if not np.isfinite(means).all():
        raise ValueError("a ring holds no sample (or only zeros): the reference fails on this input as well")
    kmeans = KMeans(init="random", n_clusters=no_parts, n_init=10, max_iter=200, random_state=42)
    kmeans.fit(means)
    labels = kmeans.labels_
    return labels, radii_from_labels(labels, no_parts)


def partition_and_stats(dataset=None, img=None, kcoords=None, show=True, no_steps=40, no_parts=4, stat="max"):
    """Returns (max / min of |img| per partition, radii) -- reference :94-134 (|.| is the elementwise abs of the
    (re, im) pairs there, not the complex magnitude; kept)."""
    img, kcoords = _arrays(dataset, img, kcoords)
    _, radii = partition_kspace(None, img, kcoords, show, no_steps, no_parts)
    dist_to_center = torch.sqrt(kcoords[..., 1] ** 2 + kcoords[..., 2] ** 2)
    stats = []
    for i in range(len(radii) - 1):
        sel = (dist_to_center >= radii[i]) & (dist_to_center <= radii[i + 1])
        v = torch.abs(img[sel])
        stats.append(v.min() if stat == "min" else v.max())
    return torch.stack(stats), radii
