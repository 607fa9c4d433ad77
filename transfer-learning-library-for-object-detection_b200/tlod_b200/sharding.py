"""Image sharding of the RoI / proposal path over ranks.

Every unit of this path belongs to exactly one image (``rois[:, 0]``, per-image proposals and
anchor targets), so the path shards by image with no data-path collective: rank r of G owns
the images ``[r*B/G, (r+1)*B/G)``, its feature maps / RPN outputs / gt boxes stay local, and
the batch index in its RoIs is re-based to the local range.  Only the training loop's gradient
all-reduce crosses ranks (NCCL over NVLink; outside this path).  Pure torch; works on CPU
tensors (the ``gloo`` tests) and on CUDA tensors alike."""
from __future__ import annotations

from typing import Tuple

import torch


def image_range(num_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: sizes differ by at most one image, earlier ranks get the extras."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("rank %d of world %d" % (rank, world))
    base, extra = divmod(int(num_images), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensor: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rows of a per-image tensor (features, RPN scores / deltas, im_info, gt boxes) owned by `rank`."""
    lo, hi = image_range(tensor.size(0), rank, world)
    return tensor[lo:hi]


def shard_rois(rois: torch.Tensor, num_images: int, rank: int, world: int):
    """RoIs (R, 5) whose image index falls in the rank's range, re-based to local image indices.
    Returns (local_rois, index) with ``index`` the positions of the kept rows in `rois` (ascending),
    so results computed on the shard can be scattered back: ``out[index] = local_out``."""
    lo, hi = image_range(num_images, rank, world)
    b = rois[:, 0]
    index = torch.nonzero((b >= lo) & (b < hi)).view(-1)
    local = rois.index_select(0, index).clone()
    local[:, 0] -= lo
    return local, index


def unshard_rows(parts, indices, total_rows: int) -> torch.Tensor:
    """Inverse of shard_rois for per-RoI results gathered from all ranks."""
    first = parts[0]
    out = first.new_zeros((total_rows,) + tuple(first.shape[1:]))
    for p, idx in zip(parts, indices):
        out[idx] = p
    return out
