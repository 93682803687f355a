"""Host-side sharding of a batch over the GPUs of one box.

(map, source) pairs and planner problems are independent, so a batch is cut into
contiguous blocks of the item index, one block per rank, with NO data-path
collective: every rank runs the same single-GPU entry points on its block.  Each rank
only uploads the maps its block refers to (re-indexed).  The only communication is
the optional gather of results / timings on the host side (torch.distributed, NCCL or
gloo)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of n items for `rank` of `world`; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(maps, items, item_map, rank: int, world: int):
    """This rank's share of a batch.

    maps      uint8 (nmaps, ny, nx)
    items     int32 (n, 2) sources or (n, 4) start/end pairs
    item_map  int32 (n,) map index per item, or None (every item uses map 0)
    Returns (local_maps, local_items, local_item_map, (lo, hi)); local_item_map indexes
    local_maps, which holds only the maps the block uses, in ascending global order."""
    items = np.ascontiguousarray(items, dtype=np.int32)
    lo, hi = shard_bounds(len(items), rank, world)
    sel = items[lo:hi]
    if item_map is None:
        return maps[:1], sel, None, (lo, hi)
    im = np.ascontiguousarray(item_map, dtype=np.int32)[lo:hi]
    used, local = np.unique(im, return_inverse=True)
    return (np.ascontiguousarray(maps[used]), sel, local.astype(np.int32).reshape(-1), (lo, hi))


def gather_blocks(local: np.ndarray, n_total: int, dist=None):
    """All ranks' blocks of a per-item result, concatenated in item order (host arrays;
    works with the gloo and nccl backends through all_gather_object).  Without a
    process group the local block is the whole result."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    out = np.concatenate(parts, axis=0)
    if len(out) != n_total:
        raise RuntimeError(f"gathered {len(out)} items, expected {n_total}")
    return out
