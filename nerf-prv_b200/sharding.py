"""View / object sharding across the GPUs of one box.

* views (config C3): rank g owns the interleaved views {v : v mod G == g}, so near-pole and near-horizon views (cheap
  vs expensive rays) spread evenly.  After the ray cast one all-gather of the coverage rows gives every rank the full
  table in RANK-MAJOR order; `gathered_view_ids` is the global view id of each gathered row, and the greedy breaks ties
  on that id, so the selected sequence is identical on every rank and identical to the single-GPU run.
* objects (config C4): object o goes to rank o mod G; no collective.
"""
import numpy as np


def shard_view_ids(n_views, rank, world):
    """Global view ids owned by `rank` (interleaved).  Every rank must own the same number of views for the
    all-gather; the caller pads with `pad_view_ids`."""
    return np.arange(rank, n_views, world, dtype=np.uint32)


def views_per_rank(n_views, world):
    return (n_views + world - 1) // world


def pad_view_ids(ids, n_views, rank, world):
    """Pad a rank's id list to views_per_rank with ids >= n_views (empty views: their rows stay zero, gain 0)."""
    per = views_per_rank(n_views, world)
    pad = per - len(ids)
    if pad <= 0:
        return ids
    extra = n_views + rank + world * np.arange(pad, dtype=np.uint32)
    return np.concatenate([ids, extra.astype(np.uint32)])


def gathered_view_ids(n_views, world):
    """Global view id of each row of the all-gathered (rank-major) table."""
    return np.concatenate([pad_view_ids(shard_view_ids(n_views, r, world), n_views, r, world) for r in range(world)])


def shard_objects(n_objects, rank, world):
    return list(range(rank, n_objects, world))
