"""Batch partitioning across the GPUs of one box: contiguous slices, no collective.

Every state is independent, so rank g owns states [g*ceil(n/G), min(n, (g+1)*ceil(n/G))).  The
host scatters inputs and gathers results; NVLink is not on the data path.
"""


def slice_bounds(n, world_size, rank):
    """Half-open range of states owned by `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    per = -(-n // world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def all_slices(n, world_size):
    return [slice_bounds(n, world_size, r) for r in range(world_size)]
