"""Contiguous-slab partitioning of query points / body poses over ranks.

Points and poses are independent units, so multi-GPU is pure data parallelism: rank r owns the
contiguous index range [first, first + count) and nothing is exchanged on the data path.  The only
collectives are the barrier and the max-over-ranks of the elapsed time in the benchmark, and an
optional gather of per-slab summaries."""


def slab_range(n_total, rank, world):
    """Balanced contiguous split: the first n_total % world ranks get one extra unit."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_total), int(world))
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first, count


def weak_slab(points_per_rank, rank, world):
    """Weak scaling: every rank owns `points_per_rank` units of a world x larger problem."""
    return rank * int(points_per_rank), int(points_per_rank)
