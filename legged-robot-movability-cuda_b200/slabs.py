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


def dealt_chunks(n_total, rank, world, chunks_per_rank=8):
    """Load-balanced split for units of very unequal cost (body poses: a pose above flat ground
    costs a few cell tests, one next to a cliff thousands of predicates): the index range is cut
    into world * chunks_per_rank contiguous chunks which are dealt to the ranks round-robin.  Still
    no communication — every rank computes the same table.  Returns this rank's [(first, count), ...]
    in ascending order; over all ranks the chunks partition [0, n_total) exactly."""
    if world <= 0 or not (0 <= rank < world) or chunks_per_rank <= 0:
        raise ValueError("bad rank/world/chunks")
    nchunks = int(world) * int(chunks_per_rank)
    out = []
    for c in range(rank, nchunks, world):
        first, count = slab_range(n_total, c, nchunks)
        if count:
            out.append((first, count))
    return out


def strong_slab(n_total, rank, world):
    """Strong scaling: a fixed problem of n_total units cut into `world` contiguous slabs."""
    return slab_range(n_total, rank, world)
