"""Contiguous sharding of a fragment list over ranks (one process per GPU).

The search path shards by fragment: every (fragment, assay) pair is independent
(tntblast_local.cpp:388-470 hands them out one by one), so ranks need no data-path collective.
A record longer than the fragment length is cut by the caller exactly like the reference does
(seq_len_increment, sequence_data.cpp:739-754, right overlap max_product_length + 2,
tntblast_local.cpp:174,510-511); this module only decides which rank owns which fragment.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_targets(lengths: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Split fragments [0, n) into `world_size` contiguous ranges with balanced base counts.

    Returns [(begin, end)] per rank; ranges are contiguous, ordered and cover every fragment once.
    """
    n = len(lengths)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    total = sum(lengths)
    out = []
    begin = 0
    acc = 0
    for r in range(world_size):
        goal = total * (r + 1) / world_size
        end = begin
        # leave at least one fragment for each remaining rank when possible
        while end < n and (acc + lengths[end] / 2.0 <= goal or end == begin) and (n - end) > (world_size - r - 1):
            acc += lengths[end]
            end += 1
        if r == world_size - 1:
            while end < n:
                acc += lengths[end]
                end += 1
        out.append((begin, end))
        begin = end
    return out


def fragment_record(length: int, max_fragment: int = 500000) -> List[Tuple[int, int]]:
    """Inclusive (start, stop) pieces of one record as the reference driver cuts it
    (seq_len_increment, sequence_data.cpp:739-754 + tntblast_local.cpp:282-289,448-468)."""
    if length <= 0:
        return []
    if max_fragment <= 0 or length <= max_fragment:
        return [(0, length - 1)]
    num = length // max_fragment + (1 if length % max_fragment else 0)
    delta = length // num + (1 if length % num else 0)
    out = []
    start, stop = 0, delta
    max_stop = length - 1
    while True:
        stop = min(stop, max_stop)
        out.append((start, stop))
        if stop == max_stop:
            break
        start = stop + 1
        stop = stop + delta
    return out


def shard_fasta_text(text, world_size: int) -> List[Tuple[int, int]]:
    """Byte ranges [begin, end) of a FASTA text, one per rank, for tnt_engine_add_fasta.

    Every range starts at a '>' that follows a '\\n' (or at the first '>' of the text), i.e. at a point
    where the reference's indexer is in its start state (sequence_data_fastx.cpp:33-58: `read_fasta` is
    false after a '\\n', so that '>' opens a record), hence parsing the ranges one by one yields exactly
    the records of the whole text, in order.  Cuts are the record starts nearest after the equal-size
    byte marks; a rank may get an empty range when there are fewer records than ranks."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if isinstance(text, str):
        data = text.encode("latin-1")
    elif isinstance(text, (bytes, bytearray)):
        data = text
    else:
        data = bytes(text)  # memoryview, numpy uint8 array, anything with the buffer protocol
    n = len(data)
    first = data.find(b">")
    if first < 0:
        return [(0, 0)] * world_size
    cuts = [first]
    for r in range(1, world_size):
        mark = max(cuts[-1], first + (n - first) * r // world_size)
        at = data.find(b"\n>", max(mark - 1, 0))
        cuts.append(n if at < 0 else max(at + 1, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]
