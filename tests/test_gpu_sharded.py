"""BASELINE configs[4] in miniature: one database cut into the reference's fragments, dealt out in
contiguous shards (sharding.shard_targets), every shard searched by its own engine, the hit lists
gathered and finished with tnt_finalize_hits -- against one engine holding the whole database.
(The engines run one after the other on cuda:0 here; bench.py --config 5 runs one per GPU.)"""
import numpy as np
import pytest

import gen

pytestmark = pytest.mark.gpu

RECORD_BP, FRAGMENT_BP, OVERLAP = 5_000_000, 500_000, 2002


def test_sharded_search_equals_single_engine(engine_lib):
    from thermonucleotideblast_b200 import Assay, Engine, search_options
    from thermonucleotideblast_b200.engine import finalize_hits
    from thermonucleotideblast_b200.sharding import fragment_record, shard_targets
    n_records = 2
    assays = gen.config5_assays(40)
    alist = [Assay(i, a[0], a[1], None) for i, a in enumerate(assays)]
    db = np.empty(n_records * RECORD_BP, dtype=np.uint8)
    for r in range(n_records):
        gen.config5_record(r, assays, db[r * RECORD_BP:(r + 1) * RECORD_BP])
    table = []
    for r in range(n_records):
        for (a, b) in fragment_record(RECORD_BP, FRAGMENT_BP):
            table.append((r, a, b, RECORD_BP - 1, min(RECORD_BP, b + 1 + OVERLAP) - a))
    opts = search_options(min_primer_tm=45.0, max_len=2000)

    def block(rows):
        with Engine() as e:
            e.set_assays(alist)
            e.add_targets([db[r * RECORD_BP + a: r * RECORD_BP + a + n] for (r, a, b, ms, n) in rows])
            e.search_raw(opts)
            n, raw, text = e.hit_records()
            return (raw, n, text, [tuple(t) for t in rows])

    key = lambda c: (c.assay_index, c.target_id, c.amp_first, c.amp_last, c.primer_strand, c.forward.tm, c.reverse.tm,
                     c.forward.num_mm, c.reverse.num_mm, c.forward_clamp, c.reverse_clamp)
    single_block = block(table)
    single = finalize_hits([single_block], alist)
    assert single_block[1] > len(single) >= 40          # the overlaps and cuts gave uniquify / the truncation filter work
    for world in (2, 3, 8):
        ranges = shard_targets([t[4] for t in table], world)
        blocks = [block(table[lo:hi]) for (lo, hi) in ranges]
        assert sum(b[1] for b in blocks) == single_block[1]
        many = finalize_hits(blocks, alist)
        assert [key(c) for (_, _, c) in many] == [key(c) for (_, _, c) in single]
