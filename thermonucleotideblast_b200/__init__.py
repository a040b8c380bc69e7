"""thermonucleotideblast_b200 -- B200-native (sm_100a) engine for the tntblast search hot path.

Seed scan of a 2-bit packed target database, NucCruc nearest-neighbour Tm/dG alignment of every
candidate window and amplicon / padlock / probe assembly, behind the C ABI in include/tntb200.h.
The CUDA shared library is required; there is no CPU fallback.
"""
from .engine import (ASSAY_MIPS, ASSAY_PADLOCK, ASSAY_PCR, ASSAY_PROBE, MINUS, OLIGO_F, OLIGO_NONE,  # noqa: F401
                     OLIGO_P, OLIGO_R, PLUS, STRAND_BOTH, STRAND_MINUS, STRAND_PLUS, Assay, Engine,
                     EngineError, FragmentList, Hit, SearchOptions, load_library, search_options)
from .sharding import shard_targets  # noqa: F401

__all__ = ["Engine", "EngineError", "FragmentList", "Assay", "Hit", "SearchOptions", "search_options", "load_library",
           "shard_targets"]
