"""mauvealigner_b200 — B200-native seed-match anchoring (seed -> multi-MUM) behind the reference's MatchFinder protocol.

Package layout: csrc/ (sm_100a CUDA kernels + the C ABI, built into libmauve_b200.so), _lib.py (ctypes binding),
finder.py (host-side mirror of MatchList / UniqueMatchFinder / SeedMatchEnumerator), seeds.py (seed patterns),
synth.py (synthetic genomes), dist.py (multi-GPU plumbing over torch.distributed)."""
from ._lib import MODE_UNIQUE, MODE_SEED_ENUM, MODE_UNIQUE_COUNT, MODE_PAIRWISE, MODE_REPEAT, MauveError, lib  # noqa: F401
from .finder import (Context, Match, MatchList, SortedMerList, MatchFinder, UniqueMatchFinder, MemHash, MaskedMemHash, PairwiseMatchFinder,  # noqa: F401
                     SeedMatchEnumerator, RepeatHash, ContextPool, find_multi, NO_MATCH, EliminateOverlaps, transposeMatches)
from .seeds import get_seed, default_seed_weight, seed_length, seed_weight, seed_valid, SOLID_SEED, CODING_SEED  # noqa: F401
from .synth import synth_genomes  # noqa: F401
