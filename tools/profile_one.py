"""One warm-up find + N timed finds of a BASELINE config; used under ncu (launch list / full capture).
usage: python tools/profile_one.py CONFIG SCALE [REPEAT]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402


def main():
    config, scale = int(sys.argv[1]), int(sys.argv[2])
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if config == 1:
        pattern, mode, kw = mb.get_seed(15, 0), mb.MODE_UNIQUE, {}
    elif config in (2, 5):
        pattern, mode, kw = mb.get_seed(15, mb.CODING_SEED), mb.MODE_UNIQUE, {}
    elif config == 3:
        pattern, mode, kw = mb.get_seed(19, 0), mb.MODE_UNIQUE_COUNT, {}
    else:
        pattern, mode, kw = mb.get_seed(15, 0), mb.MODE_SEED_ENUM, dict(min_multi=2, max_multi=500)
    seqs = mb.synth_genomes(config, scale)
    ctx = mb.Context(0)
    for s in seqs:
        ctx.add_sequence(s)
    ctx.set_seed(pattern)
    for _ in range(1 + rep):
        ctx.find_device(mode, **kw)
    r = ctx.fetch()
    st = ctx.stats()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, r["n_matches"])


if __name__ == "__main__":
    main()
