"""All ranks of a world emulated on ONE GPU (LocalFabric) over a BASELINE config: one warm-up + one measured run of the
distributed stages; used under ncu for a per-kernel launch list of the multi-GPU path (the kernels of all ranks run
back to back on the one device, so per-kernel sums / WORLD = one rank's share).
usage: python tools/profile_dist.py CONFIG SCALE WORLD"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402
from mauvealigner_b200.dist import LocalFabric, concat_results, find_unique  # noqa: E402


def main():
    config, scale, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    pattern = mb.get_seed(15, 0) if config == 1 else mb.get_seed(15, mb.CODING_SEED)
    seqs = mb.synth_genomes(config, scale)
    dev = torch.device("cuda", 0)
    ctxs = [mb.Context(0) for _ in range(world)]
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        for c in ctxs:
            c.set_stream(stream.cuda_stream)
            for s in seqs:
                c.add_sequence(s)
            c.set_seed(pattern)
        fabric = LocalFabric(world)
        for _ in range(2):
            info = find_unique(ctxs, fabric, dev)
        stream.synchronize()
        res = concat_results([c.fetch() for c in ctxs])
    print(res["n_matches"], [c.dist_stage_ms() for c in ctxs][:1], info[0])


if __name__ == "__main__":
    main()
