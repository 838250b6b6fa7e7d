"""torchrun --nproc-per-node N tools/dist_check.py [CONFIG SCALE]: the multi-GPU path over NCCL against the single-GPU
path of the same library on rank 0 (bit-exact CSR), on a scaled BASELINE config."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mauvealigner_b200 as mb  # noqa: E402
from mauvealigner_b200.dist import TorchFabric, find_unique  # noqa: E402


def main():
    config = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    scale = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(backend="nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    seqs = mb.synth_genomes(config, scale)
    pattern = mb.get_seed(15, 0) if config == 1 else mb.get_seed(15, mb.CODING_SEED)
    ctx = mb.Context(local)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for s in seqs:
        ctx.add_sequence(s)
    ctx.set_seed(pattern)
    fabric = TorchFabric(device=dev)
    info = find_unique([ctx], fabric, dev)
    torch.cuda.synchronize()
    print(f"[rank {rank}] {info[0]} stages {ctx.dist_stage_ms()}", flush=True)
    piece = ctx.fetch()
    pieces = [None] * world
    dist.all_gather_object(pieces, {k: piece[k] for k in ("n_matches", "n_comps", "length", "comp_off", "comp_seq", "comp_start")})
    if rank == 0:
        from mauvealigner_b200.dist import concat_results
        got = concat_results(pieces)
        want = ctx.find(mb.MODE_UNIQUE)
        ok = got["n_matches"] == want["n_matches"] and all(np.array_equal(np.asarray(got[k], dtype=np.int64), np.asarray(want[k], dtype=np.int64))
                                                           for k in ("length", "comp_off", "comp_seq", "comp_start"))
        print(f"DIST_CHECK world={world} config=C{config}/{scale} matches={got['n_matches']} pieces={[p['n_matches'] for p in pieces]} "
              f"{'OK bit-exact vs single-GPU path' if ok else 'MISMATCH'}", flush=True)
    dist.barrier()
    fabric.release_peer_buffers([ctx])
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
