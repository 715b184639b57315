"""Worker of tests/test_shard_gloo.py: one rank of a gloo group (CPU).  The oracle stands in for the block compressor
(the product's needs a GPU); what is under test is the sharding plumbing of idencomp_b200/shard.py."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from gpu_util import blocks_of  # noqa: E402
from idencomp_b200 import shard  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out_path, block_len = sys.argv[1], int(sys.argv[2])
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    models = [O.Model(O.simple_acid_model()), O.Model(O.simple_q_score_model())]
    reads = O.fastq_parse((ROOT / "tests" / "golden" / "1k-reads.fastq").read_bytes())
    bf = blocks_of(reads, block_len)
    n_blocks = len(bf) - 1
    ids = b"".join(m.md.identifier for m in models)
    preamble = b"IDENCOMP\x01" + bytes([1, 0, 2]) + ids

    def compress_range(lo, hi):
        out = bytearray()
        for b in range(lo, hi):
            data, crc, _ = O.compress_block(models, reads, int(bf[b]), int(bf[b + 1] - bf[b]), include_identifiers=True)
            out += len(data).to_bytes(4, "big") + crc.to_bytes(4, "big") + data
        return bytes(out)

    idn = shard.compress_sharded(compress_range, preamble, n_blocks, dist)
    if rank == 0:
        Path(out_path).write_bytes(idn)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
