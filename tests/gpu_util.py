"""Helpers shared by the GPU parity tests: feed oracle-side model data through the C-ABI."""
import re

import numpy as np


def parse_spec(name: str):
    """spec type name -> (kind, ao, qo, pb, qmax)   (idencomp-macros/src/lib.rs:166-196)"""
    if name == "dummy":
        return 0, 0, 0, 0, 0
    m = re.fullmatch(r"generic_ao(\d+)_qo(\d+)_pb(\d+)", name)
    if m:
        return (0, *map(int, m.groups()), 0)
    m = re.fullmatch(r"light_ao(\d+)_qo(\d+)_pb(\d+)_qm(\d+)", name)
    if m:
        return (1, *map(int, m.groups()))
    raise ValueError(name)


def upload(gctx, O, model) -> int:
    """Upload an oracle Model (its integer tables come from the pinned quantiser) through idn_gpu_model_upload."""
    md = model.md
    kind, ao, qo, pb, qmax = parse_spec(md.spec_name)
    return gctx.upload_model(md.mtype, kind, ao, qo, pb, qmax, model.cum_table(), md.spec_keys, md.spec_ctx)


def blocks_of(reads, max_block_total_len):
    """IdnCompressor::add_sequence block forming (idn/compressor.rs:517-540): block_first_read."""
    first = [0]
    cur = 0
    for r in range(reads.n_reads):
        ln = int(reads.read_off[r + 1] - reads.read_off[r])
        if cur + ln > max_block_total_len and cur > 0:
            first.append(r)
            cur = 0
        cur += ln
    first.append(reads.n_reads)
    return np.asarray(first, dtype=np.uint32)
