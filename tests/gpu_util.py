"""Helpers shared by the parity tests: feed oracle-side model data through the C-ABI; synthetic models per spec type."""
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


# ---- every legal context-spec type (the model!{} list, context_spec.rs:532-599): dummy + 23 generic + 26 light ----
GENERIC = [(1, 0, 0), (2, 0, 0), (4, 0, 0), (8, 0, 0), (0, 1, 0), (0, 2, 0), (0, 3, 0), (0, 0, 2), (0, 0, 4), (0, 0, 8), (4, 1, 2),
           (1, 3, 2), (2, 1, 6), (6, 2, 0), (3, 3, 0), (8, 0, 4), (4, 0, 3), (4, 0, 6), (0, 2, 6), (0, 3, 3), (4, 2, 6), (5, 2, 4),
           (3, 3, 4)]
LIGHT = [(4, 1, 2, 16), (8, 1, 2, 16), (8, 0, 0, 1), (0, 3, 3, 8), (0, 3, 3, 16), (0, 4, 3, 8), (0, 4, 3, 16), (0, 4, 0, 8),
         (0, 4, 0, 16), (3, 3, 0, 8), (3, 3, 0, 16), (2, 3, 2, 8), (0, 4, 2, 8), (2, 3, 2, 16), (0, 4, 2, 16), (2, 4, 2, 8),
         (4, 3, 4, 16), (4, 3, 2, 8), (0, 3, 0, 4), (0, 3, 0, 8), (0, 3, 0, 16), (0, 3, 0, 32), (4, 4, 4, 8), (4, 4, 4, 16),
         (5, 4, 4, 16), (3, 5, 4, 16)]
SPEC_NAMES = (["dummy"] + [f"generic_ao{a}_qo{q}_pb{p}" for a, q, p in GENERIC] +
              [f"light_ao{a}_qo{q}_pb{p}_qm{m}" for a, q, p, m in LIGHT])
assert len(SPEC_NAMES) == 50


def toy_reads(O, seed, n=48):
    """short ragged reads with few distinct quality values and some N / q = 0, so that contexts repeat"""
    rng = np.random.default_rng(seed)
    seqs = [("", [], [])]
    for _ in range(n):
        ln = int(rng.integers(1, 120))
        a = rng.choice([0, 1, 2, 3, 4], size=ln, p=[0.04, 0.3, 0.22, 0.22, 0.22])
        q = rng.choice([0, 2, 11, 25, 37, 40, 93], size=ln, p=[0.03, 0.07, 0.2, 0.3, 0.25, 0.1, 0.05])
        seqs.append(("", a, q))
    return O.Reads.from_lists(seqs)


def synthetic_model(O, mtype, spec_name, reads, seed, n_ctx=3):
    """a model of `spec_name` whose contexts own the most frequent specs of `reads` (several specs per context, i.e. binned),
    with skewed probabilities incl. exact zeros (the zero-frequency fix-up of the quantiser)"""
    rng = np.random.default_rng(seed)
    counts = {}
    for r in range(reads.n_reads):
        s0, s1 = int(reads.read_off[r]), int(reads.read_off[r + 1])
        g = O.Generator(spec_name, s1 - s0)
        for i in range(s0, s1):
            sp = g.current_context()
            counts[sp] = counts.get(sp, 0) + 1
            g.update(int(reads.acids[i]), int(reads.quals[i]))
    top = [sp for sp, _ in sorted(counts.items(), key=lambda kv: (-kv[1], kv[0]))[:3 * n_ctx]]
    nsym = 5 if mtype == O.ACID else 94
    ctxs = []
    for k in range(min(n_ctx, len(top))):
        p = rng.dirichlet(np.full(nsym, 0.3)).astype(np.float32)
        p[rng.integers(0, nsym)] = 0.0
        ctxs.append((sorted(top[k::n_ctx]), (p / p.sum()).tolist()))
    return O.Model(O.ModelData.from_contexts(mtype, spec_name, ctxs))


