"""CPU: the oracle over every legal context-spec type (context_spec.rs:532-599) with small synthetic models -- the same
models and reads tests/test_gpu_variants.py feeds to the CUDA path.  Checks that the oracle accepts each type, that
encode -> decode is the identity in both container formats and that the scorer's size is the size of a real
single-state encode (model_chooser.rs:215-243)."""
import numpy as np
import pytest

from gpu_util import SPEC_NAMES, parse_spec, synthetic_model, toy_reads


def test_spec_name_list_matches_the_macro_list():
    assert len(SPEC_NAMES) == 50 and len(set(SPEC_NAMES)) == 50 and SPEC_NAMES[0] == "dummy"
    for n in SPEC_NAMES:
        kind, ao, qo, pb, qm = parse_spec(n)
        assert kind in (0, 1) and 0 <= ao <= 8 and 0 <= qo <= 5 and pb <= 8


@pytest.mark.parametrize("i", range(len(SPEC_NAMES)), ids=SPEC_NAMES)
def test_oracle_round_trip_per_spec_type(O, i):
    name_a, name_q = SPEC_NAMES[i], SPEC_NAMES[(i + 7) % len(SPEC_NAMES)]
    reads = toy_reads(O, 100 + i)
    am, qm = synthetic_model(O, O.ACID, name_a, reads, i), synthetic_model(O, O.QSCORE, name_q, reads, 1000 + i)
    assert am.md.n_ctx >= 1 and qm.md.n_ctx >= 1
    idn = O.compress([am, qm], reads, max_block_total_len=1500, include_identifiers=False)
    back = O.decompress([am, qm], idn)
    assert np.array_equal(back.read_off, reads.read_off)
    assert np.array_equal(back.acids, reads.acids) and np.array_equal(back.quals, reads.quals)
    data, crc = O.compress_native_block([am, qm], reads, 0, reads.n_reads, lane_syms=256, include_identifiers=False)
    ln, a, q = O.decompress_native_block([am, qm], data)
    assert np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)
    # the models are actually used: a context row other than the dummy is hit
    hit = 0
    for r in range(1, min(reads.n_reads, 8)):
        s0, s1 = int(reads.read_off[r]), int(reads.read_off[r + 1])
        g = O.Generator(name_a, s1 - s0)
        for k in range(s0, s1):
            hit += am.ctx_for(g.current_context()) != 0
            g.update(int(reads.acids[k]), int(reads.quals[k]))
    assert hit > 0


def test_oracle_native_format_splits_long_reads(O, toy_models):
    """Container version 2: reads longer than the lane quantum (>= 256) are cut into lanes of `lane_syms` symbols, each with
    the 8-position history its generators need (csrc/idn_native.cuh).  CPU statement: encode -> decode is the identity,
    a long read yields several lanes, short reads are laid out as before."""
    rng = np.random.default_rng(21)
    seqs = []
    for ln in [5000, 40, 0, 256, 257, 700, 3, 12000, 90, 90, 1024, 1]:
        seqs.append(("", rng.integers(0, 5, size=ln), rng.integers(0, 94, size=ln)))
    reads = O.Reads.from_lists(seqs)
    for q in (256, 300, 1000, 4096):
        data, crc = O.compress_native_block(toy_models, reads, 0, reads.n_reads, lane_syms=q, include_identifiers=False)
        ln, a, qq = O.decompress_native_block(toy_models, data)
        assert ln.tolist() == [len(s[1]) for s in seqs]
        assert np.array_equal(a, reads.acids) and np.array_equal(qq, reads.quals)
        n_lanes = int.from_bytes(data[9:13], "big")
        split = data[17] >> 7
        assert split == (1 if q < 12000 else 0)
        assert n_lanes >= sum(-(-len(s[1]) // q) for s in seqs if len(s[1]) > q)
    # below the threshold nothing is split (one lane may hold a whole long read)
    data, _ = O.compress_native_block(toy_models, reads, 0, reads.n_reads, lane_syms=64, include_identifiers=False)
    assert data[17] >> 7 == 0
    ln, a, qq = O.decompress_native_block(toy_models, data)
    assert np.array_equal(a, reads.acids) and np.array_equal(qq, reads.quals)
