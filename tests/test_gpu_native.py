"""GPU-native multi-lane format (container version 2): the CUDA encoder/decoder through the C-ABI against the
independent CPU statement of the format in oracle/ (bit-exact), plus lossless round trips.  The format is ours (the
reference has nothing like it): DESIGN.md section 8 states it and the ratio delta against the compat output."""
import numpy as np
import pytest

from gpu_util import blocks_of, upload

pytestmark = pytest.mark.gpu

NATIVE = 2


@pytest.fixture(scope="module")
def gctx():
    from idencomp_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def toy_handles(gctx, O, toy_models):
    return [upload(gctx, O, m) for m in toy_models]


def oracle_native(O, models, reads, bf, lane_syms, names=False, fast=False):
    return [O.compress_native_block(models, reads, int(bf[b]), int(bf[b + 1] - bf[b]), lane_syms=lane_syms,
                                    include_identifiers=names, fast=fast) for b in range(len(bf) - 1)]


def check_blocks(out, block_off, crc, expect, prefix=None):
    for b, (data, ecrc) in enumerate(expect):
        blk = out[int(block_off[b]):int(block_off[b + 1])].tobytes()
        plen = 0 if prefix is None else int(prefix[b])
        assert int.from_bytes(blk[0:4], "big") == len(blk) - 8 == len(data), f"block {b} length"
        assert int.from_bytes(blk[4:8], "big") == ecrc == int(crc[b]), f"block {b} crc"
        assert blk[8 + plen:] == data[plen:], f"block {b} slices differ"


def decode_all(gctx, out, block_off, crc, handles, **kw):
    doff = np.append(block_off[:-1] + 8, block_off[-1]).astype(np.uint64)
    dlen = (block_off[1:] - block_off[:-1] - 8).astype(np.uint32)
    return gctx.decompress_blocks(out, doff, crc, handles, block_len=dlen, mode=NATIVE, **kw)


@pytest.mark.parametrize("block_len,lane_syms", [(4 * 1024 * 1024, 4096), (20000, 1000), (200, 64), (20000, 1)])
def test_native_encode_matches_cpu_statement(gctx, O, toy_models, toy_handles, reads_1k, block_len, lane_syms):
    bf = blocks_of(reads_1k, block_len)
    gctx.set_lane_symbols(lane_syms)
    out, block_off, crc, stats = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, mode=NATIVE)
    check_blocks(out, block_off, crc, oracle_native(O, toy_models, reads_1k, bf, lane_syms))
    ro, a, q = decode_all(gctx, out, block_off, crc, toy_handles)
    assert np.array_equal(ro, reads_1k.read_off) and np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)
    gctx.set_lane_symbols(2048)


def test_native_long_read_golden_input(gctx, O, toy_models, toy_handles, reads_1m):
    """One read of 500 000 symbols (the reference's 1M sample): one lane, width-4 length table absent (constant)."""
    bf = np.asarray([0, 1], dtype=np.uint32)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1m.read_off, reads_1m.acids, reads_1m.quals, bf, toy_handles, mode=NATIVE)
    check_blocks(out, block_off, crc, oracle_native(O, toy_models, reads_1m, bf, 2048))
    ro, a, q = decode_all(gctx, out, block_off, crc, toy_handles)
    assert np.array_equal(a, reads_1m.acids) and np.array_equal(q, reads_1m.quals)


def test_native_ragged_empty_and_wide_lengths(gctx, O, toy_models, toy_handles):
    rng = np.random.default_rng(21)
    lens = [0, 5, 3000, 0, 0, 17, 999, 1, 1000, 1001, 4, 70000, 0, 2, 2, 0]
    seqs = [(f"r{i}".encode(), rng.integers(0, 5, size=l), rng.integers(0, 94, size=l)) for i, l in enumerate(lens)]
    reads = O.Reads.from_lists(seqs)
    for block_len in (10 ** 9, 4000):
        bf = blocks_of(reads, block_len)
        gctx.set_lane_symbols(1000)
        out, block_off, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, toy_handles, mode=NATIVE)
        check_blocks(out, block_off, crc, oracle_native(O, toy_models, reads, bf, 1000))
        ro, a, q = decode_all(gctx, out, block_off, crc, toy_handles)
        assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)
        # the CPU statement decodes the device's blocks too
        for b in range(len(bf) - 1):
            blk = out[int(block_off[b]) + 8:int(block_off[b + 1])].tobytes()
            ln, a2, q2 = O.decompress_native_block(toy_models, blk)
            lo, hi = int(reads.read_off[bf[b]]), int(reads.read_off[bf[b + 1]])
            assert np.array_equal(a2, reads.acids[lo:hi]) and np.array_equal(q2, reads.quals[lo:hi])
    gctx.set_lane_symbols(2048)


def test_native_names_prefix_and_crc(gctx, O, toy_models, toy_handles, reads_1k):
    """Identifiers stay on the host: the device reserves the prefix, the CRC covers the names."""
    bf = blocks_of(reads_1k, 20000)
    expect = oracle_native(O, toy_models, reads_1k, bf, 2048, names=True)
    prefix = []
    for data, _ in expect:
        assert data[0] == 0
        prefix.append(6 + int.from_bytes(data[1:5], "big"))
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles,
                                                  mode=NATIVE, prefix_len=prefix, name_off=reads_1k.name_off, names=reads_1k.names)
    check_blocks(out, block_off, crc, expect, prefix)
    # fill the prefixes as the host would and decode with names
    out = out.copy()
    for b, (data, _) in enumerate(expect):
        lo = int(block_off[b]) + 8
        out[lo:lo + prefix[b]] = np.frombuffer(data[:prefix[b]], dtype=np.uint8)
    ro, a, q = decode_all(gctx, out, block_off, crc, toy_handles, name_off=reads_1k.name_off, names=reads_1k.names)
    assert np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)


def test_native_model_choice_per_lane(gctx, O, model_data, reads_1k):
    names = ["ERR174310__human__illumina_hiseq_2000__acids", "SRR8861483__human__illumina_novaseq_6000__acids",
             "SRR2962693__human__illumina_hiseq_2500__q_scores", "SRR5373739__cat__illumina_hiseq_2500__q_scores",
             "SRR8861483__human__illumina_novaseq_6000__q_scores"]
    models = [O.Model(model_data[n]) for n in names]
    handles = [upload(gctx, O, m) for m in models]
    bf = blocks_of(reads_1k, 30000)
    gctx.set_lane_symbols(2000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, handles, mode=NATIVE)
    check_blocks(out, block_off, crc, oracle_native(O, models, reads_1k, bf, 2000))
    ro, a, q = decode_all(gctx, out, block_off, crc, handles)
    assert np.array_equal(ro, reads_1k.read_off) and np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)
    gctx.set_lane_symbols(2048)
    for h in handles:
        gctx.release_model(h)


def test_native_ratio_delta_vs_compat(gctx, O, toy_models, toy_handles, reads_1k):
    """The stated ratio delta: native must not be larger than compat on the same input (here ~15-20 % smaller)."""
    bf = blocks_of(reads_1k, 4 * 1024 * 1024)
    _, _, _, st_c = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
    _, _, _, st_n = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, mode=NATIVE)
    assert st_n["out_bytes"] < 0.9 * st_c["out_bytes"]


def test_native_rejects_malformed_blocks(gctx, O, toy_models, toy_handles, reads_1k):
    from idencomp_b200.capi import IdnGpuError
    bf = blocks_of(reads_1k, 20000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, mode=NATIVE)
    # flipped payload byte -> checksum mismatch (or a lane that does not end cleanly)
    bad = out.copy()
    bad[int(block_off[1]) - 5] ^= 0x40
    with pytest.raises(IdnGpuError) as e:
        decode_all(gctx, bad, block_off, crc, toy_handles)
    assert e.value.kind in ("BlockChecksumMismatch", "SerializeError")
    # a lane length that no longer adds up -> malformed framing
    bad = out.copy()
    lo = int(block_off[0]) + 8
    n_reads = int.from_bytes(bad[lo + 5:lo + 9].tobytes(), "big")
    n_lanes = int.from_bytes(bad[lo + 9:lo + 13].tobytes(), "big")
    width = int(bad[lo + 17])
    ll0 = lo + 22 + n_reads * width + 2 * n_lanes
    bad[ll0 + 3] ^= 0x01
    with pytest.raises(IdnGpuError) as e:
        decode_all(gctx, bad, block_off, crc, toy_handles)
    assert e.value.kind == "SerializeError"
    # a compat block handed to the native decoder
    outc, boffc, crcc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
    with pytest.raises(IdnGpuError) as e:
        decode_all(gctx, outc, boffc, crcc, toy_handles)
    assert e.value.kind == "SerializeError"


@pytest.mark.parametrize("lane_syms", [256, 1000, 2048])
def test_native_long_reads_are_cut_into_lanes(gctx, O, toy_models, toy_handles, lane_syms):
    """Reads longer than the lane quantum (>= 256) are cut into pieces of lane_syms symbols, each its own lane with the
    8-position history in front of its rANS bytes, so a long read decodes as many short chains: device == CPU statement
    byte for byte, the decode is the identity, the block CRC (per-read, computed by a pass over the output in this case)
    still catches a flipped payload bit, and there are more lanes than reads."""
    from idencomp_b200.capi import IdnGpuError
    rng = np.random.default_rng(33)
    seqs = []
    for i, ln in enumerate([9000, 40, 0, 256, 257, 700, 3, 12345, 90, 90, 1024, 1, 5000, 2048, 4096, 4097, 0]):
        seqs.append((b"r%d" % i, rng.integers(0, 5, size=ln), rng.integers(0, 94, size=ln)))
    reads = O.Reads.from_lists(seqs)
    bf = np.asarray([0, 6, 6, 13, reads.n_reads], dtype=np.uint32)
    gctx.set_lane_symbols(lane_syms)
    try:
        out, block_off, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, toy_handles, mode=NATIVE)
        expect = oracle_native(O, toy_models, reads, bf, lane_syms)
        check_blocks(out, block_off, crc, expect)
        n_lanes = sum(int.from_bytes(d[9:13], "big") for d, _ in expect if d)
        assert n_lanes > reads.n_reads and any(d and d[17] >> 7 for d, _ in expect)
        ro, a, q = decode_all(gctx, out, block_off, crc, toy_handles)
        assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)
        # per-lane model choice with cut reads (two candidates per type)
        m2 = [O.Model(O.acid_model_prefer(1)), O.Model(O.acid_model_prefer(2))]
        mdl = [toy_models[0], m2[0], m2[1], toy_models[1]]
        hd = [toy_handles[0], upload(gctx, O, m2[0]), upload(gctx, O, m2[1]), toy_handles[1]]
        out2, boff2, crc2, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, hd, mode=NATIVE)
        check_blocks(out2, boff2, crc2, oracle_native(O, mdl, reads, bf, lane_syms))
        ro, a, q = decode_all(gctx, out2, boff2, crc2, hd)
        assert np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)
        # a flipped bit in the middle of the payload of the last block
        broken = out.copy()
        broken[int(block_off[-1]) - 40] ^= 0x10
        with pytest.raises(IdnGpuError) as e:
            decode_all(gctx, broken, block_off, crc, toy_handles)
        assert e.value.kind in ("BlockChecksumMismatch", "SerializeError")
    finally:
        gctx.set_lane_symbols(2048)
