"""Parity tests proper: the CUDA path, called through the C-ABI (libidn_gpu.so), against the CPU oracle and the
reference's golden container.  Everything here is integer / byte work, so the bar is BIT-EXACT.

Run on the B200 box: python -m pytest tests -m gpu
"""
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from gpu_util import blocks_of, upload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gctx():
    from idencomp_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def toy_handles(gctx, O, toy_models):
    return [upload(gctx, O, m) for m in toy_models]


@pytest.fixture(scope="module")
def bundled(gctx, O, model_data):
    """name -> (oracle Model, device handle) for every bundled model."""
    out = {}
    for name, md in model_data.items():
        m = O.Model(md)
        out[name] = (m, upload(gctx, O, m))
    return out


def oracle_blocks(O, models, reads, block_first, *, fast, names):
    """Expected bytes per block: slices only (compressor_block.rs:83-120) + crc."""
    out = []
    for b in range(len(block_first) - 1):
        data, crc, stats = O.compress_block(models, reads, int(block_first[b]), int(block_first[b + 1] - block_first[b]),
                                            include_identifiers=names, fast=fast)
        out.append((data, crc, stats))
    return out


def check_container(out, block_off, crc, expect, prefix=None):
    for b, (data, ecrc, _) in enumerate(expect):
        lo, hi = int(block_off[b]), int(block_off[b + 1])
        blk = out[lo:hi].tobytes()
        plen = 0 if prefix is None else int(prefix[b])
        assert int.from_bytes(blk[0:4], "big") == len(blk) - 8 == len(data), f"block {b} length"
        assert int.from_bytes(blk[4:8], "big") == ecrc == int(crc[b]), f"block {b} crc"
        assert blk[8 + plen:] == data[plen:], f"block {b} slices differ"


# ---- the golden container (idencomp/tests/simple_ctx.rs:19-32), encode side --------------------------------------
def test_encode_1m_matches_golden_container(gctx, toy_handles, reads_1m):
    golden = (GOLDEN / "1M.idn").read_bytes()
    block = golden[76:76 + 8 + 538705]  # header + the one data block
    names_slice = block[8:8 + 37]
    out, block_off, crc, stats = gctx.compress_blocks(reads_1m.read_off, reads_1m.acids, reads_1m.quals, [0, 1],
                                                      toy_handles, prefix_len=[37], name_off=reads_1m.name_off,
                                                      names=reads_1m.names)
    out = bytearray(out.tobytes())
    out[8:8 + 37] = names_slice  # names are compressed on the host (as in the reference); the device reserves room
    assert bytes(out) == block
    assert stats["payload_bytes"] == 538655 and stats["acid_switches"] == 1 and stats["q_switches"] == 1
    assert int(crc[0]) == 0xC1F69A94


def test_decode_1m_golden_container(gctx, toy_handles, reads_1m):
    golden = (GOLDEN / "1M.idn").read_bytes()
    payload = np.frombuffer(golden[84:84 + 538705], dtype=np.uint8)
    crc = int.from_bytes(golden[80:84], "big")
    ro, a, q = gctx.decompress_blocks(payload, [0, len(payload)], [crc], toy_handles, name_off=reads_1m.name_off,
                                      names=reads_1m.names)
    assert ro.tolist() == [0, 500000]
    assert np.array_equal(a, reads_1m.acids) and np.array_equal(q, reads_1m.quals)


def test_score_1m(gctx, toy_handles, reads_1m):
    sizes = gctx.score((reads_1m.read_off, reads_1m.acids, reads_1m.quals), toy_handles)
    assert sizes.tolist() == [[188868, 349787]]


# ---- 1k reads against the oracle, several block shapes ------------------------------------------------------------
@pytest.mark.parametrize("block_len", [4 * 1024 * 1024, 10000, 200])
@pytest.mark.parametrize("fast", [False, True])
def test_encode_1k_vs_oracle(gctx, O, toy_models, toy_handles, reads_1k, block_len, fast):
    bf = blocks_of(reads_1k, block_len)
    expect = oracle_blocks(O, toy_models, reads_1k, bf, fast=fast, names=False)
    out, block_off, crc, stats = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles,
                                                      fast=fast)
    check_container(out, block_off, crc, expect)
    assert stats["out_bytes"] == int(block_off[-1]) == sum(len(d) + 8 for d, _, _ in expect)


def test_encode_1k_with_names_and_prefix(gctx, O, toy_models, toy_handles, reads_1k):
    bf = blocks_of(reads_1k, 5000)
    expect = oracle_blocks(O, toy_models, reads_1k, bf, fast=False, names=True)
    prefix = [6 + int.from_bytes(d[1:5], "big") for d, _, _ in expect]  # Identifiers slice: 00 | u32 len | u8 | data
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles,
                                                  prefix_len=prefix, name_off=reads_1k.name_off, names=reads_1k.names)
    check_container(out, block_off, crc, expect, prefix)


def _decode_all(gctx, O, models, handles, reads, bf, fast):
    """compress on the device, decompress on the device, compare with the input; also feed the device output to the
    oracle decoder and the oracle output to the device decoder."""
    out, block_off, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, handles, fast=fast)
    nb = len(bf) - 1
    payload = np.concatenate([out[int(block_off[b]) + 8:int(block_off[b + 1])] for b in range(nb)]) if nb else out[:0]
    sizes = [int(block_off[b + 1] - block_off[b]) - 8 for b in range(nb)]
    poff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    ro, a, q = gctx.decompress_blocks(payload, poff, crc, handles)
    assert np.array_equal(ro, reads.read_off)
    assert np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)
    return out, block_off, crc


@pytest.mark.parametrize("block_len", [4 * 1024 * 1024, 200])
@pytest.mark.parametrize("fast", [False, True])
def test_round_trip_1k_device(gctx, O, toy_models, toy_handles, reads_1k, block_len, fast):
    _decode_all(gctx, O, toy_models, toy_handles, reads_1k, blocks_of(reads_1k, block_len), fast)


def test_device_output_decodes_with_oracle(gctx, O, toy_models, toy_handles, reads_1k):
    """A whole .idn file assembled from device blocks is read back by the oracle's file decoder."""
    bf = blocks_of(reads_1k, 20000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
    ids = b"".join(m.md.identifier for m in toy_models)
    idn = b"IDENCOMP\x01" + bytes([1, 0, len(toy_models)]) + ids + out.tobytes() + b"\x00" * 8
    back = O.decompress(toy_models, idn)
    assert np.array_equal(back.read_off, reads_1k.read_off)
    assert np.array_equal(back.acids, reads_1k.acids) and np.array_equal(back.quals, reads_1k.quals)
    # and the oracle's own file is bit-identical outside the names slices
    ref = O.compress(toy_models, reads_1k, max_block_total_len=20000, include_identifiers=False)
    assert ref == idn


# ---- every bundled model pair: dense maps, the sparse (hashed) 2^27-spec model, all spec kinds --------------------
def test_bundled_pairs_vs_oracle(gctx, O, bundled, reads_1k):
    acids = [k for k, (m, _) in bundled.items() if m.mtype == O.ACID]
    quals = [k for k, (m, _) in bundled.items() if m.mtype == O.QSCORE]
    bf = blocks_of(reads_1k, 30000)
    for i in range(max(len(acids), len(quals))):
        an, qn = acids[i % len(acids)], quals[i % len(quals)]
        models = [bundled[an][0], bundled[qn][0]]
        handles = [bundled[an][1], bundled[qn][1]]
        expect = oracle_blocks(O, models, reads_1k, bf, fast=True, names=False)
        out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, handles,
                                                      fast=True)
        check_container(out, block_off, crc, expect)
        _decode_all(gctx, O, models, handles, reads_1k, bf, True)


def test_scores_vs_oracle_all_models(gctx, O, bundled, reads_1k):
    names = list(bundled)
    handles = [bundled[n][1] for n in names]
    n = 120
    ro = reads_1k.read_off[:n + 1]
    sizes = gctx.score((ro, reads_1k.acids[:int(ro[-1])], reads_1k.quals[:int(ro[-1])]), handles)
    for j, name in enumerate(names):
        m = bundled[name][0]
        for r in range(n):
            lo, hi = int(ro[r]), int(ro[r + 1])
            assert int(sizes[r, j]) == O.score_read(m, reads_1k.acids[lo:hi], reads_1k.quals[lo:hi]), (name, r)


def test_model_selection_vs_oracle(gctx, O, bundled, reads_1k):
    """Non-fast mode with several candidates per type: scoring + greedy switching (model_chooser.rs:168-198,
    compressor_block.rs:232-280) must reproduce the oracle's switch slices byte for byte."""
    acids = [k for k, (m, _) in bundled.items() if m.mtype == O.ACID][:4]
    quals = [k for k, (m, _) in bundled.items() if m.mtype == O.QSCORE][:4]
    order = acids + quals  # acid ids first, then q ids (compressor_initializer.rs:57-65)
    models = [bundled[k][0] for k in order]
    handles = [bundled[k][1] for k in order]
    bf = blocks_of(reads_1k, 9000)
    expect = oracle_blocks(O, models, reads_1k, bf, fast=False, names=False)
    out, block_off, crc, stats = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, handles)
    check_container(out, block_off, crc, expect)
    assert stats["acid_switches"] == sum(s["acid_switches"] for _, _, s in expect)
    assert stats["q_switches"] == sum(s["q_switches"] for _, _, s in expect)
    _decode_all(gctx, O, models, handles, reads_1k, bf, False)


def test_interleaved_model_order(gctx, O, bundled, reads_1k):
    """SwitchModel indices are positions in the provider list whatever the type order."""
    a = [k for k, (m, _) in bundled.items() if m.mtype == O.ACID][:2]
    q = [k for k, (m, _) in bundled.items() if m.mtype == O.QSCORE][:2]
    order = [q[0], a[0], q[1], a[1]]
    models = [bundled[k][0] for k in order]
    handles = [bundled[k][1] for k in order]
    bf = blocks_of(reads_1k, 40000)
    expect = oracle_blocks(O, models, reads_1k, bf, fast=False, names=False)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, handles)
    check_container(out, block_off, crc, expect)


# ---- edge cases ----------------------------------------------------------------------------------------------------
def test_empty_and_ragged_reads(gctx, O):
    models = [O.Model(O.ModelData.empty(O.ACID)), O.Model(O.ModelData.empty(O.QSCORE))]
    handles = [upload(gctx, O, m) for m in models]
    rng = np.random.default_rng(7)
    seqs = [("", [], [])]
    for ln in [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 0, 64, 100, 1, 0, 255, 1000]:
        seqs.append(("", rng.integers(0, 5, ln), rng.integers(0, 94, ln)))
    reads = O.Reads.from_lists(seqs)
    bf = np.asarray([0, 1, 1, 5, reads.n_reads], dtype=np.uint32)  # includes an empty block and a block of one empty read
    expect = oracle_blocks(O, models, reads, bf, fast=False, names=False)
    out, block_off, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, handles)
    check_container(out, block_off, crc, expect)
    _decode_all(gctx, O, models, handles, reads, bf, False)


def test_zero_reads_batch(gctx, toy_handles):
    out, block_off, crc, stats = gctx.compress_blocks([0], [], [], [0, 0], toy_handles)
    assert out.tobytes() == b"\x00" * 8 and block_off.tolist() == [0, 8] and int(crc[0]) == 0


def test_random_symbols_all_spec_types(gctx, O, bundled):
    """Uniformly random symbols mostly fall through to the dummy context (sequence_compressor.rs:37-40) and hit every
    quality value / N handling path of the generic and light generators."""
    rng = np.random.default_rng(11)
    seqs = []
    for ln in rng.integers(1, 400, 300):
        seqs.append(("", rng.integers(0, 5, ln), rng.integers(0, 94, ln)))
    reads = O.Reads.from_lists(seqs)
    bf = blocks_of(reads, 10000)
    acids = [k for k, (m, _) in bundled.items() if m.mtype == O.ACID]
    quals = [k for k, (m, _) in bundled.items() if m.mtype == O.QSCORE]
    for i in range(max(len(acids), len(quals))):
        an, qn = acids[i % len(acids)], quals[(i + 3) % len(quals)]
        models = [bundled[an][0], bundled[qn][0]]
        handles = [bundled[an][1], bundled[qn][1]]
        expect = oracle_blocks(O, models, reads, bf, fast=True, names=False)
        out, block_off, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, handles, fast=True)
        check_container(out, block_off, crc, expect)
        _decode_all(gctx, O, models, handles, reads, bf, True)


def test_block_crc_matches_zlib(gctx, reads_1k):
    bf = blocks_of(reads_1k, 7000)
    got = gctx.block_crc(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, reads_1k.name_off, reads_1k.names)
    for b in range(len(bf) - 1):
        c = 0
        for r in range(int(bf[b]), int(bf[b + 1])):
            lo, hi = int(reads_1k.read_off[r]), int(reads_1k.read_off[r + 1])
            c = zlib.crc32(reads_1k.name(r), c)
            c = zlib.crc32(reads_1k.acids[lo:hi].tobytes(), c)
            c = zlib.crc32(reads_1k.quals[lo:hi].tobytes(), c)
        assert int(got[b]) == c


# ---- error behaviour (IdnCompressorError / IdnDecompressorError variants) -----------------------------------------
def test_errors(gctx, O, toy_models, toy_handles, reads_1k):
    from idencomp_b200.capi import IdnGpuError
    bf = blocks_of(reads_1k, 20000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
    nb = len(bf) - 1
    payload = np.concatenate([out[int(block_off[b]) + 8:int(block_off[b + 1])] for b in range(nb)])
    sizes = [int(block_off[b + 1] - block_off[b]) - 8 for b in range(nb)]
    poff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    # checksum mismatch names the block
    bad_crc = crc.copy()
    bad_crc[1] ^= 1
    with pytest.raises(IdnGpuError) as e:
        gctx.decompress_blocks(payload, poff, bad_crc, toy_handles)
    assert e.value.kind == "BlockChecksumMismatch" and e.value.bad_block == 1
    # sequence slice before any SwitchModel
    with pytest.raises(IdnGpuError) as e:
        gctx.decompress_blocks(payload[4:int(poff[1])], [0, int(poff[1]) - 4], None, toy_handles)
    assert e.value.kind == "NoActiveModel"
    # SwitchModel index out of range
    broken = payload[:int(poff[1])].copy()
    broken[1] = 9
    with pytest.raises(IdnGpuError) as e:
        gctx.decompress_blocks(broken, [0, len(broken)], None, toy_handles)
    assert e.value.kind == "InvalidModelIndex"
    # truncated block
    with pytest.raises(IdnGpuError) as e:
        gctx.decompress_blocks(payload[:int(poff[1]) - 3], [0, int(poff[1]) - 3], None, toy_handles)
    assert e.value.kind == "SerializeError"
    # output capacity
    with pytest.raises(IdnGpuError) as e:
        gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, out_cap=1000)
    assert e.value.kind == "NoSpace" and e.value.stats["required_bytes"] == int(block_off[-1])
    # invalid symbols
    bad = reads_1k.acids.copy()
    bad[5] = 7
    with pytest.raises(IdnGpuError) as e:
        gctx.compress_blocks(reads_1k.read_off, bad, reads_1k.quals, bf, toy_handles)
    assert e.value.kind == "InvalidSymbol"
    # fast mode with != 2 models, unknown handle
    with pytest.raises(IdnGpuError) as e:
        gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles + toy_handles[:1], fast=True)
    assert e.value.kind == "InvalidState"
    with pytest.raises(IdnGpuError) as e:
        gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, [999, 998])
    assert e.value.kind == "UnknownModel"


# ---- workload generator + container chunks passed as they lie on disk ------------------------------------------
def test_synth_reads_match_oracle_sampler(gctx, O, bundled):
    """idn_gpu_synth_reads_dev draws the same reads as the oracle's sampler (same SplitMix64 stream, same slot ->
    symbol search), so the CPU baseline and the GPU path of bench.py are fed identical inputs."""
    import ctypes as C

    import torch
    am, ha = bundled["ERR174310__human__illumina_hiseq_2000__acids"]
    qm, hq = bundled["SRR2962693__human__illumina_hiseq_2500__q_scores"]
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 300, size=3000).astype(np.uint64)
    ro = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=ro[1:])
    want = O.synth_reads(am, qm, ro, 1234, 20240601, 500)
    S = int(ro[-1])
    ro_d = torch.from_numpy(ro.view(np.int64)).cuda()
    a_d = torch.zeros(S + 16, dtype=torch.uint8, device="cuda")
    q_d = torch.zeros(S + 16, dtype=torch.uint8, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    gctx.check(gctx.L.idn_gpu_synth_reads_dev(gctx.h, ha, hq, ro_d.data_ptr(), len(lens), 1234, 20240601, 500,
                                              a_d.data_ptr(), q_d.data_ptr(), sp))
    torch.cuda.synchronize()
    assert np.array_equal(a_d[:S].cpu().numpy(), want.acids)
    assert np.array_equal(q_d[:S].cpu().numpy(), want.quals)


def test_decode_container_chunk_in_place(gctx, O, toy_models, toy_handles, reads_1k):
    """block_len lets the caller pass container bytes with the 8-byte block headers left in between the payloads."""
    bf = blocks_of(reads_1k, 10000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
    nb = len(bf) - 1
    doff = np.append(block_off[:-1] + 8, block_off[-1]).astype(np.uint64)
    dlen = (block_off[1:] - block_off[:-1] - 8).astype(np.uint32)
    ro, a, q = gctx.decompress_blocks(out, doff, crc, toy_handles, block_len=dlen)
    assert np.array_equal(ro, reads_1k.read_off) and np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)
    # a block that claims to run past the input is a malformed container, not an out-of-bounds read
    from idencomp_b200.capi import IdnGpuError
    bad = dlen.copy()
    bad[nb - 1] += 64
    with pytest.raises(IdnGpuError) as e:
        gctx.decompress_blocks(out, doff, crc, toy_handles, block_len=bad)
    assert e.value.kind == "SerializeError"


def test_unaligned_batch_pointers(gctx, O, toy_models, toy_handles, reads_1k):
    """Device entry points accept symbol arrays at any byte alignment (chunks of a larger resident batch)."""
    import ctypes as C

    import torch
    from idencomp_b200 import capi
    S = int(reads_1k.read_off[-1])
    for shift in (1, 2, 3):
        a_d = torch.zeros(S + 32, dtype=torch.uint8, device="cuda")
        q_d = torch.zeros(S + 32, dtype=torch.uint8, device="cuda")
        a_d[shift:shift + S] = torch.from_numpy(reads_1k.acids).cuda()
        q_d[shift + 4:shift + 4 + S] = torch.from_numpy(reads_1k.quals).cuda()
        ro_d = torch.from_numpy(reads_1k.read_off.view(np.int64)).cuda()
        bf = blocks_of(reads_1k, 20000)
        bf_d = torch.from_numpy(bf.astype(np.int32)).cuda()
        nb = len(bf) - 1
        cap = int(gctx.L.idn_gpu_compress_bound(reads_1k.n_reads, S, nb, 0))
        out_d = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")
        boff_d = torch.zeros(nb + 1, dtype=torch.int64, device="cuda")
        crc_d = torch.zeros(nb, dtype=torch.int32, device="cuda")
        st_d = torch.zeros(8, dtype=torch.int64, device="cuda")
        b = capi.Batch()
        b.n_reads, b.n_symbols, b.n_blocks = reads_1k.n_reads, S, nb
        b.acids, b.quals = a_d.data_ptr() + shift, q_d.data_ptr() + shift + 4
        b.read_off, b.block_first_read = ro_d.data_ptr(), bf_d.data_ptr()
        hd = np.asarray(toy_handles, dtype=np.int32)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        gctx.check(gctx.L.idn_gpu_compress_blocks_dev(gctx.h, C.byref(b), capi.MODE_COMPAT, hd.ctypes.data, 2, 0, None,
                                                      out_d.data_ptr(), cap, boff_d.data_ptr(), crc_d.data_ptr(),
                                                      st_d.data_ptr(), sp))
        torch.cuda.synchronize()
        n = int(st_d[0])
        got = out_d[:n].cpu().numpy().tobytes()
        ref = O.compress(toy_models, reads_1k, max_block_total_len=20000, include_identifiers=False)
        assert got == ref[9 + 3 + 64:-8]


def test_walk_skips_large_identifier_slices_and_long_reads(gctx, O, toy_models, toy_handles):
    """The slice walk stages blocks through shared memory in 8 KiB tiles: cover slices much larger than a tile
    (a names slice of ~40 KiB, reads of 30 000 symbols) next to tiny ones, in several blocks."""
    rng = np.random.default_rng(11)
    seqs = []
    for i, ln in enumerate([30000, 3, 0, 17000, 1, 250, 9000, 64, 64, 12000]):
        name = bytes(rng.integers(33, 127, size=4000 + i, dtype=np.uint8))
        seqs.append((name, rng.integers(0, 5, size=ln), rng.integers(0, 94, size=ln)))
    reads = O.Reads.from_lists(seqs)
    idn = O.compress(toy_models, reads, max_block_total_len=60000, include_identifiers=True)
    # split the container by hand: header 9 + metadata 3 + 2 ids
    pos, blocks = 9 + 3 + 64, []
    while True:
        ln = int.from_bytes(idn[pos:pos + 4], "big")
        crc = int.from_bytes(idn[pos + 4:pos + 8], "big")
        if ln == 0:
            break
        blocks.append((pos + 8, ln, crc))
        pos += 8 + ln
    assert len(blocks) >= 2
    buf = np.frombuffer(idn, dtype=np.uint8)
    doff = np.asarray([b[0] for b in blocks] + [len(idn)], dtype=np.uint64)
    dlen = np.asarray([b[1] for b in blocks], dtype=np.uint32)
    crc = np.asarray([b[2] for b in blocks], dtype=np.uint32)
    ro, a, q = gctx.decompress_blocks(buf, doff, crc, toy_handles, block_len=dlen, name_off=reads.name_off, names=reads.names)
    assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)


# ---- size-independent properties at a size the oracle would need minutes for --------------------------------------
@pytest.mark.parametrize("mode", [1, 2])
def test_large_device_round_trip_properties(gctx, O, bundled, mode):
    """2 M model-driven synthetic reads (200 M symbols, 48 blocks) stay on the device: encode -> decode is the identity,
    the block CRCs of the container equal the CRCs of the input computed independently (idn_gpu_block_crc), the
    container is deterministic, and a spot-checked block equals the oracle's encoding of the same reads."""
    import ctypes as C

    import torch
    from idencomp_b200 import capi
    am, ha = bundled["ERR174310__human__illumina_hiseq_2000__acids"]
    qm, hq = bundled["SRR2962693__human__illumina_hiseq_2500__q_scores"]
    R, L = 2_000_000, 100
    S = R * L
    dev = "cuda"
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ro = torch.arange(R + 1, dtype=torch.int64, device=dev) * L
    a_d = torch.zeros(S + 16, dtype=torch.uint8, device=dev)
    q_d = torch.zeros(S + 16, dtype=torch.uint8, device=dev)
    gctx.check(gctx.L.idn_gpu_synth_reads_dev(gctx.h, ha, hq, ro.data_ptr(), R, 7, 20240601, 500, a_d.data_ptr(), q_d.data_ptr(), sp))
    per = 4 * 1024 * 1024 // L
    bf_h = np.append(np.arange(0, R, per), R).astype(np.int32)
    nb = len(bf_h) - 1
    bf = torch.from_numpy(bf_h).to(dev)
    b = capi.Batch()
    b.n_reads, b.n_symbols, b.n_blocks = R, S, nb
    b.acids, b.quals, b.read_off, b.block_first_read = a_d.data_ptr(), q_d.data_ptr(), ro.data_ptr(), bf.data_ptr()
    hd = np.asarray([ha, hq], dtype=np.int32)
    cap = int(gctx.L.idn_gpu_compress_bound(R, S, nb, 0))
    outs = []
    for _ in range(2):
        out = torch.zeros(cap + 16, dtype=torch.uint8, device=dev)
        boff = torch.zeros(nb + 1, dtype=torch.int64, device=dev)
        crc = torch.zeros(nb, dtype=torch.int32, device=dev)
        st = torch.zeros(8, dtype=torch.int64, device=dev)
        gctx.check(gctx.L.idn_gpu_compress_blocks_dev(gctx.h, C.byref(b), mode, hd.ctypes.data, 2, 0, None, out.data_ptr(), cap,
                                                      boff.data_ptr(), crc.data_ptr(), st.data_ptr(), sp))
        torch.cuda.synchronize()
        outs.append((out, boff, crc, int(st[0])))
    (out, boff, crc, n_out), (out2, _, crc2, n_out2) = outs
    assert n_out == n_out2 and torch.equal(out[:n_out], out2[:n_out])          # deterministic
    # CRC of the input, computed by the stand-alone CRC entry point on host copies of two blocks
    a_h, q_h = a_d[:S].cpu().numpy(), q_d[:S].cpu().numpy()
    lo, hi = int(bf_h[3]) * L, int(bf_h[5]) * L
    sub_ro = np.arange(0, (hi - lo) // L + 1, dtype=np.uint64) * L
    want_crc = gctx.block_crc(sub_ro, a_h[lo:hi], q_h[lo:hi], np.asarray([0, per, 2 * per], dtype=np.uint32))
    assert want_crc.tolist() == (crc[3:5].cpu().numpy().astype(np.uint32)).tolist()
    import zlib
    assert int(want_crc[0]) == zlib.crc32(b"".join(a_h[lo + i * L:lo + (i + 1) * L].tobytes() + q_h[lo + i * L:lo + (i + 1) * L].tobytes()
                                                     for i in range(per)))
    # decode on the device: identity
    doff = torch.cat([boff[:-1] + 8, boff[-1:]]).contiguous()
    dlen = (boff[1:] - boff[:-1] - 8).to(torch.int32).contiguous()
    da = torch.zeros(S + 16, dtype=torch.uint8, device=dev)
    dq = torch.zeros(S + 16, dtype=torch.uint8, device=dev)
    dro = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    stt = torch.zeros(4, dtype=torch.int32, device=dev)
    gctx.check(gctx.L.idn_gpu_decompress_blocks_dev(gctx.h, out.data_ptr(), doff.data_ptr(), dlen.data_ptr(), crc.data_ptr(), nb, n_out, mode,
                                                    hd.ctypes.data, 2, da.data_ptr(), dq.data_ptr(), dro.data_ptr(), R, S, stt.data_ptr(), sp))
    torch.cuda.synchronize()
    assert stt.cpu().tolist()[:3] == [0, -1, R]
    assert torch.equal(da[:S], a_d[:S]) and torch.equal(dq[:S], q_d[:S]) and torch.equal(dro, ro)
    # one block against the oracle, byte for byte
    k = nb - 1  # the ragged last block
    r0, r1 = int(bf_h[k]), int(bf_h[k + 1])
    sub = O.Reads(np.arange(0, r1 - r0 + 1, dtype=np.uint64) * L, a_h[r0 * L:r1 * L], q_h[r0 * L:r1 * L])
    if mode == 1:
        data, ocrc, _ = O.compress_block([am, qm], sub, 0, r1 - r0, include_identifiers=False)
    else:
        data, ocrc = O.compress_native_block([am, qm], sub, 0, r1 - r0, lane_syms=2048, include_identifiers=False)
    lo_b, hi_b = int(boff[k]), int(boff[k + 1])
    assert out[lo_b + 8:hi_b].cpu().numpy().tobytes() == data and (int(crc[k]) & 0xffffffff) == ocrc


def test_walk_speculation_survives_header_lookalikes(gctx, O, toy_models, toy_handles):
    """The slice walk speculates on slice boundaries inside a tile.  Identifier slices full of bytes that look like
    slice headers (02 00 00 00 0c 00 00 00 01 ..., 01 00, 00 00 00 00 03 01) must not derail it: a false boundary is
    dropped when the true chain does not end there."""
    rng = np.random.default_rng(3)
    bait = (b"\x02\x00\x00\x00\x0c\x00\x00\x00\x01" + b"\x01\x00" * 6 + b"\x00\x00\x00\x00\x03\x01abc" + b"\x01\x01") * 40
    seqs = []
    for i in range(400):
        ln = int(rng.integers(20, 120))
        seqs.append((bait[i % 7:i % 7 + 30 + (i % 50)], rng.integers(0, 5, size=ln), rng.integers(0, 94, size=ln)))
    reads = O.Reads.from_lists(seqs)
    idn = O.compress(toy_models, reads, max_block_total_len=12000, include_identifiers=True)
    # store the names slices uncompressed-looking: replace every Identifiers slice by one whose DATA is the bait itself
    # (the device skips identifier data by length, whatever its bytes are)
    pos, parts, blocks = 9 + 3 + 64, [], []
    while True:
        ln = int.from_bytes(idn[pos:pos + 4], "big")
        crc = int.from_bytes(idn[pos + 4:pos + 8], "big")
        if ln == 0:
            break
        body = idn[pos + 8:pos + 8 + ln]
        assert body[0] == 0
        nlen = int.from_bytes(body[1:5], "big")
        fake = bait[:9000]
        new_body = b"\x00" + len(fake).to_bytes(4, "big") + b"\x01" + fake + body[6 + nlen:]
        blocks.append((len(b"".join(parts)) , len(new_body), crc))
        parts.append(new_body)
        pos += 8 + ln
    buf = np.frombuffer(b"".join(parts), dtype=np.uint8)
    doff = np.asarray([b[0] for b in blocks] + [len(buf)], dtype=np.uint64)
    dlen = np.asarray([b[1] for b in blocks], dtype=np.uint32)
    crc = np.asarray([b[2] for b in blocks], dtype=np.uint32)
    ro, a, q = gctx.decompress_blocks(buf, doff, crc, toy_handles, block_len=dlen, name_off=reads.name_off, names=reads.names)
    assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals)


def test_walk_drops_false_starts_that_rejoin_the_chain(gctx, toy_handles):
    """A payload that ends 02 00 00 in front of a Sequence header 02 00 00 00 41 reads, three bytes early, as a Sequence
    slice of 0x200 bytes, and with 74-byte slices that false slice ends exactly on a true header (3 + 7 * 74 = 521): the
    speculative walk's hop test cannot tell such a start from a true one.  The piece before it overshoots it, which is
    how it is found and dropped.  The index must come out as the serial walk's."""
    rng = np.random.default_rng(11)
    blocks, expect = [], []
    for n_reads in (20000, 7000):
        parts = [b"\x01\x00\x01\x01"]
        for i in range(n_reads):
            pay = bytearray(rng.integers(3, 256, size=65, dtype=np.uint8).tobytes())  # no byte looks like a slice type
            if i % 40 == 7:
                pay[-3:] = b"\x02\x00\x00"
            parts.append(b"\x02" + (65).to_bytes(4, "big") + (100 + i % 3).to_bytes(4, "big") + bytes(pay))
        blocks.append(b"".join(parts))
        expect.append((n_reads, sum(100 + i % 3 for i in range(n_reads))))
    buf = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    off = np.asarray([0, len(blocks[0]), len(buf)], dtype=np.uint64)
    n_reads, n_syms, bf = gctx.index_blocks(buf, off, toy_handles)
    assert n_reads == sum(e[0] for e in expect) and n_syms == sum(e[1] for e in expect)
    assert bf.tolist() == [0, expect[0][0], expect[0][0] + expect[1][0]]


def test_both_walks_agree_on_mutated_containers(gctx, O, toy_models, toy_handles, reads_1k):
    """Fuzz: random byte flips, truncations and splices of a valid multi-block container.  The parallel walk (with the
    serial walk as its fallback) and the serial walk alone must give the same answer for every input -- the same index
    totals, or the same error kind -- and neither may fault."""
    from idencomp_b200.capi import IdnGpuError
    bf = blocks_of(reads_1k, 9000)
    out, block_off, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles,
                                                  name_off=reads_1k.name_off, names=reads_1k.names,
                                                  prefix_len=np.full(len(bf) - 1, 40, dtype=np.uint32))
    nb = len(bf) - 1
    doff = np.append(block_off[:-1] + 8, block_off[-1]).astype(np.uint64)
    dlen = (block_off[1:] - block_off[:-1] - 8).astype(np.uint32)
    base = out.copy()
    for b in range(nb):  # a syntactically valid Identifiers slice in the reserved prefix: 00 u32be(34) 01 + 34 bytes
        lo = int(doff[b])
        base[lo:lo + 6] = np.frombuffer(b"\x00" + (34).to_bytes(4, "big") + b"\x01", dtype=np.uint8)
    rng = np.random.default_rng(2024)

    def run(buf, off, ln, mode):
        gctx.set_walk(mode)
        try:
            n_reads, n_syms, first = gctx.index_blocks(buf, off, toy_handles, block_len=ln)
            return ("ok", n_reads, n_syms, first.tolist())
        except IdnGpuError as e:
            return ("err", e.kind)
        finally:
            gctx.set_walk(0)

    n_err = 0
    for it in range(150):
        buf, off, ln = base.copy(), doff.copy(), dlen.copy()
        kind = it % 5
        if kind == 0:    # flip a few bits inside one block (one bad block: the error kind cannot depend on a race between blocks)
            b = int(rng.integers(0, nb))
            for p in rng.integers(int(doff[b]), int(doff[b]) + int(dlen[b]), size=int(rng.integers(1, 4))):
                buf[p] ^= np.uint8(1 << int(rng.integers(0, 8)))
        elif kind == 1:  # corrupt a slice header field of a random read
            b = int(rng.integers(0, nb))
            p = int(doff[b]) + 40 + int(rng.integers(0, 200))
            buf[p] = np.uint8(rng.integers(0, 4))
        elif kind == 2:  # truncate a block
            b = int(rng.integers(0, nb))
            ln[b] -= np.uint32(rng.integers(1, 60))
        elif kind == 3:  # plant slice look-alikes inside a payload
            p = int(rng.integers(int(doff[0]) + 200, len(buf) - 40))
            buf[p:p + 9] = np.frombuffer(b"\x02\x00\x00\x00\x10\x00\x00\x00\x05", dtype=np.uint8)
        else:            # valid as it is
            pass
        a, c = run(buf, off, ln, 2), run(buf, off, ln, 1)
        assert a == c, f"iteration {it} (mutation {kind}): parallel walk {a} vs serial walk {c}"
        n_err += a[0] == "err"
    assert 10 < n_err < 150  # the mutations did exercise both outcomes


# ---- both slice walks ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("walk", ["serial", "fast"])
def test_decode_side_with_either_slice_walk(walk):
    """The library picks the parallel speculative walk for calls of fewer than 512 blocks (every container in this
    file) and the one-warp-per-block walk above that and as the fallback.  IDN_WALK forces one of them for a whole
    process (it is read once), so the decode-side tests run again in a child process under each setting."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    sel = ("test_decode_1m_golden_container or test_round_trip_1k_device or test_errors or test_decode_container_chunk_in_place "
           "or test_walk_skips_large_identifier_slices_and_long_reads or test_walk_speculation_survives_header_lookalikes "
           "or test_model_selection_vs_oracle or test_empty_and_ragged_reads or test_walk_drops_false_starts_that_rejoin_the_chain")
    env = dict(os.environ, IDN_WALK=walk)
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-x", "-q", "-m", "gpu", "-k", sel],
                       env=env, capture_output=True, text=True, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
