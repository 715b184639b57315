"""Pins the CPU oracle against every golden vector / KAT the reference's tests hold for the hot path
(SURVEY.md section 4 and 8c).  If these pass the oracle may be used as the checker for the CUDA path."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN, MODELS

N, A, C, T, G = 0, 1, 2, 3, 4


# ---- quantiser: context.rs:337-339, 619-646 -------------------------------------------------------
@pytest.mark.parametrize("probs,bits,expect", [
    ([0.0, 0.0, 0.333, 0.333, 0.334], 8, [0, 1, 2, 86, 170]),
    ([0.25, 0.25, 0.25, 0.25], 4, [0, 4, 8, 12]),
    ([0.05, 0.10, 0.125, 0.125, 0.30, 0.03, 0.07, 0.05, 0.12, 0.03], 10,
     [0, 51, 154, 282, 410, 717, 748, 819, 870, 993]),
    ([0.01, 0.01, 0.49, 0.49], 4, [0, 1, 2, 9]),
])
def test_quantise_kats(O, probs, bits, expect):
    assert O.quantise(probs, bits).tolist() == expect


def test_dummy_context_freqs(O):
    # Context::dummy(5) at 14 bits: freqs 3277,3277,3276,3277,3277
    cum = O.quantise([np.float32(1.0) / np.float32(5)] * 5, 14)
    freqs = np.diff(np.append(cum, 16384)).tolist()
    assert freqs == [3277, 3277, 3276, 3277, 3277]


def test_quantise_all_bundled_contexts_valid(O, model_data):
    # as_integer_cum_freqs asserts: all unique, last < total (context.rs:366-367)
    for name, md in model_data.items():
        m = O.Model(md)
        tab = m.cum_table().astype(np.int64)
        assert (np.diff(tab, axis=1) >= 1).all(), name
        assert (tab[:, 0] == 0).all() and (tab[:, -1] == 16384).all()


# ---- context specs: context_spec.rs:269,454,652-718 ------------------------------------------------
def test_spec_num(O):
    assert O.spec_num("generic_ao1_qo0_pb0") == 8
    assert O.spec_num("generic_ao2_qo1_pb5") == 131072
    assert O.spec_num("light_ao2_qo1_pb5_qm16") == 8192
    assert O.spec_num("dummy") == 1
    assert O.spec_num("generic_ao3_qo3_pb0") == 1 << 27


def test_generic_spec_no_pos(O):
    g = O.Generator("generic_ao2_qo1_pb0", 10)
    g.update(C, 0)
    g.update(G, 92)
    assert g.current_context() == 0xB8E


def test_generic_spec_with_pos(O):
    g = O.Generator("generic_ao2_qo1_pb3", 8)
    for a, q in [(N, 0), (N, 0), (N, 0), (C, 0), (G, 92)]:
        g.update(a, q)
    assert g.current_context() == 0x5C75


def test_generator_position(O):
    g = O.Generator("generic_ao0_qo0_pb2", 7)
    seen = [g.current_context()]
    for _ in range(6):
        g.update(0, 0)
        seen.append(g.current_context())
    assert seen == [0, 0, 1, 1, 2, 2, 3]


def test_light_generator(O):
    g = O.Generator("light_ao2_qo2_pb4_qm16", 8)
    seen = [g.current_context()]
    for a, q in [(A, 0), (N, 0), (A, 93), (A, 93), (C, 93), (C, 93)]:
        g.update(a, q)
        seen.append(g.current_context())
    assert seen == [0x0, 0x2, 0x4, 0xF06, 0xFF08, 0xFF1A, 0xFF5C]


# ---- rANS glue: compressor.rs:224-240, 293-321 ------------------------------------------------------
def test_rans_small_output(O):
    cum = O.quantise([0.001, 0.001, 0.997, 0.001], 16)
    freq = np.diff(np.append(cum, 1 << 16))
    out = O.rans_encode_raw([cum[2]] * 500, [freq[2]] * 500, 1, 16)
    assert len(out) == 4


def test_rans_two_channels_layout(O):
    """compressor.rs:293-321 (round_trip_two_channels): four puts on two interleaved states, then the reference's
    assertions -- the decoder returns the pairs last-in-first-out and channel k of get() is channel k of put() -- checked
    with an independent byte-wise decoder (ryg rans_byte: RansDecInit / RansDecAdvance / renormalisation) written out here."""
    SB, L = 6, 1 << 23
    c1 = O.quantise([0.25] * 4, SB)
    c2 = O.quantise([0.125] * 8, SB)
    assert list(c1) == [0, 16, 32, 48] and list(c2) == list(range(0, 64, 8))
    pairs = [(0, 1), (1, 3), (2, 5), (3, 7)]
    starts, freqs = [], []
    for s1, s2 in pairs:  # put(ctx1, s1, ctx2, s2) = put_at(0, s1) then put_at(1, s2)   (compressor.rs:95-96)
        starts += [c1[s1], c2[s2]]
        freqs += [16, 8]
    out = O.rans_encode_raw(starts, freqs, 2, SB)
    assert len(out) == 10  # two flushed states + one renormalisation byte per state (the 4th put of each overflows 2^31 / 2^6 * freq)
    pos = 0

    def init():
        nonlocal pos
        x = int.from_bytes(out[pos:pos + 4], "little")
        pos += 4
        return x

    def get(x, cum, freq):
        nonlocal pos
        slot = x & ((1 << SB) - 1)
        sym = max(k for k in range(len(cum)) if cum[k] <= slot)
        x = freq * (x >> SB) + slot - int(cum[sym])
        while x < L:
            x = (x << 8) | out[pos]
            pos += 1
        return x, sym

    # flush() writes state 0 then state 1, each in front of what is there: state 1 comes first in memory (compressor.rs:99-106),
    # and the decoder undoes the puts in reverse: channel 1 before channel 0 (compressor.rs:181-193)
    x1, x0 = init(), init()
    got = []
    for _ in pairs:
        x1, s2 = get(x1, c2, 8)
        x0, s1 = get(x0, c1, 16)
        got.append((s1, s2))
    assert got == pairs[::-1]  # LIFO order, channel pairing kept
    assert pos == len(out) and x0 == L and x1 == L  # every byte consumed, both states back at their initial value
    # the two channels are not interchangeable: reading the states the other way round breaks the pairing
    pos = 0
    y0, y1 = init(), init()
    y1, t2 = get(y1, c2, 8)
    y0, t1 = get(y0, c1, 16)
    assert (t1, t2) != pairs[-1]


# ---- model identifiers: model.rs:314, model_serializer.rs:177-189 ------------------------------------
def test_empty_model_identifier(O):
    md = O.ModelData.empty(O.ACID)
    assert md.identifier.hex().startswith("85989ce9")
    assert md.identifier == hashlib.sha3_256(b"\x00dummy").digest()


def test_bundled_model_identifiers(O, model_data):
    assert len(model_data) >= 12  # load_msgpack raises on identifier mismatch


def test_toy_model_identifiers(O):
    # the ids stored in samples/1M.idn metadata
    assert O.simple_acid_model().identifier.hex().startswith("55b681de")
    assert O.simple_q_score_model().identifier.hex().startswith("52a858fa")


# ---- the golden container: idencomp/tests/simple_ctx.rs:19-32 -----------------------------------------
def test_decompress_simple_1m(O, toy_models, reads_1m):
    idn = (GOLDEN / "1M.idn").read_bytes()
    reads, info = O.decompress(toy_models, idn, return_info=True)
    assert reads.n_reads == 1 and info["n_blocks"] == 2
    assert np.array_equal(reads.acids, reads_1m.acids)
    assert np.array_equal(reads.quals, reads_1m.quals)
    assert reads.name(0) == reads_1m.name(0) == b"data/SRR1518133_1.fastq 0:500000"
    assert O.fastq_write(reads) == (GOLDEN / "1M.fastq").read_bytes()


def test_1m_payload_states_and_crc(O, toy_models, reads_1m):
    idn = (GOLDEN / "1M.idn").read_bytes()
    assert len(idn) == 538797
    # preamble 76 B, block header 8 B, names slice 6+31 B, two switches 4 B, sequence header 9 B
    payload = idn[76 + 8 + 37 + 4 + 9:len(idn) - 8]
    assert len(payload) == 538655
    a, q, states, used = O.decode_read(toy_models[0], toy_models[1], payload, 500000)
    assert states == (0x800000, 0x800000) and used == 538655
    crc = O.crc32(reads_1m.name(0))
    crc = O.crc32(a, crc)
    crc = O.crc32(q, crc)
    assert crc == 0xC1F69A94 == int.from_bytes(idn[80:84], "big")


def test_compress_simple_1m_bit_exact(O, toy_models, reads_1m):
    # the reference only asserts non-empty (tests/simple_ctx.rs:5-17); the golden bytes are reproducible
    out = O.compress(toy_models, reads_1m)
    assert out == (GOLDEN / "1M.idn").read_bytes()


def test_scorer_values_1m(O, toy_models, reads_1m):
    assert O.score_read(toy_models[0], reads_1m.acids, reads_1m.quals) == 188868
    assert O.score_read(toy_models[1], reads_1m.acids, reads_1m.quals) == 349787


# ---- round trips: tests/simple_ctx.rs:34-117, idn/tests.rs:13-85 ---------------------------------------
def _same(a, b, names=True):
    assert a.n_reads == b.n_reads
    assert np.array_equal(a.read_off, b.read_off)
    assert np.array_equal(a.acids, b.acids) and np.array_equal(a.quals, b.quals)
    if names:
        assert [a.name(i) for i in range(a.n_reads)] == [b.name(i) for i in range(b.n_reads)]


@pytest.mark.parametrize("block,threads", [(4 * 1024 * 1024, 0), (200, 0), (200, 8)])
def test_round_trip_1k_reads(O, toy_models, reads_1k, block, threads):
    assert reads_1k.n_reads == 1000
    idn = O.compress(toy_models, reads_1k, max_block_total_len=block, threads=threads)
    back, info = O.decompress(toy_models, idn, threads=threads, return_info=True)
    _same(back, reads_1k)
    if block == 200:
        assert info["n_blocks"] == 501  # 2 reads of 76 per block + EOF block


def test_round_trip_empty_file(O):
    models = [O.Model(O.ModelData.empty(O.ACID)), O.Model(O.ModelData.empty(O.QSCORE))]
    idn = O.compress(models, O.Reads.from_lists([]))
    assert len(idn) == 9 + 3 + 64 + 8
    assert O.decompress(models, idn).n_reads == 0


def test_round_trip_short_and_named(O):
    models = [O.Model(O.ModelData.empty(O.ACID)), O.Model(O.ModelData.empty(O.QSCORE))]
    short = ("", [A, C, T, G], [0, 1, 13, 50])
    simple = ("SEQ_ID", [G, A, T, T, T, G, G, G, G, T], [0, 6, 6, 9, 7, 7, 7, 7, 9, 9])
    empty = ("seq", [], [])
    reads = O.Reads.from_lists([short, simple, empty])
    back = O.decompress(models, O.compress(models, reads))
    _same(back, reads)
    back = O.decompress(models, O.compress(models, reads, include_identifiers=False))
    _same(back, reads, names=False)
    assert all(back.name(i) == b"" for i in range(3))


@pytest.mark.parametrize("quality", range(1, 8))
def test_round_trip_multiple_models_all_quals(O, quality):
    # idn/tests.rs:53-85: two acid models + empty q model, sequences preferring each
    models = [O.Model(O.acid_model_prefer(A)), O.Model(O.acid_model_prefer(C)),
              O.Model(O.ModelData.empty(O.QSCORE))]
    reads = O.Reads.from_lists([("PREFER_A", [A] * 100, [0] * 100), ("PREFER_C", [C] * 100, [50] * 100)])
    idn, stats = O.compress(models, reads, quality=quality, return_stats=True)
    back = O.decompress(models, idn)
    _same(back, reads)
    if quality >= 2:
        assert stats["acid_switches"] == 2  # both models retained, one switch per read


def test_error_paths(O, toy_models, reads_1k):
    idn = bytearray(O.compress(toy_models, reads_1k))
    bad = bytes(idn[:8]) + b"\x02" + bytes(idn[9:])
    with pytest.raises(O.OracleError) as e:
        O.decompress(toy_models, bad)
    assert e.value.kind == "InvalidVersion"
    with pytest.raises(O.OracleError) as e:
        O.decompress([toy_models[0]], bytes(idn))
    assert e.value.kind == "UnknownModel"
    flipped = bytearray(idn)
    flipped[76 + 4] ^= 0xFF  # block crc
    with pytest.raises(O.OracleError) as e:
        O.decompress(toy_models, bytes(flipped))
    assert e.value.kind == "BlockChecksumMismatch"
    with pytest.raises(O.OracleError) as e:
        O.compress(toy_models, reads_1k, max_block_total_len=100)
    assert e.value.kind == "SequenceTooLong"


def test_round_trip_bundled_pairs(O, model_data, reads_1k):
    # every bundled (acid, q) pair round-trips 1k-reads; unseen specs fall through to the dummy context
    acids = [k for k, v in model_data.items() if v.mtype == O.ACID]
    quals = [k for k, v in model_data.items() if v.mtype == O.QSCORE]
    built = {k: O.Model(v) for k, v in model_data.items()}
    for i, a in enumerate(acids):
        qn = quals[i % len(quals)]
        models = [built[a], built[qn]]
        idn = O.compress(models, reads_1k, fast=True, include_identifiers=False)
        _same(O.decompress(models, idn), reads_1k, names=False)


# ---- our own version-2 block format: the CPU statement round-trips (no reference golden exists for it) ----------
@pytest.mark.parametrize("lane_syms", [1, 64, 1000, 4096])
def test_native_cpu_statement_round_trip(O, toy_models, reads_1k, lane_syms):
    data, crc = O.compress_native_block(toy_models, reads_1k, 0, reads_1k.n_reads, lane_syms=lane_syms, include_identifiers=False)
    ln, a, q = O.decompress_native_block(toy_models, data)
    assert np.array_equal(np.cumsum(ln), reads_1k.read_off[1:])
    assert np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)
    compat, crc_c, _ = O.compress_block(toy_models, reads_1k, 0, reads_1k.n_reads, include_identifiers=False)
    assert crc == crc_c  # same checksum definition as version 1
    if lane_syms >= 1000:
        assert len(data) < len(compat)  # the point of the format: no per-read header and flush
