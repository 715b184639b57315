"""The host mirror (C++: IdnCompressor / IdnDecompressor over the C-ABI) against the reference's own API tests:
idencomp/src/idn/tests.rs:13-85 and idencomp/tests/simple_ctx.rs:5-117, plus byte identity with the oracle's
container wherever the reference's output is deterministic without the clustering RNG."""
import zlib

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

N, A, C, T, G = 0, 1, 2, 3, 4
SIMPLE_ACIDS = [G, A, T, T, T, G, G, G, G, T, T, C, A, A, A, G, C, A, G, T, A, T, C, G, A, T, C, A, A, A, T, A, G, T, A, A, A,
                T, C, C, A, T, T, T, G, T, T, C, A, A, C, T, C, A, C, A, G, T, T, T]
SIMPLE_QUALS = [0, 6, 6, 9, 7, 7, 7, 7, 9, 9, 9, 10, 8, 8, 4, 4, 4, 10, 10, 8, 7, 4, 4, 4, 4, 8, 13, 16, 9, 9, 9, 12, 10, 9, 6,
                6, 8, 8, 9, 9, 20, 20, 34, 34, 37, 29, 29, 29, 29, 29, 29, 34, 34, 34, 34, 34, 34, 34, 21, 20]
SHORT = (b"", [A, C, T, G], [0, 1, 13, 50])                       # SHORT_TEST_SEQUENCE
SIMPLE = (b"SEQ_ID", SIMPLE_ACIDS, SIMPLE_QUALS)                  # SIMPLE_TEST_SEQUENCE
PREFER_A = (b"PREFER_A", [A] * 100, [0] * 100)                    # TEST_SEQUENCE_PREFER_A
PREFER_C = (b"PREFER_C", [C] * 100, [50] * 100)                   # TEST_SEQUENCE_PREFER_C


@pytest.fixture(scope="module")
def H():
    from idencomp_b200 import host
    host.load()
    return host


def hmodel(H, md):
    return H.Model.new(md.mtype, md.spec_name, md.probs, md.spec_keys, md.spec_ctx)


@pytest.fixture(scope="module")
def simple_provider(H, O):
    """SIMPLE_MODEL_PROVIDER (_internal_test_data.rs:146-149)"""
    return [hmodel(H, O.simple_acid_model()), hmodel(H, O.simple_q_score_model())]


@pytest.fixture(scope="module")
def default_provider(H):
    """ModelProvider::default(): one empty model per type"""
    return [H.Model.empty(0), H.Model.empty(1)]


def round_trip(H, provider, seqs_in, seqs_out=None, **params):
    c = H.IdnCompressor(provider, **params)
    for name, acids, quals in seqs_in:
        c.add_sequence(name, acids, quals)
    idn = c.finish()
    d = H.decompress(provider, idn)
    want = seqs_in if seqs_out is None else seqs_out
    assert len(d["read_off"]) - 1 == len(want)
    for i, (name, acids, quals) in enumerate(want):
        lo, hi = int(d["read_off"][i]), int(d["read_off"][i + 1])
        assert d["acids"][lo:hi].tolist() == list(acids) and d["quals"][lo:hi].tolist() == list(quals)
        assert d["names"][int(d["name_off"][i]):int(d["name_off"][i + 1])].tobytes() == name
    return idn, c


# ---- idn/tests.rs ------------------------------------------------------------------------------------------------
def test_round_trip_empty_file(H, default_provider):
    idn, _ = round_trip(H, default_provider, [])
    assert len(idn) == 9 + 3 + 64 + 8  # header, metadata with 2 ids, terminator block


def test_round_trip_short_sequence(H, default_provider):
    round_trip(H, default_provider, [SHORT])


def test_round_trip_sequence_with_name(H, default_provider):
    round_trip(H, default_provider, [SIMPLE])


def test_round_trip_sequence_identifiers_disabled(H, default_provider):
    round_trip(H, default_provider, [SIMPLE], [(b"", SIMPLE[1], SIMPLE[2])], include_identifiers=False)


def test_round_trip_multiple_sequences(H, default_provider):
    round_trip(H, default_provider, [SHORT, SIMPLE])


@pytest.fixture(scope="module")
def prefer_provider(H, O):
    return [hmodel(H, O.acid_model_prefer(A)), hmodel(H, O.acid_model_prefer(C)), H.Model.empty(1)]


def test_round_trip_multiple_models(H, prefer_provider):
    idn, c = round_trip(H, prefer_provider, [PREFER_A, PREFER_C])
    st = c.stats()
    assert st["acid_model_switches"] == 2 and st["q_score_model_switches"] == 1  # both acid models retained, one switch per read
    assert len(c.retained_models()) == 3


@pytest.mark.parametrize("quality", range(1, 10))  # 8-9 switch the identifiers slice to Brotli (compressor_block.rs:146-163)
def test_round_trip_all_quals(H, prefer_provider, quality):
    idn, _ = round_trip(H, prefer_provider, [PREFER_A, PREFER_C], quality=quality)
    n_ids = idn[11]
    pos = 9 + 3 + 32 * n_ids + 8  # first block's first slice
    assert idn[pos] == 0 and idn[pos + 5] == (0 if quality >= 8 else 1)  # IdnIdentifierCompression::{Brotli, Deflate}


# ---- tests/simple_ctx.rs -----------------------------------------------------------------------------------------
def test_decompress_simple_1m(H, simple_provider, reads_1m):
    d = H.decompress(simple_provider, (GOLDEN / "1M.idn").read_bytes())
    assert d["version"] == 1 and len(d["read_off"]) == 2
    assert np.array_equal(d["acids"], reads_1m.acids) and np.array_equal(d["quals"], reads_1m.quals)
    assert d["names"].tobytes() == reads_1m.name(0)


def test_compress_simple_1m_is_the_golden_container(H, simple_provider, reads_1m):
    c = H.IdnCompressor(simple_provider)
    c.add_batch(reads_1m.read_off, reads_1m.acids, reads_1m.quals, reads_1m.name_off, reads_1m.names)
    idn = c.finish()
    golden = (GOLDEN / "1M.idn").read_bytes()
    assert len(idn) > 0
    # identical except (possibly) the bytes of the Deflate stream of the names: compare around the identifiers slice
    def split(b):
        pos = 9 + 3 + 64
        ln = int.from_bytes(b[pos:pos + 4], "big")
        crc = b[pos + 4:pos + 8]
        assert b[pos + 8] == 0
        nlen = int.from_bytes(b[pos + 9:pos + 13], "big")
        names = zlib.decompress(b[pos + 14:pos + 14 + nlen], -15)
        return b[:pos], crc, names, b[pos + 14 + nlen:pos + 8 + ln], b[pos + 8 + ln:]
    assert split(idn) == split(golden)


@pytest.mark.parametrize("params", [{}, {"max_block_total_len": 200}, {"max_block_total_len": 200, "thread_num": 8},
                                    {"max_block_total_len": 200, "batch_blocks": 7}])
def test_round_trip_many_sequences(H, O, simple_provider, toy_models, reads_1k, params):
    c = H.IdnCompressor(simple_provider, **params)
    c.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals, reads_1k.name_off, reads_1k.names)
    idn = c.finish()
    d = H.decompress(simple_provider, idn, batch_blocks=params.get("batch_blocks", 32))
    assert np.array_equal(d["read_off"], reads_1k.read_off) and np.array_equal(d["acids"], reads_1k.acids)
    assert np.array_equal(d["quals"], reads_1k.quals)
    assert np.array_equal(d["name_off"], reads_1k.name_off) and np.array_equal(d["names"], reads_1k.names)
    # the oracle (reference restatement) reads the same file
    back = O.decompress(toy_models, idn)
    assert np.array_equal(back.acids, reads_1k.acids) and np.array_equal(back.names, reads_1k.names)


def test_container_identical_to_oracle_without_identifiers(H, O, simple_provider, toy_models, reads_1k):
    for kw in ({}, {"max_block_total_len": 5000}, {"fast": True}):
        c = H.IdnCompressor(simple_provider, include_identifiers=False, batch_blocks=5, **kw)
        c.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals)
        assert c.finish() == O.compress(toy_models, reads_1k, include_identifiers=False, **kw)


def test_quality_1_ranking_matches_oracle(H, O, model_data, reads_1k):
    """quality 1 = get_model_ranking (deterministic): same retained models, same container bytes as the oracle."""
    names = ["ERR174310__human__illumina_hiseq_2000__acids", "SRR8861483__human__illumina_novaseq_6000__acids",
             "m64187e__sars_cov_2__sequel_ii_e__acids", "SRR2962693__human__illumina_hiseq_2500__q_scores",
             "SRR5373739__cat__illumina_hiseq_2500__q_scores", "SRR8861483__human__illumina_novaseq_6000__q_scores"]
    provider = [hmodel(H, model_data[n]) for n in names]
    omodels = [O.Model(model_data[n]) for n in names]
    c = H.IdnCompressor(provider, quality=1, include_identifiers=False)
    c.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals)
    idn = c.finish()
    assert idn == O.compress(omodels, reads_1k, quality=1, include_identifiers=False)
    assert len(c.retained_models()) == 2
    d = H.decompress(provider, idn)
    assert np.array_equal(d["acids"], reads_1k.acids) and np.array_equal(d["quals"], reads_1k.quals)


def test_native_container_round_trip(H, simple_provider, reads_1k):
    from idencomp_b200 import capi
    c = H.IdnCompressor(simple_provider, mode=capi.MODE_NATIVE, max_block_total_len=20000, lane_symbols=1000)
    c.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals, reads_1k.name_off, reads_1k.names)
    idn = c.finish()
    assert idn[8] == 2  # version byte: the reference decoder refuses it cleanly (InvalidVersion)
    d = H.decompress(simple_provider, idn)
    assert d["version"] == 2
    assert np.array_equal(d["acids"], reads_1k.acids) and np.array_equal(d["quals"], reads_1k.quals)
    assert np.array_equal(d["names"], reads_1k.names)
    compat = H.IdnCompressor(simple_provider, max_block_total_len=20000)
    compat.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals, reads_1k.name_off, reads_1k.names)
    assert len(idn) < len(compat.finish())


# ---- error behaviour (idn/compressor.rs:22-33, idn/decompressor.rs:25-48) ------------------------------------------
def test_errors(H, O, simple_provider, default_provider, toy_models, reads_1k):
    c = H.IdnCompressor(simple_provider, max_block_total_len=100)
    with pytest.raises(H.HostError) as e:
        c.add_sequence(b"x", [A] * 51, [0] * 51)
    assert e.value.kind == "SequenceTooLong"
    c.add_sequence(b"x", [A] * 50, [0] * 50)
    c.finish()
    with pytest.raises(H.HostError) as e:
        c.finish()
    assert e.value.kind == "InvalidState"
    idn = bytearray(O.compress(toy_models, reads_1k))
    with pytest.raises(H.HostError) as e:
        H.decompress(simple_provider, bytes(idn[:8]) + b"\x07" + bytes(idn[9:]))
    assert e.value.kind == "InvalidVersion"
    with pytest.raises(H.HostError) as e:
        H.decompress(default_provider, bytes(idn))
    assert e.value.kind == "UnknownModel"
    flipped = bytearray(idn)
    flipped[76 + 4] ^= 0xFF
    with pytest.raises(H.HostError) as e:
        H.decompress(simple_provider, bytes(flipped))
    assert e.value.kind == "BlockChecksumMismatch"
    with pytest.raises(H.HostError) as e:
        H.decompress(simple_provider, bytes(idn[:200]))
    assert e.value.kind == "IoError"
