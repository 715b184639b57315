"""FASTQ text -> .idn -> FASTQ text through the host mirror without a symbol round trip over PCIe (row f1 wired into
IdnCompressor / IdnDecompressor): the text is split into records, blocks are formed and the block kernels run on the device
(csrc/idn_textpath.inc).  The container must be the one the add_sequence path writes for the same reads, however the text
is cut into pieces, and the text that comes back must be what the reference's FastqWriter prints."""
import numpy as np
import pytest

from conftest import GOLDEN, MODELS

pytestmark = pytest.mark.gpu


def _host_models(H, toy_models):
    return [H.Model.new(m.md.mtype, m.md.spec_name, m.md.probs, m.md.spec_keys, m.md.spec_ctx) for m in toy_models]


def _by_batch(H, models, reads, **kw):
    c = H.IdnCompressor(models, **kw)
    names = kw.get("include_identifiers", True)
    c.add_batch(reads.read_off, reads.acids, reads.quals, reads.name_off if names else None, reads.names if names else None)
    idn = c.finish()
    c.close()
    return idn


def _by_text(H, models, text, cuts, **kw):
    c = H.IdnCompressor(models, **kw)
    pos = 0
    for cut in list(cuts) + [len(text)]:
        c.add_fastq_text(text[pos:cut])
        pos = cut
    idn = c.finish()
    st = c.stats()
    c.close()
    return idn, st


@pytest.mark.parametrize("mode", [1, 2], ids=["compat", "native"])
@pytest.mark.parametrize("names", [True, False], ids=["names", "no_names"])
def test_text_in_equals_sequences_in(O, toy_models, reads_1k, mode, names):
    from idencomp_b200 import host as H
    hm = _host_models(H, toy_models)
    text = (GOLDEN / "1k-reads.fastq").read_bytes()
    kw = dict(max_block_total_len=3000, include_identifiers=names, mode=mode, batch_blocks=4)
    want = _by_batch(H, hm, reads_1k, **kw)
    rng = np.random.default_rng(4)
    for chunk, cuts in ((1 << 30, []), (20_000, sorted(rng.integers(1, len(text), size=7).tolist())), (9_000, list(range(5000, len(text), 5000)))):
        got, st = _by_text(H, hm, text, cuts, text_chunk_bytes=chunk, **kw)
        assert got == want, f"chunk {chunk}, cuts {cuts[:3]}..."
        assert st["in_reads"] == reads_1k.n_reads and st["in_symbols"] == int(reads_1k.read_off[-1])
    if mode == 1 and not names:
        assert want == O.compress(toy_models, reads_1k, max_block_total_len=3000, include_identifiers=False)
    # two workers on the device: a chunk's kernels overlap the next chunk's upload and parse
    got, _ = _by_text(H, hm, text, [], text_chunk_bytes=15_000, devices=[0, 0], **kw)
    assert got == want


@pytest.mark.parametrize("mode", [1, 2], ids=["compat", "native"])
def test_text_out_is_what_the_reference_writer_prints(O, toy_models, reads_1k, mode):
    from idencomp_b200 import host as H
    hm = _host_models(H, toy_models)
    text = (GOLDEN / "1k-reads.fastq").read_bytes()
    idn, _ = _by_text(H, hm, text, [], max_block_total_len=5000, mode=mode, batch_blocks=3)
    want = O.fastq_write(reads_1k)  # FastqWriter::write_sequence for every read (fastq/writer.rs:190-245)
    assert H.decompress_text(hm, idn, batch_blocks=5) == want
    assert H.decompress_text(hm, idn, batch_blocks=2, thread_num=4) == want
    sep = H.decompress_text(hm, idn, batch_blocks=7, title_with_separator=True)
    lines = sep.split(b"\n")
    assert lines[2] == b"+" + lines[0][1:] and len(sep) == len(want) + sum(len(reads_1k.name(r)) for r in range(reads_1k.n_reads))
    # the streaming reader (pieces in the library's page-locked memory) and the caller-buffer form give the same bytes
    rd = H.FastqTextReader(hm, idn, devices=[0, 0], batch_blocks=2)
    pieces = [bytes(p) for p in rd]
    rd.close()
    assert len(pieces) > 2 and b"".join(pieces) == want
    buf = np.zeros(len(want) + 10, dtype=np.uint8)
    assert H.decompress_text_into(hm, idn, buf, batch_blocks=3) == len(want) and bytes(buf[:len(want)]) == want
    with pytest.raises(H.HostError, match="too small"):
        H.decompress_text_into(hm, idn, buf[:len(want) - 1], batch_blocks=3)
    # the compressor writing into the caller's buffer
    c = H.IdnCompressor(hm, max_block_total_len=5000, mode=mode, batch_blocks=3)
    mine = np.zeros(len(idn) + 8, dtype=np.uint8)
    c.set_output(mine)
    c.add_fastq_text(text)
    c.finish()
    assert bytes(c.output_view()) == idn and bytes(mine[:len(idn)]) == idn
    c.close()
    # without identifiers the titles are empty
    idn2, _ = _by_text(H, hm, text, [], max_block_total_len=5000, mode=mode, include_identifiers=False)
    back = O.fastq_parse(H.decompress_text(hm, idn2))
    assert np.array_equal(back.acids, reads_1k.acids) and np.array_equal(back.quals, reads_1k.quals) and int(back.name_off[-1]) == 0


def test_text_path_errors_and_model_selection(O, toy_models, reads_1k):
    from idencomp_b200 import host as H
    hm = _host_models(H, toy_models)
    text = (GOLDEN / "1k-reads.fastq").read_bytes()
    # a malformed record names the reference's FastqReaderError variant
    bad = text.replace(b"\n+\n", b"\n-\n", 1)
    c = H.IdnCompressor(hm, max_block_total_len=3000)
    with pytest.raises(H.HostError) as e:
        c.add_fastq_text(bad)
        c.finish()
    assert "InvalidFormat" in str(e.value)
    c.close()
    # SequenceTooLong: a read longer than half a block (idn/compressor.rs:542-544)
    c = H.IdnCompressor(hm, max_block_total_len=100)
    with pytest.raises(H.HostError) as e:
        c.add_fastq_text(text)
        c.finish()
    assert e.value.code == 4
    c.close()
    # mixing the two input forms is refused
    c = H.IdnCompressor(hm)
    c.add_fastq_text(text[:1000])
    with pytest.raises(H.HostError):
        c.add_batch(reads_1k.read_off, reads_1k.acids, reads_1k.quals)
    c.close()
    # quality 7 over several models: the first block decides the retained set exactly as on the add_sequence path
    stems = sorted(p.stem for p in MODELS.glob("*.msgpack"))[:8]
    dirm = [H.Model.load(MODELS / (s + ".msgpack")) for s in stems]
    kw = dict(max_block_total_len=6000, quality=7, batch_blocks=2)
    want = _by_batch(H, dirm, reads_1k, **kw)
    got, _ = _by_text(H, dirm, text, [33_333], text_chunk_bytes=40_000, **kw)
    assert got == want
