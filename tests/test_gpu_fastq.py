"""Row f1: FASTQ text <-> symbol arrays on the device against the oracle's restatement of fastq/reader.rs and
fastq/writer.rs, on the reference's sample files and on the reader's corner cases."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gctx():
    from idencomp_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()


def same_as_oracle(gctx, O, text):
    want = O.fastq_parse(text)
    ro, a, q, no, nm = gctx.fastq_parse(text)
    assert np.array_equal(ro, want.read_off) and np.array_equal(a, want.acids) and np.array_equal(q, want.quals)
    assert np.array_equal(no, want.name_off) and np.array_equal(nm, want.names)
    return ro, a, q, no, nm


@pytest.mark.parametrize("name", ["1k-reads.fastq", "1M.fastq"])
def test_parse_reference_samples(gctx, O, name):
    text = (GOLDEN / name).read_bytes()
    ro, a, q, no, nm = same_as_oracle(gctx, O, text)
    # and back: FastqWriter output of the parsed reads == oracle's writer == the sample itself (canonical 4-line records)
    out = gctx.fastq_format(ro, a, q, no, nm)
    assert out == O.fastq_write(O.fastq_parse(text))
    assert out == text


def test_reader_corner_cases(gctx, O):
    cases = [
        b"",                                                      # empty input: no reads
        b"\n\n  \n",                                              # only blank lines
        b"@r1\nACGT\n+\n!!!!",                                    # no trailing newline
        b"\n\n@r1\nACGT\n+\n!!!!\n\n\n@r2 extra words \nN\n+r2\n~\n",  # blank lines before titles, trimmed title, +title
        b"@e\n\n+\n\n@f\nA\n+\n#\n",                              # zero-length read: blank symbol lines are NOT skipped
        b"@\nACGTN\n+\n!#%~I\n",                                  # empty title
    ]
    for text in cases:
        same_as_oracle(gctx, O, text)
    ro, a, q, no, nm = gctx.fastq_parse(cases[3])
    assert nm.tobytes() == b"r1r2 extra words" and a.tolist() == [1, 2, 4, 3, 0] and q.tolist() == [0, 0, 0, 0, 93]
    assert gctx.fastq_format(ro, a, q, no, nm, title_with_separator=True) == b"@r1\nACGT\n+r1\n!!!!\n@r2 extra words\nN\n+r2 extra words\n~\n"


def test_many_records_across_tiles(gctx, O):
    rng = np.random.default_rng(9)
    parts = []
    for i in range(30000):
        ln = int(rng.integers(0, 40))
        parts.append(b"@read.%d len=%d\n" % (i, ln))
        parts.append(bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=ln)) + b"\n+\n")
        parts.append(bytes(rng.integers(33, 127, size=ln, dtype=np.uint8)) + b"\n")
        if i % 97 == 0:
            parts.append(b"\n \t\n")
    text = b"".join(parts)
    same_as_oracle(gctx, O, text)


@pytest.mark.parametrize("text,kind,rec", [
    (b"@a\nACGT\n+\n!!!!\nr2\nAC\n+\n!!\n", "InvalidFormat", 1),              # title without '@'
    (b"@a\nACXT\n+\n!!!!\n", "InvalidAcid", 0),
    (b"@a\nACGT\r\n+\r\n!!!!\r\n", "InvalidAcid", 0),                         # CRLF: '\r' is not an acid (as in the reference)
    (b"@a\nACGT\n-\n!!!!\n", "InvalidFormat", 0),                             # separator must start with '+'
    (b"@a\nACGT\n+\n!! !\n", "InvalidQualityScore", 0),
    (b"@a\nACGT\n+\n!!!\n", "AcidAndQualityScoreLengthMismatch", 0),
    (b"@a\nACGT\n+\n!!!!\n@b\nAC\n", "EofReached", 1),                        # input ends inside a record
])
def test_reader_errors(gctx, O, text, kind, rec):
    from idencomp_b200.capi import IdnGpuError
    with pytest.raises(IdnGpuError) as e:
        gctx.fastq_parse(text)
    assert e.value.kind == "SerializeError" and e.value.fastq_error == kind and e.value.bad_record == rec
    with pytest.raises(O.OracleError):
        O.fastq_parse(text)


def test_text_to_container_on_the_device(gctx, O, toy_models):
    """FASTQ text -> parse_dev -> compress_blocks_dev without the symbols leaving the device."""
    import ctypes as C

    import torch
    from gpu_util import blocks_of, upload
    from idencomp_b200 import capi
    text = (GOLDEN / "1k-reads.fastq").read_bytes()
    reads = O.fastq_parse(text)
    handles = np.asarray([upload(gctx, O, m) for m in toy_models], dtype=np.int32)
    t_d = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    info = capi.FastqInfo()
    gctx.check(gctx.L.idn_gpu_fastq_parse_dev(gctx.h, t_d.data_ptr(), len(text), C.byref(info), sp))
    assert info.n_reads == reads.n_reads and info.n_symbols == len(reads.acids)
    b = capi.Batch()
    gctx.check(gctx.L.idn_gpu_fastq_batch_dev(gctx.h, C.byref(b)))
    bf = blocks_of(reads, 20000)
    bf_d = torch.from_numpy(bf.astype(np.int32)).cuda()
    b.n_blocks, b.block_first_read = len(bf) - 1, bf_d.data_ptr()
    b.names, b.name_off = None, None  # identifiers excluded here (the host would Deflate them)
    cap = int(gctx.L.idn_gpu_compress_bound(b.n_reads, b.n_symbols, b.n_blocks, 0))
    out_d = torch.zeros(cap + 16, dtype=torch.uint8, device="cuda")
    boff_d = torch.zeros(b.n_blocks + 1, dtype=torch.int64, device="cuda")
    crc_d = torch.zeros(b.n_blocks, dtype=torch.int32, device="cuda")
    st_d = torch.zeros(8, dtype=torch.int64, device="cuda")
    gctx.check(gctx.L.idn_gpu_compress_blocks_dev(gctx.h, C.byref(b), capi.MODE_COMPAT, handles.ctypes.data, 2, 0, None,
                                                  out_d.data_ptr(), cap, boff_d.data_ptr(), crc_d.data_ptr(), st_d.data_ptr(), sp))
    torch.cuda.synchronize()
    got = out_d[:int(st_d[0])].cpu().numpy().tobytes()
    want = O.compress(toy_models, reads, max_block_total_len=20000, include_identifiers=False)
    assert got == want[9 + 3 + 64:-8]
