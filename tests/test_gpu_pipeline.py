"""The host-pointer C-ABI calls are pipelined inside the call (sub-chunks of whole blocks on upload / compute / download
streams, csrc/idn_pipeline.inc) and the host mirror spreads batches over several devices: neither may change a byte."""
import numpy as np
import pytest

from conftest import MODELS
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


@pytest.mark.parametrize("mode", [1, 2], ids=["compat", "native"])
@pytest.mark.parametrize("sub", [1, 2, 3, 7])
def test_sub_chunking_changes_nothing(gctx, O, toy_models, toy_handles, reads_1k, sub, mode):
    """41 blocks through sub-chunks of 1, 2, 3 and 7 blocks (ragged last sub-chunk) == one sub-chunk, with names (CRC) and
    a reserved prefix per block; and the decode of the result, whose outputs are stitched from the sub-chunks."""
    bf = blocks_of(reads_1k, 1900)
    nb = len(bf) - 1
    assert nb > 30
    prefix = (np.arange(nb) % 5 * 3).astype(np.uint32)
    kw = dict(prefix_len=prefix, name_off=reads_1k.name_off, names=reads_1k.names, mode=mode)
    gctx.set_pipeline_blocks(1000)
    want = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, **kw)
    gctx.set_pipeline_blocks(sub)
    try:
        got = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, **kw)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]) and got[3] == want[3]
        a, b = got[0].copy(), want[0].copy()
        for k in range(nb):  # the reserved bytes behind each block header are the host's to fill (identifiers slice)
            lo = int(want[1][k]) + 8
            a[lo:lo + int(prefix[k])] = 0
            b[lo:lo + int(prefix[k])] = 0
        assert np.array_equal(a, b)
        # decode without names on the device: compress again without them (the CRC then covers the symbols only)
        out, boff, crc, _ = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, mode=mode)
        doff = np.append(boff[:-1] + 8, boff[-1]).astype(np.uint64)
        dlen = (boff[1:] - boff[:-1] - 8).astype(np.uint32)
        ro, a, q = gctx.decompress_blocks(out, doff, crc, toy_handles, block_len=dlen, mode=mode)
        assert np.array_equal(ro, reads_1k.read_off) and np.array_equal(a, reads_1k.acids) and np.array_equal(q, reads_1k.quals)
        # the bytes an index call left on the device decode to the same reads (blocks == NULL), and only right after it
        ro2, a2, q2 = gctx.decompress_blocks(out, doff, crc, toy_handles, block_len=dlen, mode=mode, resident=True)
        assert np.array_equal(ro2, ro) and np.array_equal(a2, a) and np.array_equal(q2, q)
        # errors keep their block numbers across sub-chunks
        from idencomp_b200.capi import IdnGpuError
        bad = crc.copy()
        bad[nb - 2] ^= 5
        with pytest.raises(IdnGpuError) as e:
            gctx.decompress_blocks(out, doff, bad, toy_handles, block_len=dlen, mode=mode)
        assert e.value.kind == "BlockChecksumMismatch" and e.value.bad_block == nb - 2
    finally:
        gctx.set_pipeline_blocks(32)


def test_pipelined_call_reports_nospace_and_bad_symbols(gctx, toy_handles, reads_1k):
    from idencomp_b200.capi import IdnGpuError
    bf = blocks_of(reads_1k, 1900)
    gctx.set_pipeline_blocks(4)
    try:
        full = gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles)
        with pytest.raises(IdnGpuError) as e:
            gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, reads_1k.quals, bf, toy_handles, out_cap=len(full[0]) - 1)
        assert e.value.kind == "NoSpace" and e.value.stats["required_bytes"] == len(full[0])
        bad = reads_1k.quals.copy()
        bad[-3] = 200  # in the last sub-chunk
        with pytest.raises(IdnGpuError) as e:
            gctx.compress_blocks(reads_1k.read_off, reads_1k.acids, bad, bf, toy_handles)
        assert e.value.kind == "InvalidSymbol"
    finally:
        gctx.set_pipeline_blocks(32)


def _host_models(H, O, toy_models):
    return [H.Model.new(m.md.mtype, m.md.spec_name, m.md.probs, m.md.spec_keys, m.md.spec_ctx) for m in toy_models]


def _compress(H, models, reads, devices, mode, names=True, **kw):
    c = H.IdnCompressor(models, max_block_total_len=1500, include_identifiers=names, devices=devices, mode=mode, batch_blocks=3, **kw)
    c.add_batch(reads.read_off, reads.acids, reads.quals, reads.name_off if names else None, reads.names if names else None)
    idn = c.finish()
    c.close()
    return idn


@pytest.mark.parametrize("mode", [1, 2], ids=["compat", "native"])
def test_host_mirror_workers_commit_in_order(O, toy_models, reads_1k, mode):
    """Several workers (here: two and three contexts on device 0) take the batches round robin; the file must be the one a
    single worker writes, and in compat mode the oracle's."""
    from idencomp_b200 import host as H
    hm = _host_models(H, O, toy_models)
    one = _compress(H, hm, reads_1k, [0], mode)
    for devs in ([0, 0], [0, 0, 0]):
        assert _compress(H, hm, reads_1k, devs, mode) == one
    if mode == 1:
        ref = O.compress(toy_models, reads_1k, max_block_total_len=1500, include_identifiers=False)
        assert _compress(H, hm, reads_1k, [0, 0], mode, names=False) == ref
    back = H.decompress(hm, one, device=0, batch_blocks=4)
    assert np.array_equal(back["read_off"], reads_1k.read_off) and np.array_equal(back["acids"], reads_1k.acids)
    assert np.array_equal(back["quals"], reads_1k.quals) and back["names"].tobytes() == reads_1k.names.tobytes()


@pytest.mark.parametrize("mode", [1, 2], ids=["compat", "native"])
def test_n_gpu_container_equals_1_gpu_container(O, toy_models, reads_1k, mode):
    """One file shared by every GPU of the box == the file one GPU writes, byte for byte; decoding shares the file too."""
    from idencomp_b200 import capi, host as H
    n = int(capi.load().idn_gpu_device_count())
    if n < 2:
        pytest.skip("needs at least two GPUs")
    hm = _host_models(H, O, toy_models)
    one = _compress(H, hm, reads_1k, [0], mode)
    many = _compress(H, hm, reads_1k, list(range(n)), mode)
    assert many == one
    # quality 7 over a directory of models: the selection runs on device 0, every device gets the retained set
    stems = sorted(p.stem for p in MODELS.glob("*.msgpack"))[:8]
    dirm = [H.Model.load(MODELS / (s + ".msgpack")) for s in stems]
    a = _compress(H, dirm, reads_1k, [0], mode, quality=7)
    b = _compress(H, dirm, reads_1k, list(range(n)), mode, quality=7)
    assert a == b
    back = H.decompress(dirm, b, n_devices=n, batch_blocks=2)
    assert np.array_equal(back["acids"], reads_1k.acids) and np.array_equal(back["quals"], reads_1k.quals)
    assert back["names"].tobytes() == reads_1k.names.tobytes()
