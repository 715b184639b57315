#!/usr/bin/env python
"""Randomised parity run on a GPU box: random reads, random synthetic models of random spec types, random block sizes,
both container modes, per-read / per-lane selection with several models per type, sub-chunked host-pointer calls, and the
text path through the host mirror with random chunking -- every container compared with the CPU oracle byte for byte and
decoded back.  The checker is oracle/ (test infrastructure); the code under test is reached through the C-ABI only.

    python tools/fuzz_gpu.py --seconds 240 --seed 1 > gpurun_out/fuzz.log

Prints one line per failing case (with the seed that reproduces it) and a summary; exit status 1 if anything failed."""
import argparse
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import oracle as O  # noqa: E402  (checker)
from idencomp_b200 import capi, host as H  # noqa: E402
from gpu_util import SPEC_NAMES, blocks_of, synthetic_model, upload  # noqa: E402

COMPAT, NATIVE = 1, 2


def random_reads(rng):
    n = int(rng.integers(1, 160))
    kind = rng.integers(0, 4)
    seqs = []
    qvals = rng.choice(94, size=int(rng.integers(2, 9)), replace=False)
    for i in range(n):
        if kind == 0:
            ln = int(rng.integers(0, 130))
        elif kind == 1:
            ln = int(rng.choice([0, 1, 2, 7, 8, 9, 100, 101, 255, 256, 257]))
        elif kind == 2:
            ln = int(rng.integers(80, 120)) if rng.random() < 0.9 else int(rng.integers(600, 5000))
        else:
            ln = 100
        a = rng.choice(5, size=ln, p=[0.04, 0.3, 0.22, 0.22, 0.22])
        q = rng.choice(qvals, size=ln)
        seqs.append((b"r%d/%d" % (i, int(rng.integers(0, 1000))) if rng.random() < 0.9 else b"", a, q))
    return O.Reads.from_lists(seqs)


def check(out, block_off, crc, expect):
    for b, e in enumerate(expect):
        data, ecrc = e[0], e[1]
        blk = out[int(block_off[b]):int(block_off[b + 1])].tobytes()
        assert int.from_bytes(blk[0:4], "big") == len(blk) - 8 == len(data), f"block {b} length"
        assert int.from_bytes(blk[4:8], "big") == ecrc == int(crc[b]), f"block {b} crc"
        assert blk[8:] == data, f"block {b} bytes differ"


def one_case(gctx, seed, log):
    rng = np.random.default_rng(seed)
    reads = random_reads(rng)
    max_len = int(np.max(np.diff(reads.read_off.astype(np.int64)))) if reads.n_reads else 0
    n_a, n_q = int(rng.integers(1, 4)), int(rng.integers(1, 4))
    models = []
    for k in range(n_a):
        models.append(synthetic_model(O, O.ACID, str(rng.choice(SPEC_NAMES)), reads, seed * 10 + k, n_ctx=int(rng.integers(1, 5))))
    for k in range(n_q):
        models.append(synthetic_model(O, O.QSCORE, str(rng.choice(SPEC_NAMES)), reads, seed * 10 + 5 + k, n_ctx=int(rng.integers(1, 5))))
    ids = [m.md.identifier for m in models]
    if len(set(ids)) != len(ids):
        return "skip"
    if rng.random() < 0.3:
        perm = rng.permutation(len(models))
        models = [models[i] for i in perm]
    handles = [upload(gctx, O, m) for m in models]
    try:
        block_len = int(max(2 * max_len, rng.choice([200, 1000, 5000, 60000])))
        bf = blocks_of(reads, block_len)
        nb = len(bf) - 1
        fast = bool(rng.random() < 0.4) and n_a == 1 and n_q == 1  # fast mode takes exactly one model per type (compressor_block.rs:96)
        gctx.set_pipeline_blocks(int(rng.integers(1, 6)))
        what = f"seed {seed}: {reads.n_reads} reads, max len {max_len}, {n_a}+{n_q} models {[m.md.spec_name for m in models]}, {nb} blocks of {block_len}, fast {fast}"
        # ---- compat
        out, boff, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, handles, fast=fast)
        expect = [O.compress_block(models, reads, int(bf[b]), int(bf[b + 1] - bf[b]), include_identifiers=False, fast=fast) for b in range(nb)]
        check(out, boff, crc, expect)
        doff = np.append(boff[:-1] + 8, boff[-1]).astype(np.uint64)
        dlen = (boff[1:] - boff[:-1] - 8).astype(np.uint32)
        ro, a, q = gctx.decompress_blocks(out, doff, crc, handles, block_len=dlen, mode=COMPAT)
        assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals), "compat decode"
        # ---- native
        lane_syms = int(rng.choice([1, 64, 256, 300, 1024, 2048]))
        gctx.set_lane_symbols(lane_syms)
        out, boff, crc, _ = gctx.compress_blocks(reads.read_off, reads.acids, reads.quals, bf, handles, mode=NATIVE, fast=fast)
        expect = [O.compress_native_block(models, reads, int(bf[b]), int(bf[b + 1] - bf[b]), lane_syms=lane_syms, include_identifiers=False,
                                          fast=fast) for b in range(nb)]
        check(out, boff, crc, expect)
        doff = np.append(boff[:-1] + 8, boff[-1]).astype(np.uint64)
        dlen = (boff[1:] - boff[:-1] - 8).astype(np.uint32)
        ro, a, q = gctx.decompress_blocks(out, doff, crc, handles, block_len=dlen, mode=NATIVE)
        assert np.array_equal(ro, reads.read_off) and np.array_equal(a, reads.acids) and np.array_equal(q, reads.quals), "native decode"
        # ---- text path through the host mirror (one model per type: the file-level subset selection is not under test here)
        if reads.n_reads and rng.random() < 0.5:
            pair = [next(m for m in models if m.md.mtype == O.ACID), next(m for m in models if m.md.mtype == O.QSCORE)]
            hm = [H.Model.new(m.md.mtype, m.md.spec_name, m.md.probs, m.md.spec_keys, m.md.spec_ctx) for m in pair]
            text = O.fastq_write(reads)
            names = bool(rng.random() < 0.5)
            mode = int(rng.choice([COMPAT, NATIVE]))
            kw = dict(max_block_total_len=block_len, include_identifiers=names, mode=mode, batch_blocks=int(rng.integers(1, 6)),
                      lane_symbols=max(lane_syms, 2))
            c = H.IdnCompressor(hm, **kw)
            c.add_batch(reads.read_off, reads.acids, reads.quals, reads.name_off if names else None, reads.names if names else None)
            want = c.finish()
            c.close()
            cuts = sorted(rng.integers(1, len(text), size=int(rng.integers(0, 5))).tolist())
            c = H.IdnCompressor(hm, text_chunk_bytes=int(rng.choice([1 << 30, 3 * block_len, 20000])), devices=[0, 0] if rng.random() < 0.5 else [0], **kw)
            pos = 0
            for cut in cuts + [len(text)]:
                c.add_fastq_text(text[pos:cut])
                pos = cut
            got = c.finish()
            c.close()
            assert got == want, f"text in != sequences in (names {names}, mode {mode}, cuts {cuts})"
            back = H.decompress_text(hm, got, batch_blocks=int(rng.integers(1, 6)))
            if names:
                assert back == text, "text out"
            else:
                p = O.fastq_parse(back)
                assert np.array_equal(p.acids, reads.acids) and np.array_equal(p.quals, reads.quals), "text out (no names)"
        return "ok"
    except AssertionError as e:
        log(f"FAIL {what}: {e}")
        return "fail"
    except Exception as e:  # noqa: BLE001
        log(f"ERROR {what}: {type(e).__name__}: {e}\n{traceback.format_exc(limit=3)}")
        return "fail"
    finally:
        gctx.set_lane_symbols(2048)
        gctx.set_pipeline_blocks(32)
        for h in handles:
            gctx.release_model(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=0, help="stop after this many cases (0: run for --seconds)")
    args = ap.parse_args()
    gctx = capi.Context(0)
    t0 = time.time()
    counts = {"ok": 0, "fail": 0, "skip": 0}
    seed = args.seed

    def log(s):
        print(s, flush=True)
    while (time.time() - t0 < args.seconds) and (args.cases == 0 or sum(counts.values()) < args.cases):
        counts[one_case(gctx, seed, log)] += 1
        seed += 1
    gctx.close()
    H.release_cached()
    print(f"fuzz: {counts['ok']} ok, {counts['fail']} failed, {counts['skip']} skipped, seeds {args.seed}..{seed - 1}, {time.time() - t0:.0f} s", flush=True)
    sys.exit(1 if counts["fail"] else 0)


if __name__ == "__main__":
    main()
