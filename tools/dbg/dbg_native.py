import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ctypes as C
import numpy as np, torch
from idencomp_b200 import capi
from oracle import oracle as O
from gpu_util import blocks_of, upload
ctx = capi.Context(0)
models = [O.Model(O.simple_acid_model()), O.Model(O.simple_q_score_model())]
h = np.asarray([upload(ctx, O, m) for m in models], dtype=np.int32)
reads = O.fastq_parse(open("tests/golden/1k-reads.fastq","rb").read())
for nr, Q in ((10, 4096), (10, 100), (300, 1000), (1000, 4096)):
    ro = reads.read_off[:nr+1]; S = int(ro[-1])
    a, q = reads.acids[:S], reads.quals[:S]
    bf = np.asarray([0, nr], dtype=np.uint32)
    ctx.set_lane_symbols(Q)
    out, boff, crc, st = ctx.compress_blocks(ro, a, q, bf, list(h), mode=2)
    exp, _ = O.compress_native_block(models, O.Reads(ro, a, q), 0, nr, lane_syms=Q, include_identifiers=False)
    print("encode equal:", out[8:].tobytes() == exp, len(exp))
    ln, a3, q3 = O.decompress_native_block(models, out[8:].tobytes())
    print("oracle decodes gpu bytes:", np.array_equal(a3, a), np.array_equal(q3, q))
    dev = "cuda"
    blk = torch.from_numpy(out.copy()).to(dev)
    doff = torch.tensor([8, len(out)], dtype=torch.int64, device=dev)
    dlen = torch.tensor([len(out) - 8], dtype=torch.int32, device=dev)
    ao = torch.zeros(S + 64, dtype=torch.uint8, device=dev); qo = torch.zeros(S + 64, dtype=torch.uint8, device=dev)
    roo = torch.zeros(nr + 2, dtype=torch.int64, device=dev); stt = torch.zeros(4, dtype=torch.int32, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = ctx.L.idn_gpu_decompress_blocks_dev(ctx.h, blk.data_ptr(), doff.data_ptr(), dlen.data_ptr(), None, 1, len(out), 2, h.ctypes.data, 2,
                                             ao.data_ptr(), qo.data_ptr(), roo.data_ptr(), nr, S, stt.data_ptr(), sp)
    torch.cuda.synchronize()
    print("rc", rc, "status", stt.cpu().tolist())
    print("ro ok", np.array_equal(roo.cpu().numpy()[:nr+1], ro.astype(np.int64)))
    a2 = ao[:S].cpu().numpy(); q2 = qo[:S].cpu().numpy()
    print("acid match prefix:", int(np.argmax(a2 != a)) if (a2 != a).any() else "all", "q:", int(np.argmax(q2 != q)) if (q2 != q).any() else "all")
    bad = np.nonzero((a2 != a) | (q2 != q))[0]
    print("n bad", len(bad), bad[:10], "reads", (bad[:10] // 76))
