"""N > 1 path on CPU: world_size 2 (and 3) gloo groups shard the blocks of a file, exchange sizes, concatenate.  The
result must be byte-identical to the single-process container (no data-path collective)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT
from idencomp_b200 import shard


def test_block_ranges_cover_everything():
    for n in (0, 1, 5, 8, 966):
        for w in (1, 2, 3, 8):
            r = shard.block_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    assert shard.exclusive_offsets([5, 0, 7]) == ([0, 5, 5], 12)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,block_len", [(2, 5000), (3, 200), (2, 4 * 1024 * 1024)])
def test_sharded_container_equals_single_process(O, toy_models, reads_1k, tmp_path, world, block_len):
    out = tmp_path / "sharded.idn"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "_shard_worker.py"), str(out), str(block_len)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    want = O.compress(toy_models, reads_1k, max_block_total_len=block_len)
    assert out.read_bytes() == want
    back = O.decompress(toy_models, out.read_bytes())
    assert np.array_equal(back.acids, reads_1k.acids) and np.array_equal(back.names, reads_1k.names)
