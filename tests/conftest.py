"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs the reference's golden vectors, host logic, C-ABI symbol/loader checks.
`-m gpu`: parity tests proper -- CUDA path (through the C-ABI) vs the oracle, bit-exact.
Nothing here reads /root/reference at run time: goldens live in tests/golden/, models in models/.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
MODELS = ROOT / "models"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def toy_models(O):
    """SIMPLE_MODEL_PROVIDER (_internal_test_data.rs:146-149)."""
    return [O.Model(O.simple_acid_model()), O.Model(O.simple_q_score_model())]


@pytest.fixture(scope="session")
def model_data(O):
    """name -> oracle ModelData for every bundled msgpack file in models/."""
    out = {}
    for p in sorted(MODELS.glob("*.msgpack")):
        out[p.stem] = O.ModelData.load_msgpack(p)
    return out


@pytest.fixture(scope="session")
def reads_1k(O):
    return O.fastq_parse((GOLDEN / "1k-reads.fastq").read_bytes())


@pytest.fixture(scope="session")
def reads_1m(O):
    return O.fastq_parse((GOLDEN / "1M.fastq").read_bytes())
