"""Host mirror (libidn_host.so, C++): model loading, identifiers and the f32 quantiser, against the reference's
known answers and against the oracle.  CPU only -- no compute call goes to the GPU library here."""
import numpy as np
import pytest

from conftest import MODELS


@pytest.fixture(scope="module")
def H():
    from idencomp_b200 import host
    host.load()
    return host


def test_libraries_export_every_declared_symbol(H):
    import re
    from conftest import ROOT
    from idencomp_b200 import capi
    L = capi.load()
    declared = set(re.findall(r"\b(idn_gpu_\w+)\s*\(", (ROOT / "include" / "idn_gpu.h").read_text()))
    assert declared == set(capi.EXPORTS)
    assert not [s for s in declared if not hasattr(L, s)]
    LH = H.load()
    declared_h = set(re.findall(r"\b(idn_host_\w+)\s*\(", (ROOT / "include" / "idn_host.h").read_text()))
    assert declared_h == set(H.EXPORTS)
    assert not [s for s in declared_h if not hasattr(LH, s)]


# context.rs:337-339, 619-646
@pytest.mark.parametrize("probs,bits,expect", [
    ([0.0, 0.0, 0.333, 0.333, 0.334], 8, [0, 1, 2, 86, 170]),
    ([0.25, 0.25, 0.25, 0.25], 4, [0, 4, 8, 12]),
    ([0.05, 0.10, 0.125, 0.125, 0.30, 0.03, 0.07, 0.05, 0.12, 0.03], 10, [0, 51, 154, 282, 410, 717, 748, 819, 870, 993]),
    ([0.01, 0.01, 0.49, 0.49], 4, [0, 1, 2, 9]),
])
def test_quantise_kats(H, probs, bits, expect):
    assert H.quantise(probs, bits).tolist() == expect


def test_empty_model_identifier(H):
    # model.rs:314: Model::empty(Acids) identifier starts 85989ce9 (SHA3-256 of 00 "dummy")
    m = H.Model.empty(0)
    assert m.identifier.hex().startswith("85989ce9")
    assert len(m) == 0 and m.spec_name == "dummy"


def test_bundled_models_load_and_match_oracle(H, O, model_data):
    # read_model recomputes the SHA3 identifier and refuses a mismatch (model_serializer.rs:111-114)
    for name, md in model_data.items():
        m = H.Model.load(MODELS / (name + ".msgpack"))
        assert m.identifier == md.identifier, name
        assert m.spec_name == md.spec_name and len(m) == md.n_ctx and m.model_type == md.mtype
        assert np.array_equal(m.cum_table(), O.Model(md).cum_table()), name


def test_toy_models_match_reference_ids(H, O):
    # SIMPLE_MODEL_PROVIDER ids in samples/1M.idn's metadata
    for md, prefix in ((O.simple_acid_model(), "55b681de"), (O.simple_q_score_model(), "52a858fa")):
        m = H.Model.new(md.mtype, md.spec_name, md.probs, md.spec_keys, md.spec_ctx)
        assert m.identifier.hex().startswith(prefix)


def test_corrupt_model_is_rejected(H):
    data = bytearray((MODELS / "SRR20210997__salmonella__illumina_hiseq_2500__q_scores.msgpack").read_bytes())
    data[len(data) // 2] ^= 0x01
    with pytest.raises(H.HostError):
        H.Model.from_bytes(bytes(data))
    with pytest.raises(H.HostError):
        H.Model.load(MODELS / "does-not-exist.msgpack")


# ---- file-level model subset selection: clustering.rs:232-271 (the reference's own test of this code) ------------
def _point_costs(points, centroids):
    return [[(p[0] - c[0]) ** 2 + (p[1] - c[1]) ** 2 for c in centroids] for p in points]


def test_cluster_trivial(H):
    cent, vc = H.cluster(_point_costs([(0, 0)], [(2, 1), (-2, 2), (0, 0), (3, -3)]), 1)
    assert cent == [2] and vc == [0]


def test_cluster_points(H):
    points = [(2, 2), (2, 3), (4, 1), (-1, 1), (-2, 1), (-3, 2), (-2, -2), (2, -2), (2, -3)]
    centroids = [(-6, -7), (0, 0), (2, 1), (-2, 2), (-1, -1), (3, -3)]
    cent, vc = H.cluster(_point_costs(points, centroids), 4)
    clusters = sorted((c, [v for v, k in enumerate(vc) if k == i]) for i, c in enumerate(cent))
    assert clusters == [(2, [0, 1, 2]), (3, [3, 4, 5]), (4, [6]), (5, [7, 8])]


def test_rank_models(H):
    # read 0 prefers model 2, reads 1-2 prefer model 0; ties keep provider order (stable sorts, model_chooser.rs:114-136)
    cost = [[5, 9, 1], [1, 9, 5], [1, 9, 5], [7, 7, 7]]
    assert H.rank(cost, 1) == [0]
    assert H.rank(cost, 2) == [0, 2]
    assert H.rank(cost, 5) == [0, 2, 1]
