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


# ---- the third-party random stream behind Clustering (row f3): rand_xoshiro 0.6.0 + rand 0.8.5, restated in
# csrc/host/clustering.hpp.  Pinned here by the PUBLISHED known-answer vectors of the generators (the same ones the
# crates' own unit tests hold) and by an independent Python statement of the sampler.  The ORDER of the retained models
# against the Rust crates themselves stays unpinned (no cargo in this image).
M64 = (1 << 64) - 1


def test_xoshiro256plusplus_reference_vector(H):
    # xoshiro256plusplus.c by Blackman & Vigna, state {1, 2, 3, 4}; rand_xoshiro/src/xoshiro256plusplus.rs `reference`
    assert H.xoshiro256pp(10, state=[1, 2, 3, 4]) == [
        41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205, 9973669472204895162,
        14011001112246962877, 12406186145184390807, 15849039046786891736, 10450023813501588000]


def test_splitmix64_reference_vector(H):
    # splitmix64.c by Vigna; rand_xoshiro/src/splitmix64.rs `reference` (seed 1477776061723855037)
    assert H.splitmix64(1477776061723855037, 10) == [
        1985237415132408290, 2979275885539914483, 13511426838097143398, 8488337342461049707, 15141737807933549159,
        17093170987380407015, 16389528042912955399, 13177319091862933652, 10841969400225389492, 17094824097954834098]
    # seed_from_u64 fills the four state words from that stream (rand_xoshiro overrides rand_core's default)
    s = H.splitmix64(404, 4)
    assert H.xoshiro256pp(6, seed=404) == H.xoshiro256pp(6, state=s)


class _PyXoshiro:
    """Independent statement of the generator + rand 0.8.5's UniformInt<u32>::sample_single_inclusive + Floyd sampling."""

    def __init__(self, seed):
        x, self.s = seed, []
        for _ in range(4):
            x = (x + 0x9E3779B97F4A7C15) & M64
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
            self.s.append(z ^ (z >> 31))

    def next_u64(self):
        rotl = lambda v, k: ((v << k) | (v >> (64 - k))) & M64
        s = self.s
        r = (rotl((s[0] + s[3]) & M64, 23) + s[0]) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
        return r

    def gen_range_inclusive(self, high):  # rand-0.8.5/src/distributions/uniform.rs, uniform_int_impl sample_single_inclusive
        rng = (high + 1) & 0xffffffff
        if rng == 0:
            return self.next_u64() >> 32
        lz = 32 - rng.bit_length()
        zone = ((rng << lz) - 1) & 0xffffffff
        while True:
            m = (self.next_u64() >> 32) * rng
            if (m & 0xffffffff) <= zone:
                return m >> 32

    def sample(self, length, amount):  # rand-0.8.5/src/seq/index.rs sample_floyd, amount < 50: the shuffled variant
        idx = []
        for j in range(length - amount, length):
            t = self.gen_range_inclusive(j)
            if t in idx:
                idx.insert(idx.index(t), j)
            else:
                idx.append(t)
        return idx


@pytest.mark.parametrize("seed,length,amount", [(404, 1000, 4), (404, 5, 5), (404, 3, 1), (7, 41943, 5), (1, 12, 11), (99, 2, 2)])
def test_choose_multiple_indices(H, seed, length, amount):
    got = H.sample_indices(seed, length, amount)
    assert got == _PyXoshiro(seed).sample(length, amount)
    assert len(set(got)) == amount and all(0 <= i < length for i in got)


def test_gen_range_rejection_zone(H):
    # a range just above a power of two rejects almost half of the draws: exercises the zone test
    for high in (0, 1, 2, 4, 6, 9, 2**31, 2**32 - 1, 3 * 2**30):
        g = _PyXoshiro(404)
        assert H.gen_range(404, high, 64) == [g.gen_range_inclusive(high) for _ in range(64)]


def test_floyd_duplicates_take_the_place_of_the_first_draw(H):
    # hand-computed: amount == length forces j = 0 -> t = 0; j = 1 -> t in {0, 1}; a repeated t inserts j BEFORE it
    for seed in range(40):
        got = H.sample_indices(seed, 4, 4)
        g = _PyXoshiro(seed)
        want = []
        for j in range(4):
            t = g.gen_range_inclusive(j)
            if t in want:
                want.insert(want.index(t), j)
            else:
                want.append(t)
        assert got == want and sorted(got) == [0, 1, 2, 3]


def test_clustering_keeps_one_random_stream(H):
    """The reference's ModelChooser owns ONE Clustering: the q-score clustering draws its initial centroids from where
    the acid clustering left the generator (model_chooser.rs:14-24, compressor_initializer.rs:57-64)."""
    rng = np.random.default_rng(3)
    cost_a = rng.integers(50, 90, size=(500, 6)).astype(np.uint32)
    cost_q = rng.integers(50, 90, size=(500, 7)).astype(np.uint32)
    shared = H.Clustering()
    a1 = shared.make_clusters(cost_a, 4)
    q1 = shared.make_clusters(cost_q, 4)
    fresh_a, fresh_q = H.cluster(cost_a, 4), H.cluster(cost_q, 4)
    assert a1 == fresh_a  # the first use of a Clustering equals a fresh one
    # the second draw comes from the advanced stream: same index sample as the Python statement continued
    g = _PyXoshiro(404)
    first = g.sample(500, 4)
    second = g.sample(500, 4)
    assert first == H.sample_indices(404, 500, 4) and second != first
    # and the converged clusters are valid whatever the start (distinct centroids, every value assigned)
    for cent, vc in (q1, fresh_q):
        assert len(set(cent)) == 4 and set(vc) <= {0, 1, 2, 3}
