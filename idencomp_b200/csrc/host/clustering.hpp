// clustering.hpp -- file-level model subset selection on the host.
//
//   Clustering                clustering.rs:8-118       one Xoshiro256PlusPlus::seed_from_u64(404) per ModelChooser
//                                                       (model_chooser.rs:14-24), used first for the acid models and then,
//                                                       WITHOUT re-seeding, for the q-score models
//                                                       (compressor_initializer.rs:57-64)
//   rank_models               idn/model_chooser.rs:103-138
//
// The random part is third-party code that is not under /root/reference: rand 0.8.5 (`SliceRandom::choose_multiple` ->
// `seq::index::sample` -> Floyd's algorithm for amount < 12, `UniformInt<u32>::sample_single_inclusive`) and
// rand_xoshiro 0.6.0 (Xoshiro256++ with SplitMix64 `seed_from_u64`, `next_u32` = upper half of `next_u64`).  It is
// restated here from the published algorithms; tests/test_host_models.py pins the generator against the published
// known-answer vectors of xoshiro256++ / SplitMix64 and the sampler against hand-computed cases.  PARITY of the
// retained-model ORDER against the Rust crates themselves stays unpinned (no cargo in this image, no golden in the
// reference).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <vector>

namespace idencomp {

struct SplitMix64 {  // rand_xoshiro::SplitMix64 (the seeding stream of seed_from_u64)
    uint64_t x;
    explicit SplitMix64(uint64_t seed) : x(seed) {}
    uint64_t next_u64() {
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

struct Xoshiro256PlusPlus {
    uint64_t s[4];
    explicit Xoshiro256PlusPlus(uint64_t seed) {  // seed_from_u64
        SplitMix64 sm(seed);
        for (auto& w : s) w = sm.next_u64();
    }
    explicit Xoshiro256PlusPlus(const uint64_t state[4]) {  // from_seed (little-endian words)
        for (int i = 0; i < 4; i++) s[i] = state[i];
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // UniformInt<u32>::sample_single_inclusive(0, high): widening multiply with a rejection zone
    uint32_t gen_range_inclusive(uint32_t high) {
        uint32_t range = high + 1;
        if (range == 0) return next_u32();
        uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            uint64_t m = (uint64_t)next_u32() * range;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
};

// rand::seq::index::sample(rng, length, amount) for amount < 12 (the model counts of this path: <= 5): Floyd's
// algorithm, fully shuffled variant (an index already drawn makes room for j at its position)
inline std::vector<uint32_t> sample_floyd(Xoshiro256PlusPlus& rng, uint32_t length, uint32_t amount) {
    std::vector<uint32_t> idx;
    for (uint32_t j = length - amount; j < length; j++) {
        uint32_t t = rng.gen_range_inclusive(j);
        auto pos = std::find(idx.begin(), idx.end(), t);
        if (pos != idx.end()) {
            idx.insert(pos, j);
            continue;
        }
        idx.push_back(t);
    }
    return idx;
}

class Clustering {
public:
    Clustering() : rng_(404) {}  // Clustering::new (clustering.rs:14-18)
    // cost[value * n_centroids + centroid]; returns the centroid index of every cluster, in cluster order
    std::vector<size_t> make_clusters(const std::vector<uint32_t>& cost, size_t n_values, size_t n_centroids, size_t num_clusters,
                                      std::vector<std::vector<size_t>>* members) {
        std::vector<size_t> best;
        if (num_clusters == 0) return best;
        num_clusters = std::min(num_clusters, n_centroids);
        std::vector<bool> avail(n_centroids, true);
        std::vector<size_t> value_cluster(n_values, 0);
        auto best_centroid_for = [&](const std::vector<size_t>& vals) {
            std::vector<uint32_t> sum(n_centroids, 0);
            for (size_t v : vals)
                for (size_t c = 0; c < n_centroids; c++) sum[c] += cost[v * n_centroids + c];
            size_t pick = n_centroids;
            for (size_t c = 0; c < n_centroids; c++)  // stable sort by cost, first available
                if (avail[c] && (pick == n_centroids || sum[c] < sum[pick])) pick = c;
            return pick;
        };
        size_t amount = std::min(num_clusters, n_values);  // choose_multiple yields at most `len` items
        for (uint32_t v : sample_floyd(rng_, (uint32_t)n_values, (uint32_t)amount)) {
            size_t c = best_centroid_for({v});
            best.push_back(c);
            avail[c] = false;
        }
        for (;;) {
            size_t cluster_changes = 0, centroid_changes = 0;
            for (size_t v = 0; v < n_values; v++) {
                size_t pick = 0;
                for (size_t k = 1; k < best.size(); k++)
                    if (cost[v * n_centroids + best[k]] < cost[v * n_centroids + best[pick]]) pick = k;
                if (value_cluster[v] != pick) {
                    value_cluster[v] = pick;
                    cluster_changes++;
                }
            }
            std::fill(avail.begin(), avail.end(), true);
            for (size_t k = 0; k < best.size(); k++) {
                std::vector<size_t> vals;
                for (size_t v = 0; v < n_values; v++)
                    if (value_cluster[v] == k) vals.push_back(v);
                size_t c = best_centroid_for(vals);
                if (best[k] != c) {
                    best[k] = c;
                    centroid_changes++;
                }
                avail[c] = false;
            }
            if (cluster_changes == 0 && centroid_changes == 0) break;
        }
        if (members) {
            members->assign(best.size(), {});
            for (size_t v = 0; v < n_values; v++) (*members)[value_cluster[v]].push_back(v);
        }
        return best;
    }
    Xoshiro256PlusPlus& rng() { return rng_; }

private:
    Xoshiro256PlusPlus rng_;
};

// get_model_ranking (idn/model_chooser.rs:103-138): per read, models sorted by size (stable) get rank 1, 2, ...; the
// models with the smallest rank sums win (stable)
inline std::vector<size_t> rank_models(const std::vector<uint32_t>& cost, size_t n_values, size_t n_models, size_t model_num) {
    std::vector<uint32_t> score(n_models, 0);
    std::vector<size_t> order(n_models);
    for (size_t v = 0; v < n_values; v++) {
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[v * n_models + a] < cost[v * n_models + b]; });
        for (size_t i = 0; i < n_models; i++) score[order[i]] += (uint32_t)i + 1;
    }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return score[a] < score[b]; });
    order.resize(std::min(model_num, n_models));
    return order;
}

}  // namespace idencomp
