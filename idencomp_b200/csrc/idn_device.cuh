// idn_device.cuh -- device-side building blocks of the rANS hot path (sm_100a).
//
// Everything here is integer arithmetic that must reproduce the reference bit for bit:
//   * context-spec generators   idencomp/src/context_spec.rs:218-529, int_queue.rs:40-70
//   * ryg rans_byte put/get     crate rans 0.2.1 -> ryg-rans-sys 1.0.7 (call sites compressor.rs:53-98,173-193)
// The table layouts are ours (see DESIGN.md "Data layout in HBM"): the reference keeps a usize map of
// spec_num entries and a 128 KiB slot->symbol LUT per context (compressor.rs:124-128), which would be GBs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace idn {

constexpr uint32_t kScaleBits = 14;
constexpr uint32_t kAcidSyms = 5, kQSyms = 94;
constexpr uint32_t kSlotMask = (1u << kScaleBits) - 1;
constexpr uint32_t kRansL = 1u << 23;  // RANS_BYTE_L
constexpr int kHist = 8;               // longest acid / q-score history any legal spec type uses
constexpr int kQRowStride = 112;       // u16 per q-score decode row: 16 pivots + 96 cumulative freqs
constexpr int kQBucket = 8;            // symbols per pivot bucket

// divide a value < 2^31 by a constant: q = m ? umulhi(x, m) >> s : x >> s
struct FastDiv {
    uint32_t m, s;
};

__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f{0, 0};
    if (d == 0) d = 1;
    if ((d & (d - 1)) == 0) {  // power of two (incl. 1)
        while ((1u << f.s) < d) f.s++;
        return f;
    }
    uint32_t c = 0;  // ceil(log2 d)
    while ((1ull << c) < d) c++;
    f.s = c - 1;
    f.m = (uint32_t)((((unsigned long long)1 << (32 + f.s)) + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, FastDiv f) {
    return (f.m ? __umulhi(x, f.m) : x) >> f.s;
}

// parameters of one ContextSpecGenerator (generic or light); "dummy" is generic<0,0,0>
struct SpecDev {
    uint32_t ao, qo, pb, light, qmax;
    uint32_t base_a, base_q;  // 5/94 generic, 4/qmax light
    uint32_t abits, qbits;    // IntQueue::num_bits
    uint32_t pow_a, pow_q;    // base^(order-1)  (IntQueue::last_pow), 0 when order == 0
    FastDiv div_base_a, div_base_q;  // backward slide: state / base
    FastDiv div_pow_a, div_pow_q;    // forward push: state % last_pow
};

struct ModelDev {
    SpecDev spec;
    uint32_t type, nsym, n_rows;  // n_rows = n_ctx + 1, row 0 = dummy context
    const uint16_t* map;          // dense spec -> row, or nullptr
    const uint32_t* hkeys;        // open-addressing hash (sparse spec types)
    const uint16_t* hvals;
    uint32_t hmask;
    const uint2* enc;             // [n_rows][nsym] {rcp_freq, start | freq << 14 | rcp_shift << 28}
    const uint16_t* dec;          // acid: [n_rows][4] = cum[1..4]; q: [n_rows][kQRowStride]
};

__device__ __forceinline__ uint32_t hash32(uint32_t k) {
    k ^= k >> 16;
    k *= 0x7feb352dU;
    k ^= k >> 15;
    k *= 0x846ca68bU;
    k ^= k >> 16;
    return k;
}

// RansEncModel/RansDecModel::context_for  (sequence_compressor.rs:60-62, 203-205)
__device__ __forceinline__ uint32_t ctx_row(const ModelDev& m, uint32_t spec) {
    if (m.map) return __ldg(m.map + spec);
    if (!m.hkeys) return 0;  // model without contexts
    uint32_t h = hash32(spec) & m.hmask;
    for (;;) {
        uint32_t k = __ldg(m.hkeys + h);
        if (k == spec) return __ldg(m.hvals + h);
        if (k == 0xffffffffu) return 0;
        h = (h + 1) & m.hmask;
    }
}

// LightContextSpecGenerator::update mapping  (context_spec.rs:516-529)
__device__ __forceinline__ void light_map(uint32_t a, uint32_t q, uint32_t qmax, uint32_t& va, uint32_t& vq) {
    if (a == 0 || q == 0) {
        va = 0;
        vq = 0;
    } else {
        va = a - 1;
        vq = (q * qmax * 11156u) >> 20;  // q*qmax/94, exact for q*qmax <= 8742
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward generator: used by the scorer and the decoder (symbols become known one by one).
//   current_context  context_spec.rs:377-383 / 508-514      update  :385-389 / :516-529
// ---------------------------------------------------------------------------------------------------
struct GenFwd {
    uint32_t sa, sq;   // queue states, initial 0 (int_queue.rs:24-30)
    uint32_t pos, rem; // pos = floor(i * 2^pb / len), rem = i * 2^pb - pos * len

    __device__ __forceinline__ void init() { sa = sq = pos = rem = 0; }

    __device__ __forceinline__ uint32_t spec(const SpecDev& s) const {
        return (((sq << s.abits) | sa) << s.pb) | pos;
    }

    __device__ __forceinline__ void update(const SpecDev& s, uint32_t a, uint32_t q, uint32_t len) {
        uint32_t va = a, vq = q;
        if (s.light) light_map(a, q, s.qmax, va, vq);
        if (s.ao) sa = (sa - fastdiv(sa, s.div_pow_a) * s.pow_a) * s.base_a + va;
        if (s.qo) sq = (sq - fastdiv(sq, s.div_pow_q) * s.pow_q) * s.base_q + vq;
        if (s.pb) {
            rem += 1u << s.pb;
            while (rem >= len) {
                rem -= len;
                pos++;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Symbol window for the backward (encode) pass: entries e_k = symbol at position i-k, k = 0..8.
// acids: 3 bits per entry in a u32; quality scores: 7 bits per entry in a u64.
// ---------------------------------------------------------------------------------------------------
struct SymWindow {
    uint32_t wa;
    unsigned long long wq;
    __device__ __forceinline__ void init() {
        wa = 0;
        wq = 0;
    }
    __device__ __forceinline__ void shift_in(uint32_t a, uint32_t q) {
        wa = (wa >> 3) | (a << (3 * kHist));
        wq = (wq >> 7) | ((unsigned long long)q << (7 * kHist));
    }
    __device__ __forceinline__ uint32_t acid(uint32_t k) const { return (wa >> (3 * k)) & 7u; }
    __device__ __forceinline__ uint32_t qual(uint32_t k) const { return (uint32_t)(wq >> (7 * k)) & 127u; }
};

// Backward generator: state at position i is rebuilt from the symbols i-order .. i-1 held in the window.
//   state_i = floor(state_{i+1} / base) + v[i - order] * base^(order-1)
struct GenBack {
    uint32_t sa, sq;
    uint32_t pos, rem;

    // state at position `len` (all symbols consumed); the window holds e_k = symbol at len-k
    __device__ __forceinline__ void init(const SpecDev& s, const SymWindow& w, uint32_t len) {
        sa = sq = 0;
        for (uint32_t k = s.ao; k >= 1; k--) {
            uint32_t va = w.acid(k), vq = w.qual(k);
            if (s.light) light_map(va, vq, s.qmax, va, vq);
            sa = sa * s.base_a + va;
        }
        for (uint32_t k = s.qo; k >= 1; k--) {
            uint32_t va = w.acid(k), vq = w.qual(k);
            if (s.light) light_map(va, vq, s.qmax, va, vq);
            sq = sq * s.base_q + vq;
        }
        // position(len) = len * 2^pb / len
        pos = 1u << s.pb;
        rem = 0;
        if (s.pb == 0) pos = 0;
    }

    // move from position i+1 to i; the window has already been shifted so that e_0 = symbol i
    __device__ __forceinline__ void step_back(const SpecDev& s, const SymWindow& w, uint32_t len) {
        if (s.ao) {
            uint32_t va = w.acid(s.ao), vq = w.qual(s.ao);
            if (s.light) light_map(va, vq, s.qmax, va, vq);
            sa = fastdiv(sa, s.div_base_a) + va * s.pow_a;
        }
        if (s.qo) {
            uint32_t va = w.acid(s.qo), vq = w.qual(s.qo);
            if (s.light) light_map(va, vq, s.qmax, va, vq);
            sq = fastdiv(sq, s.div_base_q) + vq * s.pow_q;
        }
        if (s.pb) {  // i*P = pos*len + rem  ->  (i-1)*P
            uint32_t P = 1u << s.pb;
            while (rem < P) {
                rem += len;
                pos--;
            }
            rem -= P;
        }
    }

    __device__ __forceinline__ uint32_t spec(const SpecDev& s) const {
        return (((sq << s.abits) | sa) << s.pb) | pos;
    }
};

// ---------------------------------------------------------------------------------------------------
// rANS encode step (RansEncPutSymbol).  `emit(byte)` receives renormalisation bytes in emission order
// (the reference writes them at decreasing addresses).
// ---------------------------------------------------------------------------------------------------
template <class Emit>
__device__ __forceinline__ void rans_put(uint32_t& x, uint2 e, Emit&& emit) {
    uint32_t freq = (e.y >> 14) & 0x3fffu;
    uint32_t start = e.y & 0x3fffu;
    uint32_t x_max = freq << (23 - kScaleBits + 8);  // ((L >> scale_bits) << 8) * freq
    while (x >= x_max) {
        emit(x & 0xffu);
        x >>= 8;
    }
    // exact floor(x / freq); ryg's reciprocal (rcp_freq, rcp_shift) is exact for x < 2^31, freq >= 2
    uint32_t q = freq == 1 ? x : (__umulhi(x, e.x) >> (e.y >> 28));
    x = x + start + q * ((1u << kScaleBits) - freq);
}

// size-only variant for the scorer (ModelTester::compute_size): counts emitted bytes
__device__ __forceinline__ void rans_put_count(uint32_t& x, uint2 e, uint32_t& bytes) {
    uint32_t freq = (e.y >> 14) & 0x3fffu;
    uint32_t start = e.y & 0x3fffu;
    uint32_t x_max = freq << (23 - kScaleBits + 8);
    while (x >= x_max) {
        bytes++;
        x >>= 8;
    }
    uint32_t q = freq == 1 ? x : (__umulhi(x, e.x) >> (e.y >> 28));
    x = x + start + q * ((1u << kScaleBits) - freq);
}

// count of 16-bit lanes a (<= 0x7fff each) with a <= s, over a 32-bit word holding two of them
__device__ __forceinline__ uint32_t le_mask2(uint32_t w, uint32_t ss) {
    return ((ss | 0x80008000u) - w) & 0x80008000u;
}

// acid symbol search: packed = cum[1..4] as 4 x u16.  returns symbol, sets start/freq
__device__ __forceinline__ uint32_t acid_find(uint2 packed, uint32_t slot, uint32_t& start, uint32_t& freq) {
    uint32_t ss = slot | (slot << 16);
    uint32_t sym = __popc(le_mask2(packed.x, ss)) + __popc(le_mask2(packed.y, ss));  // #(cum[1..4] <= slot)
    unsigned long long all = ((unsigned long long)packed.y << 32) | packed.x;        // field k = cum[k+1]
    start = sym == 0 ? 0u : (uint32_t)(all >> (16 * (sym - 1))) & 0xffffu;
    uint32_t next = sym == 4 ? (1u << kScaleBits) : (uint32_t)(all >> (16 * sym)) & 0xffffu;
    freq = next - start;
    return sym;
}

__device__ __forceinline__ uint32_t sel4(uint4 v, uint32_t i) {
    return (i & 2) ? ((i & 1) ? v.w : v.z) : ((i & 1) ? v.y : v.x);
}
__device__ __forceinline__ uint32_t half_of(uint32_t w, uint32_t odd) { return odd ? (w >> 16) : (w & 0xffffu); }

// q-score symbol search in a decode row: 16 pivots (cum[8k], padded 0x7fff) then 96 cums (padded 0x7fff
// after cum[94] = 16384).  Two dependent 16/32-byte vector loads, everything else in registers.
__device__ __forceinline__ uint32_t q_find(const uint16_t* __restrict__ row, uint32_t slot, uint32_t& start,
                                           uint32_t& freq) {
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
    uint4 p0 = __ldg(r4), p1 = __ldg(r4 + 1);
    uint32_t ss = slot | (slot << 16);
    uint32_t m = le_mask2(p0.x, ss) | (le_mask2(p0.y, ss) >> 1) | (le_mask2(p0.z, ss) >> 2) |
                 (le_mask2(p0.w, ss) >> 3) | (le_mask2(p1.x, ss) >> 4) | (le_mask2(p1.y, ss) >> 5) |
                 (le_mask2(p1.z, ss) >> 6) | (le_mask2(p1.w, ss) >> 7);
    uint32_t b = __popc(m) - 1;   // pivot[0] = 0 <= slot always
    uint4 c = __ldg(r4 + 2 + b);  // cum[8b .. 8b+7]
    uint32_t mm = le_mask2(c.x, ss) | (le_mask2(c.y, ss) >> 1) | (le_mask2(c.z, ss) >> 2) | (le_mask2(c.w, ss) >> 3);
    uint32_t j = __popc(mm) - 1;  // 0..7
    start = half_of(sel4(c, j >> 1), j & 1);
    uint32_t j1 = j + 1, next;
    if (j1 < 8) {
        next = half_of(sel4(c, j1 >> 1), j1 & 1);
    } else {  // cum[8(b+1)] = pivot[b+1]; b <= 10 here (bucket 11 ends at j = 5)
        uint32_t b1 = b + 1;
        next = half_of((b1 & 8) ? sel4(p1, (b1 >> 1) & 3) : sel4(p0, (b1 >> 1) & 3), b1 & 1);
    }
    freq = next - start;
    return b * kQBucket + j;
}

}  // namespace idn
