// idn_device.cuh -- device-side building blocks of the rANS hot path (sm_100a).
//
// Everything here is integer arithmetic that must reproduce the reference bit for bit:
//   * context-spec generators   idencomp/src/context_spec.rs:218-529, int_queue.rs:40-70
//   * ryg rans_byte put/get     crate rans 0.2.1 -> ryg-rans-sys 1.0.7 (call sites compressor.rs:53-98,173-193)
// The table layouts are ours (see DESIGN.md "Data layout in HBM"): the reference keeps a usize map of
// spec_num entries and a 128 KiB slot->symbol LUT per context (compressor.rs:124-128), which would be GBs.
//
// The codec kernels are instruction-issue bound (profiles/r1_ncu_v1.md), so the generators are written
// branch-free: one formula covers every legal spec type (generic / light, any order, any position bits) with
// per-model constants instead of per-order branches.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace idn {

constexpr uint32_t kScaleBits = 14;
constexpr uint32_t kAcidSyms = 5, kQSyms = 94;
constexpr uint32_t kSlotMask = (1u << kScaleBits) - 1;
constexpr uint32_t kRansL = 1u << 23;  // RANS_BYTE_L
constexpr int kHist = 8;               // longest acid / q-score history any legal spec type uses
// q-score decode row (512 bytes): 128-byte bucket LUT (slot >> 7 -> (symbol that owns the bucket's first slot) >> 2) followed
// by 24 overlapping 16-byte windows, window g = the cumulative frequencies ("starts", u16) of the symbols 4g .. 4g+7
// (everything from symbol 94 on = 2^14).  The search reads one LUT byte and then ONE 16-byte window (a single gather
// request: the kernels are bound by the number of divergent memory requests, profiles/r2_ncu_*.md) and resolves the symbol
// in registers.
constexpr int kQLutBytes = 128;
#ifndef IDN_QROW352
constexpr int kQWindows = 24;                            // g = 0 .. 23 (s0 <= 93; the window of g = 23 ends at 2^14: no advance)
constexpr int kQRowBytes = kQLutBytes + 16 * kQWindows;  // 512
#else  // round-1 layout, kept for A/B measurements: LUT + the 104 starts once, the window is two 8-byte loads
constexpr int kQStarts = 104;
constexpr int kQRowBytes = 352;
#endif
// q-score "window" rows (decoder): one 16-byte entry per 128-slot bucket = {first symbol s0 of the bucket, starts of the
// symbols s0 .. s0+6}, i.e. bucket LUT and starts window in ONE gather (2 KB per row instead of 352 B, one dependent
// load level less).  The compact row stays for slots the window cannot resolve and for the workload generator.
constexpr int kQWinBytes = 128 * 16;

// floor(x / d) for x < 2^31 as umulhi(x, m) >> s.  d = 2^k (k >= 1): m = 2^(32-k), s = 0; otherwise the round-up
// reciprocal with s = ceil(log2 d) - 1 (the same argument as ryg's RansEncSymbolInit).  d <= 1 is never divided by.
struct Magic {
    uint32_t m, s;
};
__host__ __device__ constexpr Magic make_magic(uint32_t d) {
    Magic f{0, 0};
    if (d <= 1) return f;
    if ((d & (d - 1)) == 0) {
        uint32_t k = 0;
        while ((1u << k) < d) k++;
        f.m = 1u << (32 - k);
        return f;
    }
    uint32_t c = 0;  // ceil(log2 d)
    while ((1ull << c) < d) c++;
    f.s = c - 1;
    f.m = (uint32_t)((((unsigned long long)1 << (32 + f.s)) + d - 1) / d);
    return f;
}

// One IntQueue<B, n> (int_queue.rs:40-70), newest digit lowest.
//   forward push (with_pushed_back):  s' = (s mod B^(n-1)) * B + v
//        = s*B - floor(s / M)*M*B + v*vmul   (mod 2^32),  M = B^(n-1)
//        n == 0: B = MB = vmul = 0 -> s stays 0;   n == 1: B = MB = 0 -> s' = v
//   backward slide (undo the newest push, re-append the digit that had dropped out):
//        s = floor(s' / B) + v_old * M        n == 0: everything 0
struct QueueDev {
    uint32_t B, m, sh, MB, vmul;  // forward
    uint32_t mb, shb, powmul;     // backward
    uint32_t depth;               // n
};

// parameters of one ContextSpecGenerator (generic or light); "dummy" is generic<0,0,0>
struct SpecDev {
    QueueDev qa, qq;  // acid queue, quality score queue
    uint32_t abits;   // IntQueue::num_bits of the acid queue
    uint32_t pb;      // position bits
    uint32_t sbits;   // bits of the two queue states together (spec bits without the position)
    // symbol mapping (LightContextSpecGenerator::update, context_spec.rs:516-529): generic asub = 0, qmul = 2^20
    uint32_t asub, qmul, light;
};

// ---- generator parameters from (kind, acid order, q order, position bits, q max)   context_spec.rs:218-529 ----
// constexpr so that the same code builds the run-time tables (idn_gpu_model_upload) and the compile-time constants of
// the kernels specialised for the spec types of the bundled models (StaticSpecs below).
struct SpecBuild {
    SpecDev spec;
    uint32_t total_bits;
    bool ok;
};

__host__ __device__ constexpr uint64_t spec_ipow(uint64_t b, uint32_t n) {
    uint64_t r = 1;
    while (n--) r *= b;
    return r;
}
__host__ __device__ constexpr uint32_t spec_bitlen(uint64_t v) {  // IntQueue::num_bits (int_queue.rs:40-43) of v = B^n - 1
    uint32_t n = 0;
    while (v) {
        n++;
        v >>= 1;
    }
    return n;
}

// IntQueue<B, n> constants for the branch-free push / slide; returns false when B^n does not fit
__host__ __device__ constexpr bool make_queue(uint32_t base, uint32_t order, QueueDev& q, uint32_t& bits) {
    q = QueueDev{0, 0, 0, 0, 0, 0, 0, 0, order};
    uint64_t pw = spec_ipow(base, order);  // B^n
    if (pw > (1ull << 31)) return false;
    bits = spec_bitlen(pw - 1);
    if (order == 0) return true;  // everything 0: the state stays 0
    q.vmul = 1;
    uint32_t M = (uint32_t)spec_ipow(base, order - 1);
    q.powmul = M;
    Magic mb = make_magic(base);  // base >= 2 whenever order >= 1 (light qmax = 1 is handled by the caller)
    q.mb = mb.m;
    q.shb = mb.s;
    if (order >= 2) {
        q.B = base;
        Magic mm = make_magic(M);
        q.m = mm.m;
        q.sh = mm.s;
        q.MB = (uint32_t)((uint64_t)M * base);
    }
    return true;
}

__host__ __device__ constexpr SpecBuild make_spec(int kind, int ao, int qo, int pb, int qmax) {
    SpecBuild r{SpecDev{QueueDev{0, 0, 0, 0, 0, 0, 0, 0, 0}, QueueDev{0, 0, 0, 0, 0, 0, 0, 0, 0}, 0, 0, 0, 0, 0, 0}, 0, false};
    if (kind != 0 && kind != 1) return r;  // IDN_SPEC_GENERIC / IDN_SPEC_LIGHT
    if (ao < 0 || ao > kHist || qo < 0 || qo > kHist || pb < 0 || pb > 16) return r;
    SpecDev& s = r.spec;
    s.pb = (uint32_t)pb;
    s.light = kind == 1;
    uint32_t base_a = 5, base_q = 94;
    s.asub = 0;
    s.qmul = 1u << 20;
    if (s.light) {
        if (qmax < 1 || qmax > 94) return r;
        base_a = 4;
        base_q = (uint32_t)qmax;
        s.asub = 1;
        s.qmul = (uint32_t)qmax * 11156u;
        for (uint32_t q = 0; q < 94; q++)  // the multiply-shift in map_syms must equal q*qmax/94
            if (((q * s.qmul) >> 20) != q * (uint32_t)qmax / 94) return r;
    }
    uint32_t qbits = 0;
    if (!make_queue(base_a, (uint32_t)ao, s.qa, s.abits)) return r;
    if (base_q == 1) {
        // IntQueue<1, n>: every digit is 0 and the state stays 0 (light qmax = 1, e.g. light_ao8_qo0_pb0_qm1)
        s.qq = QueueDev{0, 0, 0, 0, 0, 0, 0, 0, (uint32_t)qo};
        qbits = 0;
    } else if (!make_queue(base_q, (uint32_t)qo, s.qq, qbits)) {
        return r;
    }
    if (s.abits + qbits + s.pb > 31) return r;
    s.sbits = s.abits + qbits;
    r.total_bits = s.abits + qbits + s.pb;
    r.ok = true;
    return r;
}

// Spec policies of the codec kernels: DynSpecs reads the generator parameters from the model (any legal spec type);
// StaticSpecs<...> makes them compile-time constants for one (acid spec, q spec) pair, which turns the generic
// multiply/reciprocal queue arithmetic into shifts and masks and deletes what the pair does not use.
struct DynSpecs {
    static constexpr bool kStatic = false;
    static constexpr bool kQWin = false;
    __host__ __device__ static constexpr SpecDev sa() { return make_spec(0, 0, 0, 0, 0).spec; }
    __host__ __device__ static constexpr SpecDev sq() { return make_spec(0, 0, 0, 0, 0).spec; }
};
template <int KA, int AOA, int QOA, int PBA, int QMA, int KQ, int AOQ, int QOQ, int PBQ, int QMQ>
struct StaticSpecs {
    static constexpr bool kStatic = true;
    // the decoder searches q-score symbols through the 2 KB "window" rows (one gather level less, 6 x the row footprint).
    // Measured per bundled pair on the 10 GB workloads: NovaSeq (generic_ao2_qo1_pb6, 2 154 evenly used rows) decode
    // 46.9 -> 44.1 ms; HiSeq (generic_ao0_qo2_pb6, a few hot rows) 34.4 -> 35.7 ms; Sequel II 25.6 -> 26.6 ms.
#ifdef IDN_QWIN_ALL
    static constexpr bool kQWin = true;
#elif defined(IDN_QWIN_NONE)
    static constexpr bool kQWin = false;
#else
    static constexpr bool kQWin = KQ == 0 && AOQ == 2 && QOQ == 1 && PBQ == 6;
#endif
    __host__ __device__ static constexpr SpecDev sa() { return make_spec(KA, AOA, QOA, PBA, QMA).spec; }
    __host__ __device__ static constexpr SpecDev sq() { return make_spec(KQ, AOQ, QOQ, PBQ, QMQ).spec; }
};

struct ModelDev {
    SpecDev spec;
    uint32_t type, nsym, n_rows;  // n_rows = n_ctx + 1, row 0 = dummy context
    const uint16_t* map;          // dense spec -> row, or nullptr
    const uint32_t* hkeys;        // open-addressing hash (sparse spec types)
    const uint32_t* hvals;        // row per key (u32: a model may hold 65 536 contexts + the dummy row, sequence_compressor.rs:209-219)
    uint32_t hmask;
    const uint2* enc;             // [n_rows][nsym] {rcp_freq, start | freq << 14 | rcp_shift << 28}
    const uint8_t* dec;           // acid: [n_rows] x 8 bytes = cum[1..4] u16; q: [n_rows] x kQRowBytes
    const uint4* qwin;            // q-scores: [n_rows][128] window entries (kQWinBytes per row), or nullptr
    const uint2* adirect;         // acid, small dense spec spaces: the decode row of every spec (spec -> cum[1..4]), or nullptr
    const uint2* aenc;            // acid, small dense spec spaces: the encoder entries of every spec, [spec][5] (spec -> row -> entry
                                  // becomes ONE gather in the encoder and the scorer), or nullptr
};

// 8-byte read-only gather that does not allocate in L1: the per-spec acid tables (hundreds of KB, one random line per lane
// per position) otherwise sweep L1 clean once per position and take the small position-local q-score tables with them
__device__ __forceinline__ uint2 ldg_stream8(const uint2* p) {
#ifdef IDN_ACID_NOALLOC
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

__device__ __forceinline__ uint32_t hash32(uint32_t k) {
    k ^= k >> 16;
    k *= 0x7feb352dU;
    k ^= k >> 15;
    k *= 0x846ca68bU;
    k ^= k >> 16;
    return k;
}

// Index of a spec in the dense per-spec tables (map, adirect, aenc): the spec with its position field moved to the TOP,
//     index = position << sbits | queue states        (spec = queue states << pb | position, context_spec.rs:377-383)
// The reads of a thread block have similar lengths and advance in lock step, so at any moment its threads look up specs
// of (nearly) one position: position-major tables turn those gathers into hits in one small slice (2^sbits entries, e.g.
// 17 KB of the 2 MB map of generic_ao0_qo2_pb6) that stays in L1, and the slices are walked front to back.
__host__ __device__ constexpr uint32_t spec_table_index(const SpecDev& s, uint32_t spec) {
    return s.pb == 0 ? spec : (((spec & ((1u << s.pb) - 1u)) << s.sbits) | (spec >> s.pb));
}

// RansEncModel/RansDecModel::context_for  (sequence_compressor.rs:60-62, 203-205) through the hash of a sparse spec space
__device__ __forceinline__ uint32_t ctx_row_hash(const ModelDev& m, uint32_t spec) {
    if (!m.hkeys) return 0;  // model without contexts
    uint32_t h = hash32(spec) & m.hmask;
    for (;;) {
        uint32_t k = __ldg(m.hkeys + h);
        if (k == spec) return __ldg(m.hvals + h);
        if (k == 0xffffffffu) return 0;
        h = (h + 1) & m.hmask;
    }
}

// context row of the spec a generator stands at.  kDense: the caller guarantees the dense table (the specialised kernels
// are only launched for such models)
template <bool kDense = false, class Gen>
__device__ __forceinline__ uint32_t gen_row(const ModelDev& m, const SpecDev& s, const Gen& g, uint32_t pos_shared, uint32_t pshift) {
    if (kDense || m.map) return __ldg(m.map + g.index(s, pos_shared, pshift));
    return ctx_row_hash(m, g.spec(s, pos_shared, pshift));
}

// (acid, q) -> queue digits.  z = (acid == N || q == 0), computed once per symbol for both generators.
__device__ __forceinline__ void map_syms(const SpecDev& s, uint32_t a, uint32_t q, bool z, uint32_t& va, uint32_t& vq) {
    va = a - s.asub;
    vq = (q * s.qmul) >> 20;  // light: q * qmax / 94 (exact for q * qmax <= 8742); generic: q
    const bool kill = z && s.light != 0;
    va = kill ? 0u : va;
    vq = kill ? 0u : vq;
}

__device__ __forceinline__ uint32_t queue_push(const QueueDev& k, uint32_t s, uint32_t v) {
    uint32_t q = __umulhi(s, k.m) >> k.sh;
    return s * k.B - q * k.MB + v * k.vmul;
}
__device__ __forceinline__ uint32_t queue_slide_back(const QueueDev& k, uint32_t s, uint32_t v_old) {
    return (__umulhi(s, k.mb) >> k.shb) + v_old * k.powmul;
}

// position(i) = floor(i * 2^pb / len)   (context_spec.rs:313-316), kept for pbmax = the larger pb of the two
// generators of a read; floor(floor(x * 2^k) / 2^k) = floor(x), so each generator shifts it down to its own pb.
struct PosFwd {
    uint32_t pos, rem, step, frac, len;
    __device__ __forceinline__ void init(uint32_t length, uint32_t pbmax) {
        len = length ? length : 1;
        uint32_t P = 1u << pbmax;
        step = P / len;
        frac = P - step * len;
        pos = rem = 0;
    }
    // the state after i advances: pos * len + rem = i * 2^pbmax
    __device__ __forceinline__ void init_at(uint32_t length, uint32_t pbmax, uint32_t i) {
        init(length, pbmax);
        const unsigned long long t = (unsigned long long)i << pbmax;
        pos = (uint32_t)(t / len);
        rem = (uint32_t)(t % len);
    }
    __device__ __forceinline__ void advance() {
        pos += step;
        rem += frac;
        if (rem >= len) {
            rem -= len;
            pos++;
        }
    }
};
struct PosBack {  // starts at i = len and steps down
    uint32_t pos, rem, step, frac, len;
    __device__ __forceinline__ void init(uint32_t length, uint32_t pbmax) {
        len = length ? length : 1;
        uint32_t P = 1u << pbmax;
        step = P / len;
        frac = P - step * len;
        pos = P;  // len * P / len
        rem = 0;
    }
    // standing at position i (<= length) instead of at the end
    __device__ __forceinline__ void init_at(uint32_t length, uint32_t pbmax, uint32_t i) {
        init(length, pbmax);
        const unsigned long long t = (unsigned long long)i << pbmax;
        pos = (uint32_t)(t / len);
        rem = (uint32_t)(t % len);
    }
    __device__ __forceinline__ void retreat() {
        pos -= step;
        if (rem < frac) {
            rem += len;
            pos--;
        }
        rem -= frac;
    }
};

// ---------------------------------------------------------------------------------------------------
// Forward generator: scorer, decoder, workload generator (symbols become known one by one).
//   current_context  context_spec.rs:377-383 / 508-514      update  :385-389 / :516-529
// ---------------------------------------------------------------------------------------------------
struct GenFwd {
    uint32_t sa, sq;  // queue states, initial 0 (int_queue.rs:24-30)
    __device__ __forceinline__ void init() { sa = sq = 0; }
    // pos_shared = PosFwd::pos at pbmax, pshift = pbmax - s.pb
    __device__ __forceinline__ uint32_t spec(const SpecDev& s, uint32_t pos_shared, uint32_t pshift) const {
        if (s.pb == 0) return (sq << s.abits) | sa;  // position(i) < 2^pbmax for i < len: nothing to add
        return (((sq << s.abits) | sa) << s.pb) | (pos_shared >> pshift);
    }
    // spec_table_index(spec()) without the detour
    __device__ __forceinline__ uint32_t index(const SpecDev& s, uint32_t pos_shared, uint32_t pshift) const {
        if (s.pb == 0) return (sq << s.abits) | sa;
        return ((pos_shared >> pshift) << s.sbits) | ((sq << s.abits) | sa);
    }
    __device__ __forceinline__ void update(const SpecDev& s, uint32_t a, uint32_t q, bool z) {
        uint32_t va, vq;
        map_syms(s, a, q, z, va, vq);
        sa = queue_push(s.qa, sa, va);
        sq = queue_push(s.qq, sq, vq);
    }
};

// ---------------------------------------------------------------------------------------------------
// Backward generator (encoder: symbols are walked last -> first).  Keeps the digits of the positions
// i-k, k = 0..kHist, in two shift registers (3 bits per acid digit, 7 bits per q digit) fed kHist positions
// ahead of the encode position; the state at i is the state at i+1 with the newest digit removed and the digit
// of position i - depth re-appended.
// ---------------------------------------------------------------------------------------------------
struct GenBack {
    uint32_t sa, sq;
    uint32_t wa;             // entry k at bits 3k
    unsigned long long wq;   // entry k at bits 7k
    uint32_t sha, shq;       // 3 * depth_a, 7 * depth_q

    __device__ __forceinline__ void clear(const SpecDev& s) {
        sa = sq = 0;
        wa = 0;
        wq = 0;
        sha = 3 * s.qa.depth;
        shq = 7 * s.qq.depth;
    }
    // pull the symbol of the position kHist places before the one that becomes entry 0 (zeros before the read starts)
    __device__ __forceinline__ void shift_in(const SpecDev& s, uint32_t a, uint32_t q, bool z) {
        uint32_t va, vq;
        map_syms(s, a, q, z, va, vq);
        wa = (wa >> 3) | (va << (3 * kHist));
        wq = (wq >> 7) | ((unsigned long long)vq << (7 * kHist));
    }
    // state at position `len` once the window holds entry k = digit of position len - k:  push digits oldest first
    __device__ __forceinline__ void init_state(const SpecDev& s) {
        sa = sq = 0;
        for (uint32_t k = s.qa.depth; k >= 1; k--) sa = queue_push(s.qa, sa, (wa >> (3 * k)) & 7u);
        for (uint32_t k = s.qq.depth; k >= 1; k--) sq = queue_push(s.qq, sq, (uint32_t)(wq >> (7 * k)) & 127u);
    }
    // from position i+1 to i; the window has already been shifted so that entry 0 is position i
    __device__ __forceinline__ void step_back(const SpecDev& s) {
        sa = queue_slide_back(s.qa, sa, (wa >> sha) & 7u);
        sq = queue_slide_back(s.qq, sq, (uint32_t)(wq >> shq) & 127u);
    }
    __device__ __forceinline__ uint32_t spec(const SpecDev& s, uint32_t pos_shared, uint32_t pshift) const {
        if (s.pb == 0) return (sq << s.abits) | sa;
        return (((sq << s.abits) | sa) << s.pb) | (pos_shared >> pshift);
    }
    __device__ __forceinline__ uint32_t index(const SpecDev& s, uint32_t pos_shared, uint32_t pshift) const {
        if (s.pb == 0) return (sq << s.abits) | sa;
        return ((pos_shared >> pshift) << s.sbits) | ((sq << s.abits) | sa);
    }
};

// ---------------------------------------------------------------------------------------------------
// rANS encode step (RansEncPutSymbol: while (x >= x_max) { *--ptr = x & 0xff; x >>= 8; } then the reciprocal
// division), without branches: a state emits 0, 1 or 2 bytes per symbol (x < 2^31 and x_max >= 2^17).  `out` is a
// BackWriter-like sink with push_bits(bytes in emission order, 8 * count).
template <class Out>
__device__ __forceinline__ void rans_put_bf(uint32_t& x, uint2 e, Out& out) {
    const uint32_t freq = (e.y >> 14) & 0x3fffu;
    const uint32_t start = e.y & 0x3fffu;
    const uint32_t x_max = freq << (23 - kScaleBits + 8);
    const bool one = x >= x_max, two = (x >> 8) >= x_max;  // two implies one
    const uint32_t sel = two ? 0x4401u : (one ? 0x4440u : 0x4444u);  // [x1 x0] / [x0] / nothing, first emitted byte highest
    const uint32_t kbits = two ? 16u : (one ? 8u : 0u);
    out.push_bits(__byte_perm(x, 0, sel), kbits);
    x >>= kbits;
    const uint32_t q = freq == 1 ? x : (__umulhi(x, e.x) >> (e.y >> 28));
    x = x + start + q * ((1u << kScaleBits) - freq);
}

// size-only variant for the scorer (ModelTester::compute_size): counts emitted bytes
__device__ __forceinline__ void rans_put_count(uint32_t& x, uint2 e, uint32_t& bytes) {
    uint32_t freq = (e.y >> 14) & 0x3fffu;
    uint32_t start = e.y & 0x3fffu;
    uint32_t x_max = freq << (23 - kScaleBits + 8);
    const uint32_t k = (x >= x_max ? 1u : 0u) + ((x >> 8) >= x_max ? 1u : 0u);  // 0, 1 or 2 bytes (see rans_put_bf)
    bytes += k;
    x >>= 8 * k;
    uint32_t q = freq == 1 ? x : (__umulhi(x, e.x) >> (e.y >> 28));
    x = x + start + q * ((1u << kScaleBits) - freq);
}

// acid symbol search: packed = cum[1..4] as 4 x u16.  returns symbol, sets start/freq
__device__ __forceinline__ uint32_t acid_find(uint2 packed, uint32_t slot, uint32_t& start, uint32_t& freq) {
    const uint32_t c1 = packed.x & 0xffffu, c2 = packed.x >> 16, c3 = packed.y & 0xffffu, c4 = packed.y >> 16;
    const bool p1 = slot >= c1, p2 = slot >= c2, p3 = slot >= c3, p4 = slot >= c4;
    start = p4 ? c4 : (p3 ? c3 : (p2 ? c2 : (p1 ? c1 : 0u)));
    const uint32_t next = !p1 ? c1 : (!p2 ? c2 : (!p3 ? c3 : (!p4 ? c4 : (1u << kScaleBits))));
    freq = next - start;
    return p4 ? 4u : (p3 ? 3u : (p2 ? 2u : (p1 ? 1u : 0u)));
}

// q-score symbol search in a decode row (layout above): bucket LUT -> window of 8 starts -> compare in registers.
// starts[4g] <= slot holds by construction; the window resolves the symbols 4g .. 4g+6 (the freq of a symbol needs the
// next start).  A slot beyond them (>= 4 symbols inside what is left of a 128-slot bucket: only in the flat tail of
// a distribution) moves on to the next window (4 symbols further).
#ifndef IDN_QROW352
__device__ __forceinline__ uint32_t q_find(const uint8_t* __restrict__ row, uint32_t slot, uint32_t& start,
                                           uint32_t& freq) {
    uint32_t g = __ldg(row + (slot >> 7));
    const uint4* wins = reinterpret_cast<const uint4*>(row + kQLutBytes);
    // halves of K - W: bit 15 set <=> slot >= start (K = 0x8000 | slot in both halves, starts <= 2^14: no borrow)
    const uint32_t K = (slot | 0x8000u) * 0x10001u;
    uint4 w = __ldg(wins + g);
    uint32_t flags;
    for (;;) {
        flags = ((K - w.x) & 0x80008000u) | (((K - w.y) & 0x80008000u) >> 1) | (((K - w.z) & 0x80008000u) >> 2) |
                (((K - w.w) & 0x80008000u) >> 3);
#ifndef IDN_ABL_NOADV
        if (!(flags & 0x10000000u)) break;  // bit 28 = the last start of the window: still <= slot -> move on
        g++;
        w = __ldg(wins + g);
#else
        break;
#endif
    }
    const uint32_t j = __popc(flags) - 1;  // index of the last start <= slot (the flags are monotone)
    const uint32_t p = j >> 1;
    const uint32_t w0 = p == 0 ? w.x : (p == 1 ? w.y : (p == 2 ? w.z : w.w));
    const uint32_t w1 = p == 0 ? w.y : (p == 1 ? w.z : w.w);
    const uint32_t pair = __funnelshift_r(w0, w1, 16 * (j & 1));  // starts[j] | starts[j+1] << 16
    start = pair & 0xffffu;
    freq = (pair >> 16) - start;
    return 4 * g + j;
}
#else
__device__ __forceinline__ uint32_t q_find(const uint8_t* __restrict__ row, uint32_t slot, uint32_t& start,
                                           uint32_t& freq) {
    uint32_t g = __ldg(row + (slot >> 7));
    const uint2* starts = reinterpret_cast<const uint2*>(row + kQLutBytes);
    const uint32_t K = (slot | 0x8000u) * 0x10001u;
    uint2 lo = __ldg(starts + g), hi = __ldg(starts + g + 1);
    uint32_t flags;
    for (;;) {
        flags = ((K - lo.x) & 0x80008000u) | (((K - lo.y) & 0x80008000u) >> 1) | (((K - hi.x) & 0x80008000u) >> 2) |
                (((K - hi.y) & 0x80008000u) >> 3);
        if (!(flags & 0x10000000u)) break;
        g++;
        lo = hi;
        hi = __ldg(starts + g + 1);
    }
    const uint32_t j = __popc(flags) - 1;
    const uint32_t p = j >> 1;
    const uint32_t w0 = p == 0 ? lo.x : (p == 1 ? lo.y : (p == 2 ? hi.x : hi.y));
    const uint32_t w1 = p == 0 ? lo.y : (p == 1 ? hi.x : hi.y);
    const uint32_t pair = __funnelshift_r(w0, w1, 16 * (j & 1));
    start = pair & 0xffffu;
    freq = (pair >> 16) - start;
    return 4 * g + j;
}
#endif

// q-score symbol search through the window rows: entry = u16[8] {s0, start(s0), ..., start(s0+6)} of the slot's bucket;
// start(s0) <= slot by construction.  The six symbols s0 .. s0+5 resolve in registers; a slot beyond them (seven symbol
// boundaries inside what is left of one 128-slot bucket) goes through the compact row.
__device__ __forceinline__ uint32_t q_find_win(const uint4* __restrict__ win_row, const uint8_t* __restrict__ row, uint32_t slot,
                                               uint32_t& start, uint32_t& freq) {
    const uint4 w = __ldg(win_row + (slot >> 7));
    const uint32_t K = (slot | 0x8000u) * 0x10001u;
    // flags of start(s0+1) .. start(s0+6) <= slot (start(s0) always is): monotone
    const uint32_t flags = ((K - w.y) & 0x80008000u) | (((K - w.z) & 0x80008000u) >> 1) | (((K - w.w) & 0x80008000u) >> 2);
    const uint32_t j = __popc(flags);  // 0 .. 6: the slot lies in symbol s0 + j
    if (j == 6) return q_find(row, slot, start, freq);  // start(s0+7) is not in the window
    const uint32_t i = j + 1;          // u16 index of start(s0+j) in the entry
    const uint32_t p = i >> 1;
    const uint32_t w0 = p == 0 ? w.x : (p == 1 ? w.y : (p == 2 ? w.z : w.w));
    const uint32_t w1 = p == 0 ? w.y : (p == 1 ? w.z : w.w);
    const uint32_t pair = __funnelshift_r(w0, w1, 16 * (i & 1));
    start = pair & 0xffffu;
    freq = (pair >> 16) - start;
    return (w.x & 0xffffu) + j;
}

}  // namespace idn
