// msgpack_lite.hpp -- the subset of MessagePack that rmp-serde's compact encoding of SerializableModel uses
// (model_serializer.rs:11-72): arrays, strings, unsigned/signed ints, float32/float64, bin, nil, bool.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace idencomp {

struct MsgpackError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class MsgpackReader {
public:
    MsgpackReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
    bool at_end() const { return p_ == end_; }

    size_t read_array() {
        uint8_t t = byte();
        if ((t & 0xf0) == 0x90) return t & 0x0f;
        if (t == 0xdc) return be(2);
        if (t == 0xdd) return be(4);
        throw MsgpackError("expected an array");
    }

    std::string read_str() {
        uint8_t t = byte();
        size_t n;
        if ((t & 0xe0) == 0xa0) n = t & 0x1f;
        else if (t == 0xd9) n = be(1);
        else if (t == 0xda) n = be(2);
        else if (t == 0xdb) n = be(4);
        else throw MsgpackError("expected a string");
        need(n);
        std::string s(reinterpret_cast<const char*>(p_), n);
        p_ += n;
        return s;
    }

    uint64_t read_uint() {
        uint8_t t = byte();
        if (t <= 0x7f) return t;
        switch (t) {
            case 0xcc: return be(1);
            case 0xcd: return be(2);
            case 0xce: return be(4);
            case 0xcf: return be(8);
            case 0xd0: return nonneg((int8_t)be(1));
            case 0xd1: return nonneg((int16_t)be(2));
            case 0xd2: return nonneg((int32_t)be(4));
            case 0xd3: return nonneg((int64_t)be(8));
        }
        throw MsgpackError("expected an unsigned integer");
    }

    float read_f32() {
        uint8_t t = byte();
        if (t == 0xca) {
            uint32_t u = (uint32_t)be(4);
            float f;
            std::memcpy(&f, &u, 4);
            return f;
        }
        if (t == 0xcb) {
            uint64_t u = be(8);
            double d;
            std::memcpy(&d, &u, 8);
            return (float)d;
        }
        throw MsgpackError("expected a float");
    }

    // [u8; 32] arrives either as an array of ints or as bin
    void read_bytes(uint8_t* out, size_t n) {
        uint8_t t = peek();
        if (t == 0xc4 || t == 0xc5 || t == 0xc6) {
            byte();
            size_t m = be(t == 0xc4 ? 1 : t == 0xc5 ? 2 : 4);
            if (m != n) throw MsgpackError("unexpected bin length");
            need(n);
            std::memcpy(out, p_, n);
            p_ += n;
            return;
        }
        if (read_array() != n) throw MsgpackError("unexpected byte-array length");
        for (size_t i = 0; i < n; i++) out[i] = (uint8_t)read_uint();
    }

private:
    const uint8_t* p_;
    const uint8_t* end_;
    void need(size_t n) const {
        if ((size_t)(end_ - p_) < n) throw MsgpackError("truncated msgpack data");
    }
    uint8_t peek() const {
        need(1);
        return *p_;
    }
    uint8_t byte() {
        need(1);
        return *p_++;
    }
    uint64_t be(int n) {
        need((size_t)n);
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v = (v << 8) | *p_++;
        return v;
    }
    static uint64_t nonneg(int64_t v) {
        if (v < 0) throw MsgpackError("negative integer where unsigned expected");
        return (uint64_t)v;
    }
};

}  // namespace idencomp
