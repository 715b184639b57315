// sha3.hpp -- SHA3-256 (FIPS 202), used for ModelIdentifier (model.rs:458-482 uses the `sha3` crate).
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace idencomp {

class Sha3_256 {
public:
    Sha3_256() { std::memset(st_, 0, sizeof st_); }

    void update(const void* data, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        while (n--) {
            reinterpret_cast<uint8_t*>(st_)[pos_++] ^= *p++;
            if (pos_ == kRate) {
                permute();
                pos_ = 0;
            }
        }
    }

    std::array<uint8_t, 32> finish() {
        uint8_t* b = reinterpret_cast<uint8_t*>(st_);
        b[pos_] ^= 0x06;  // SHA3 domain separation + first pad bit
        b[kRate - 1] ^= 0x80;
        permute();
        std::array<uint8_t, 32> out;
        std::memcpy(out.data(), st_, 32);
        return out;
    }

private:
    static constexpr size_t kRate = 136;  // 1088-bit rate for a 256-bit digest
    uint64_t st_[25];
    size_t pos_ = 0;

    static uint64_t rotl(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

    void permute() {  // Keccak-f[1600]; the state is kept little-endian (x86-64 / aarch64 hosts)
        static const uint64_t rc[24] = {
            0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
            0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
            0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
            0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
            0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
            0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
        static const int rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
        uint64_t* a = st_;
        for (int round = 0; round < 24; round++) {
            uint64_t c[5], d[5], b[25];
            for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
            for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
            for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
            for (int x = 0; x < 5; x++)
                for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y], rot[x + 5 * y]);
            for (int y = 0; y < 5; y++)
                for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
            a[0] ^= rc[round];
        }
    }
};

}  // namespace idencomp
