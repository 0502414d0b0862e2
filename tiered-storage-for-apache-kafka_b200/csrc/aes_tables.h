// aes_tables.h — AES S-box and the Te0 round table, generated at compile time (FIPS-197 §5.1).
// Te0[x] = {02·S[x], S[x], S[x], 03·S[x]} packed little-endian (byte 0 = row 0), matching a little-endian
// load of a state column.  Te1..Te3 are byte rotations of Te0 and are formed with PRMT in the kernel.
#pragma once
#include <stdint.h>

namespace ts {

struct AesTables {
    uint8_t sbox[256];
    uint32_t te0[256];
};

constexpr uint8_t aes_xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }

constexpr AesTables make_aes_tables() {
    AesTables t{};
    uint8_t p = 1, q = 1;
    do {
        p = (uint8_t)(p ^ (uint8_t)(p << 1) ^ ((p & 0x80) ? 0x1b : 0));
        q = (uint8_t)(q ^ (q << 1)); q = (uint8_t)(q ^ (q << 2)); q = (uint8_t)(q ^ (q << 4));
        if (q & 0x80) q = (uint8_t)(q ^ 0x09);
        uint8_t x = (uint8_t)(q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6)) ^
                              (uint8_t)((q << 3) | (q >> 5)) ^ (uint8_t)((q << 4) | (q >> 4)));
        t.sbox[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    t.sbox[0] = 0x63;
    for (int i = 0; i < 256; i++) {
        uint8_t s = t.sbox[i];
        uint8_t s2 = aes_xtime(s);
        uint8_t s3 = (uint8_t)(s2 ^ s);
        t.te0[i] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
    }
    return t;
}

// AES-256 key expansion into 60 little-endian words (word i = bytes 4i..4i+3 of the FIPS-197 schedule).
// Runs on the host inside tsgpu_transform/detransform (240 bytes of set-up per call, passed to the
// kernels by value so round keys are read from the constant bank).
struct Aes256RoundKeys { uint32_t w[60]; };

inline Aes256RoundKeys aes256_expand_key(const uint8_t key[32]) {
    static const AesTables T = make_aes_tables();
    uint8_t rk[240];
    for (int i = 0; i < 32; i++) rk[i] = key[i];
    uint8_t rcon = 1;
    for (int i = 32; i < 240; i += 4) {
        uint8_t t[4] = { rk[i - 4], rk[i - 3], rk[i - 2], rk[i - 1] };
        if (i % 32 == 0) {
            uint8_t u = t[0];
            t[0] = (uint8_t)(T.sbox[t[1]] ^ rcon); t[1] = T.sbox[t[2]]; t[2] = T.sbox[t[3]]; t[3] = T.sbox[u];
            rcon = aes_xtime(rcon);
        } else if (i % 32 == 16) {
            for (int k = 0; k < 4; k++) t[k] = T.sbox[t[k]];
        }
        for (int k = 0; k < 4; k++) rk[i + k] = (uint8_t)(rk[i - 32 + k] ^ t[k]);
    }
    Aes256RoundKeys out;
    for (int i = 0; i < 60; i++)
        out.w[i] = (uint32_t)rk[4 * i] | ((uint32_t)rk[4 * i + 1] << 8) | ((uint32_t)rk[4 * i + 2] << 16) |
                   ((uint32_t)rk[4 * i + 3] << 24);
    return out;
}

}  // namespace ts
