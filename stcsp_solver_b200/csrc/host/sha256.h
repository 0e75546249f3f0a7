// Streaming SHA-256 (FIPS 180-4): canonical automata are hashed line by line, so a partialorder_20-size
// automaton (10 GB of canonical text) never has to exist as one string.  Checked against hashlib in tests/test_cabi.py.
#ifndef STCSP_HOST_SHA256_H
#define STCSP_HOST_SHA256_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#if defined(__x86_64__) && defined(__GNUC__)
#include <cpuid.h>
#include <immintrin.h>
#define STCSP_SHA_NI 1
#endif

namespace stcsp {

class Sha256 {
  public:
    Sha256() { reset(); }
    void reset() {
        static const uint32_t init[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au,
                                         0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
        memcpy(h_, init, sizeof h_);
        fill_ = 0;
        total_ = 0;
    }
    void update(const void *data, size_t n) {
        const unsigned char *p = (const unsigned char *)data;
        total_ += n;
        if (fill_ == 0 && n >= 64) {        // whole blocks straight from the caller's buffer
            const size_t nb = n / 64;
            blocks(p, nb);
            p += nb * 64;
            n -= nb * 64;
        }
        while (n > 0) {
            size_t take = 64 - fill_;
            if (take > n) take = n;
            memcpy(buf_ + fill_, p, take);
            fill_ += take;
            p += take;
            n -= take;
            if (fill_ == 64) {
                blocks(buf_, 1);
                fill_ = 0;
                if (n >= 64) {
                    const size_t nb = n / 64;
                    blocks(p, nb);
                    p += nb * 64;
                    n -= nb * 64;
                }
            }
        }
    }
    std::string hex() {
        uint64_t bits = total_ * 8;
        unsigned char pad = 0x80;
        update(&pad, 1);
        unsigned char zero = 0;
        while (fill_ != 56) update(&zero, 1);
        unsigned char len[8];
        for (int i = 0; i < 8; i++) len[i] = (unsigned char)(bits >> (56 - 8 * i));
        update(len, 8);
        static const char *digits = "0123456789abcdef";
        std::string out;
        for (int i = 0; i < 8; i++)
            for (int s = 28; s >= 0; s -= 4) out += digits[(h_[i] >> s) & 15];
        return out;
    }

  private:
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    static const uint32_t *round_constants() {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
            0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
            0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
            0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
            0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        return K;
    }
    void blocks(const unsigned char *p, size_t nb) {
#ifdef STCSP_SHA_NI
        static const bool ni = have_sha_ni() && getenv("STCSP_NO_SHA_NI") == nullptr;      // (the variable: tests of the portable rounds)
        if (ni) { blocks_ni(p, nb); return; }
#endif
        for (size_t i = 0; i < nb; i++) block(p + 64 * i);
    }
#ifdef STCSP_SHA_NI
    // the x86 SHA extensions (CPUID.7.0:EBX bit 29) hash ~5x faster than the portable rounds below; same digest
    static bool have_sha_ni() {
        unsigned a, b, c, d;
        if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
        const bool sha = (b >> 29) & 1u;
        if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
        return sha && ((c >> 19) & 1u) && ((c >> 9) & 1u);      // + SSE4.1, SSSE3
    }
    __attribute__((target("sha,sse4.1,ssse3"))) void blocks_ni(const unsigned char *p, size_t nb) {
        const uint32_t *K = round_constants();
        const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bll, 0x0405060700010203ll);
        __m128i t = _mm_loadu_si128((const __m128i *)&h_[0]);       // a b c d
        __m128i s1 = _mm_loadu_si128((const __m128i *)&h_[4]);      // e f g h
        t = _mm_shuffle_epi32(t, 0xB1);                             // c d a b
        s1 = _mm_shuffle_epi32(s1, 0x1B);                           // e f g h reversed
        __m128i s0 = _mm_alignr_epi8(t, s1, 8);                     // a b e f
        s1 = _mm_blend_epi16(s1, t, 0xF0);                          // c d g h
        for (; nb; nb--, p += 64) {
            const __m128i save0 = s0, save1 = s1;
            __m128i m[4];
            for (int i = 0; i < 4; i++) m[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(p + 16 * i)), bswap);
#pragma GCC unroll 16
            for (int r = 0; r < 16; r++) {
                // four rounds per step; m[r & 3] holds w[4r .. 4r+3]
                __m128i msg = _mm_add_epi32(m[r & 3], _mm_loadu_si128((const __m128i *)(K + 4 * r)));
                s1 = _mm_sha256rnds2_epu32(s1, s0, msg);
                msg = _mm_shuffle_epi32(msg, 0x0E);
                s0 = _mm_sha256rnds2_epu32(s0, s1, msg);
                if (r < 12) {
                    // schedule w[4(r+4) .. 4(r+4)+3] into the slot that is no longer needed
                    __m128i x = _mm_sha256msg1_epu32(m[r & 3], m[(r + 1) & 3]);
                    x = _mm_add_epi32(x, _mm_alignr_epi8(m[(r + 3) & 3], m[(r + 2) & 3], 4));
                    m[r & 3] = _mm_sha256msg2_epu32(x, m[(r + 3) & 3]);
                }
            }
            s0 = _mm_add_epi32(s0, save0);
            s1 = _mm_add_epi32(s1, save1);
        }
        t = _mm_shuffle_epi32(s0, 0x1B);                            // f e b a
        s1 = _mm_shuffle_epi32(s1, 0xB1);                           // d c h g
        s0 = _mm_blend_epi16(t, s1, 0xF0);                          // a b c d
        s1 = _mm_alignr_epi8(s1, t, 8);                             // e f g h
        _mm_storeu_si128((__m128i *)&h_[0], s0);
        _mm_storeu_si128((__m128i *)&h_[4], s1);
    }
#endif
    void block(const unsigned char *p) {
        const uint32_t *K = round_constants();
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h_[0], b = h_[1], c = h_[2], d = h_[3], e = h_[4], f = h_[5], g = h_[6], h = h_[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = h + S1 + ch + K[i] + w[i];
            uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h_[0] += a; h_[1] += b; h_[2] += c; h_[3] += d; h_[4] += e; h_[5] += f; h_[6] += g; h_[7] += h;
    }
    uint32_t h_[8];
    unsigned char buf_[64];
    size_t fill_;
    uint64_t total_;
};

}  // namespace stcsp

#endif
