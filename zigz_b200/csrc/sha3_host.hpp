// Host-side SHA3-256 for the parts of the path the reference keeps on the CPU:
//   - FiatShamirTranscript            /root/reference/src/core/hash.zig:255-324 (streaming state + "peek" finalisation)
//   - LassoProver.commitToPolynomial  /root/reference/src/lookups/lasso_prover.zig:242-252 (one long sequential sponge)
//   - SimpleMerkleTree.verify         /root/reference/src/commitments/merkle_tree.zig:362-373
// The reference gets SHA3 from Zig's std (std.crypto.hash.sha3.Sha3_256); this is FIPS 202 written for x86-64:
// whole-lane absorption of 8-byte words (the hot case: every message on the path is a sequence of le64 words),
// Keccak-f[1600] unrolled two rounds at a time so the pi permutation is pure register renaming.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace zigz {

class Sha3_256 {
  public:
    static constexpr size_t RATE = 136;
    Sha3_256() { reset(); }
    void reset() {
        memset(a_, 0, sizeof(a_));
        pos_ = 0;
    }
    void update(const void *data, size_t len) {
        const uint8_t *p = static_cast<const uint8_t *>(data);
        // byte-wise until lane aligned
        while (len && (pos_ & 7)) {
            xor_byte(*p++);
            len--;
        }
        // whole lanes (long aligned runs take the word path, which may use the AVX-512 absorb loop)
        if (len >= 8 * 17 * 4 && (pos_ & 7) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0) {
            const size_t words = len / 8;
            update_words(reinterpret_cast<const uint64_t *>(p), words);
            p += words * 8;
            len -= words * 8;
        }
        while (len >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            a_[pos_ >> 3] ^= w;
            pos_ += 8;
            if (pos_ == RATE) {
                permute(a_);
                pos_ = 0;
            }
            p += 8;
            len -= 8;
        }
        while (len) {
            xor_byte(*p++);
            len--;
        }
    }
    // absorb n little-endian u64 words (field elements as the reference serialises them)
    void update_words(const uint64_t *w, size_t n) {
        if (pos_ & 7) {
            update(w, n * 8);
            return;
        }
        if (pos_ != 0 && n >= 17 * 5) { // finish the partial block first so the bulk starts block-aligned
            size_t lane0 = pos_ >> 3;
            while (lane0 != 0) {
                a_[lane0++] ^= *w++;
                n--;
                if (lane0 == RATE / 8) {
                    permute(a_);
                    lane0 = 0;
                }
            }
            pos_ = 0;
        }
        if (pos_ == 0 && n >= 17 * 4 && have_avx512()) { // long runs: whole blocks through the AVX-512 absorb loop
            const size_t blocks = n / 17;
            absorb_blocks_avx512(a_, w, blocks);
            w += blocks * 17;
            n -= blocks * 17;
        }
        size_t lane = pos_ >> 3;
        for (size_t i = 0; i < n; i++) {
            a_[lane++] ^= w[i];
            if (lane == RATE / 8) {
                permute(a_);
                lane = 0;
            }
        }
        pos_ = lane << 3;
    }
    // absorb n field elements stored as canonical u32 (the device representation), each as its 8-byte LE encoding
    void update_words_u32(const uint32_t *w, size_t n) {
        while (n && (pos_ & 7)) { // not lane aligned: byte path
            uint64_t v = *w++;
            update(&v, 8);
            n--;
        }
        size_t lane = pos_ >> 3;
        while (n && lane != 0) { // finish the partial block
            a_[lane++] ^= *w++;
            n--;
            if (lane == RATE / 8) {
                permute(a_);
                lane = 0;
            }
        }
        if (lane == 0 && n >= 17 * 4 && have_avx512()) {
            const size_t blocks = n / 17;
            absorb_blocks_u32_avx512(a_, w, blocks);
            w += blocks * 17;
            n -= blocks * 17;
        }
        for (size_t i = 0; i < n; i++) {
            a_[lane++] ^= w[i];
            if (lane == RATE / 8) {
                permute(a_);
                lane = 0;
            }
        }
        pos_ = lane << 3;
    }
    // digest of everything absorbed so far WITHOUT disturbing the running state (hash.zig:304-306 clones the hasher)
    void peek(uint8_t out[32]) const {
        uint64_t s[25];
        memcpy(s, a_, sizeof(s));
        s[pos_ >> 3] ^= (uint64_t)0x06 << (8 * (pos_ & 7));
        s[16] ^= 0x8000000000000000ull;
        permute(s);
        memcpy(out, s, 32);
    }
    static void hash(const void *data, size_t len, uint8_t out[32]) {
        Sha3_256 h;
        h.update(data, len);
        h.peek(out);
    }
    static void permute(uint64_t s[25]);
    static bool have_avx512();
    static void absorb_blocks_avx512(uint64_t s[25], const uint64_t *words, size_t nblocks);
    static void absorb_blocks_u32_avx512(uint64_t s[25], const uint32_t *words, size_t nblocks);

  private:
    void xor_byte(uint8_t b) {
        a_[pos_ >> 3] ^= (uint64_t)b << (8 * (pos_ & 7));
        if (++pos_ == RATE) {
            permute(a_);
            pos_ = 0;
        }
    }
    uint64_t a_[25];
    size_t pos_;
};

} // namespace zigz
