// ac_pack.cuh -- int8 letter words <-> 2-bit code strings, four / eight letters at a time.
//
// The padded int8 row layout is the reference's (ac_solver/envs/utils.py:4-7); the packed
// form is ac_core.cuh's.  Requires letters in {0, +-1, +-2} with zeros on the right.
#pragma once
#include <cstdint>

#include "ac_core.cuh"

namespace acs {

// four letters (one 32-bit word) -> their four codes in the TOP byte of the result.
// Two multiply-adds gather the magnitude plane (bit 0 of each byte -> bits 24,26,28,30)
// and the sign plane (bit 7 -> bits 25,27,29,31); all stray partial products land on
// distinct bits below the window, so there are no carries.  IMAD runs on the FMA pipe,
// which this ALU-bound kernel leaves idle.
__device__ __forceinline__ uint32_t codes_top8(uint32_t w) {
    return (w & 0x01010101u) * 0x01041040u + (w & 0x80808080u) * 0x00041041u;
}
// top bytes of four products -> one 32-bit word of 16 codes
__device__ __forceinline__ uint32_t gather4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
    const uint32_t a = __byte_perm(p0, p1, 0x0073);
    const uint32_t b = __byte_perm(p2, p3, 0x0073);
    return __byte_perm(a, b, 0x5410);
}

// NW int8 words (4*NW letters) -> packed relator of N = ceil(NW/4) code words
// COUNT = false skips the letter count (the caller knows the length and sets r.len).
template <int NW, int N, bool COUNT = true>
__device__ __forceinline__ Rel<N> pack_words(const uint32_t (&w)[NW]) {
    Rel<N> r;
    uint32_t nz = 0;  // per byte lane: 2 * (number of non-zero letters seen in that lane)
#pragma unroll
    for (int q = 0; q < N; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = 4 * q + k;
            if (j < NW) {
                p[k] = codes_top8(w[j]);
                if (COUNT) nz += (w[j] | (w[j] * 2u)) & 0x02020202u;  // letters are 0 or have bit0|bit1 set
            } else {
                p[k] = 0;
            }
        }
        r.b.w[q] = gather4(p[0], p[1], p[2], p[3]);
    }
    r.len = COUNT ? (int)((nz * 0x01010101u) >> 25) : 0;
    return r;
}

// int8 words 2i and 2i+1 (letters 8i..8i+7) of a packed relator, zero beyond len
template <int N>
__device__ __forceinline__ void unpack_pair(const Rel<N>& r, int i, uint32_t& lo, uint32_t& hi) {
    const uint32_t c16 = (i & 1) ? (r.b.w[i >> 1] >> 16) : (r.b.w[i >> 1] & 0xFFFFu);
    // spread eight 2-bit codes to eight selector nibbles
    uint32_t t = (c16 | (c16 << 8)) & 0x00FF00FFu;
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;
    const uint32_t b0 = __byte_perm(0xFFFE0102u, 0u, t);  // code -> letter table lookup
    const uint32_t b1 = __byte_perm(0xFFFE0102u, 0u, t >> 16);
    const int n = 8 * r.len - 64 * i;
    lo = b0 & __funnelshift_lc(0xFFFFFFFFu, 0u, max(n, 0));  // low min(n,32) bits kept
    hi = b1 & __funnelshift_lc(0xFFFFFFFFu, 0u, max(n - 32, 0));
}

// number of non-zero letters in NW int8 words (zeros are on the right)
template <int NW>
__device__ __forceinline__ int count_letters(const uint32_t (&w)[NW]) {
    uint32_t nz = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) nz += (w[j] | (w[j] * 2u)) & 0x02020202u;
    return (int)((nz * 0x01010101u) >> 25);
}

// int8 value (as a byte) of the conjugating letter of move ids 4..11:
//   id 4: x^-1, 5: y^-1, 6: y^-1, 7: x, 8: x, 9: y, 10: y, 11: x^-1
__device__ __forceinline__ uint32_t conj_letter_byte(int id) {
    const uint32_t lo = 0x01FEFEFFu, hi = 0xFF020201u;  // ids 4..7 | ids 8..11
    return (((id & 4) ? lo : hi) >> (8 * (id & 3))) & 0xFFu;
}

// Conjugation by the letter g (byte value) of a freely AND cyclically reduced word under
// cyclic reduction, in the BYTE domain: g u g^-1 reduces back to u unless u starts with
// g^-1 (then u rotates left by one letter) or ends with g (rotates right), see
// ac_core.cuh apply_move<TRUSTED>.  u: NW int8 words, len > 0 letters, first/last its end
// letters.  Returns true if the word changed; the caller then stores u and finally writes
// the single byte fix_val at letter position fix_pos (fix_pos < 0: nothing to patch).
template <int NW>
__device__ __forceinline__ bool conj_rotate_words(uint32_t (&u)[NW], int len, uint32_t first, uint32_t last,
                                                  uint32_t g, int& fix_pos, uint32_t& fix_val) {
    const bool s = first == ((0u - g) & 0xFFu);
    const bool e = !s && last == g;
    const uint32_t sh = s ? 8u : (e ? 24u : 0u);
    uint32_t o[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        const uint32_t prev = j ? u[j - 1] : (g << 24);
        const uint32_t next = (j + 1 < NW) ? u[j + 1] : 0u;
        o[j] = __funnelshift_r(e ? prev : u[j], s ? next : u[j], sh);
    }
#pragma unroll
    for (int j = 0; j < NW; ++j) u[j] = o[j];
    fix_pos = s ? len - 1 : ((e && len < 4 * NW) ? len : -1);
    fix_val = s ? first : 0u;
    return s | e;
}

}  // namespace acs
