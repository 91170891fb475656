// ac_core.cuh -- the AC' move transition function on bit-packed relators (device code).
//
// What it computes (reference, paths relative to /root/reference):
//   ACMove                 ac_solver/envs/ac_moves.py:159-231
//   concatenate_relators   ac_solver/envs/ac_moves.py:4-76
//   conjugate              ac_solver/envs/ac_moves.py:79-156
//   simplify_presentation  ac_solver/envs/utils.py:243-280
//   simplify_relator       ac_solver/envs/utils.py:175-240
//
// How (B200-first, not a translation): one thread owns one presentation and keeps each
// relator in registers as a string of 2-bit codes, letter t at bits [2t, 2t+2):
//     y = +2 -> 00    x = +1 -> 01    y^-1 = -2 -> 10    x^-1 = -1 -> 11
// (code = sign bit << 1 | bit 0 of the int8 letter), so "formal inverse of a letter" is
// XOR 0b10, "inverse of a word" is a bit reversal + shift, junction cancellation and
// cyclic reduction are a count-trailing-zeros of an XOR, and concatenation is a shift +
// OR.  No loops, no divergence on the 12 move ids except a 2-way concat/conj split; the
// only loop is the rare general free reduction of a caller-supplied non-reduced word.
// W = number of 64-bit words per relator: W=1 holds 32 letters, W=2 holds 64.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace acs {

// per-row status, shared with the C ABI (include/acsolver_b200.h)
enum : int { ST_OK = 0, ST_ASSERT = 1, ST_INDEX = 2 };

template <int W>
struct Bits {
    uint64_t w[W];
};

template <int W>
__device__ __forceinline__ Bits<W> bz() {
    Bits<W> r;
#pragma unroll
    for (int k = 0; k < W; ++k) r.w[k] = 0;
    return r;
}
template <int W>
__device__ __forceinline__ bool is_zero(const Bits<W>& a) {
    uint64_t o = 0;
#pragma unroll
    for (int k = 0; k < W; ++k) o |= a.w[k];
    return o == 0;
}
template <int W>
__device__ __forceinline__ Bits<W> operator^(const Bits<W>& a, const Bits<W>& b) {
    Bits<W> r;
#pragma unroll
    for (int k = 0; k < W; ++k) r.w[k] = a.w[k] ^ b.w[k];
    return r;
}
template <int W>
__device__ __forceinline__ Bits<W> operator|(const Bits<W>& a, const Bits<W>& b) {
    Bits<W> r;
#pragma unroll
    for (int k = 0; k < W; ++k) r.w[k] = a.w[k] | b.w[k];
    return r;
}
template <int W>
__device__ __forceinline__ Bits<W> operator&(const Bits<W>& a, const Bits<W>& b) {
    Bits<W> r;
#pragma unroll
    for (int k = 0; k < W; ++k) r.w[k] = a.w[k] & b.w[k];
    return r;
}
template <int W>
__device__ __forceinline__ bool operator==(const Bits<W>& a, const Bits<W>& b) {
    return is_zero(a ^ b);
}

// logical shifts by n bits, 0 <= n <= 64*W (n == 64*W gives zero)
__device__ __forceinline__ uint64_t shl64(uint64_t x, int n) { return n >= 64 ? 0ull : x << n; }
__device__ __forceinline__ uint64_t shr64(uint64_t x, int n) { return n >= 64 ? 0ull : x >> n; }

template <int W>
__device__ __forceinline__ Bits<W> shl(const Bits<W>& a, int n) {
    Bits<W> r;
    if constexpr (W == 1) {
        r.w[0] = shl64(a.w[0], n);
    } else {
        if (n >= 64) {
            r.w[1] = shl64(a.w[0], n - 64);
            r.w[0] = 0;
        } else {
            r.w[1] = (a.w[1] << n) | (n ? (a.w[0] >> (64 - n)) : 0ull);
            r.w[0] = a.w[0] << n;
        }
    }
    return r;
}
template <int W>
__device__ __forceinline__ Bits<W> shr(const Bits<W>& a, int n) {
    Bits<W> r;
    if constexpr (W == 1) {
        r.w[0] = shr64(a.w[0], n);
    } else {
        if (n >= 64) {
            r.w[0] = shr64(a.w[1], n - 64);
            r.w[1] = 0;
        } else {
            r.w[0] = (a.w[0] >> n) | (n ? (a.w[1] << (64 - n)) : 0ull);
            r.w[1] = a.w[1] >> n;
        }
    }
    return r;
}
// low n bits set, 0 <= n <= 64*W
template <int W>
__device__ __forceinline__ Bits<W> low_mask(int n) {
    Bits<W> r;
    if constexpr (W == 1) {
        r.w[0] = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    } else {
        r.w[0] = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
        r.w[1] = n <= 64 ? 0ull : (n >= 128 ? ~0ull : ((1ull << (n - 64)) - 1ull));
    }
    return r;
}
// index of the lowest set bit; 64*W if none
template <int W>
__device__ __forceinline__ int ctz(const Bits<W>& a) {
    if constexpr (W == 1) {
        return a.w[0] ? __ffsll((long long)a.w[0]) - 1 : 64;
    } else {
        if (a.w[0]) return __ffsll((long long)a.w[0]) - 1;
        return a.w[1] ? 64 + __ffsll((long long)a.w[1]) - 1 : 128;
    }
}

constexpr uint64_t kEven = 0x5555555555555555ull;  // bit 0 of every 2-bit group
constexpr uint64_t kOdd = 0xAAAAAAAAAAAAAAAAull;   // bit 1 of every group == "negate" mask

// A relator: len letters, codes beyond len are zero (canonical, so equality of
// (bits,len) pairs is equality of padded int8 rows).
template <int W>
struct Rel {
    Bits<W> b;
    int len;
};

// 2-bit code of an int8 letter in {+-1,+-2} and back
__device__ __forceinline__ uint32_t code_of(int8_t v) {
    return ((uint32_t)(uint8_t)v >> 6 & 2u) | ((uint32_t)v & 1u);
}
__device__ __forceinline__ int8_t letter_of(uint32_t c) {
    // 0 -> +2, 1 -> +1, 2 -> -2, 3 -> -1
    return (int8_t)((0xFFFE0102u >> (8 * c)) & 0xFFu);
}

template <int W>
__device__ __forceinline__ uint32_t get_code(const Bits<W>& b, int t) {
    if constexpr (W == 1) return (uint32_t)(b.w[0] >> (2 * t)) & 3u;
    else return (uint32_t)((t < 32 ? b.w[0] >> (2 * t) : b.w[1] >> (2 * t - 64))) & 3u;
}

// inverse word: letters reversed and negated (ac_moves.py:43-48)
template <int W>
__device__ __forceinline__ Bits<W> inverse_bits(const Bits<W>& a, int len) {
    // bit-reverse the whole string: group t moves to group (32W-1-t) with its two bits
    // swapped; swap them back, right-align to len groups, flip the sign bit of each.
    Bits<W> r;
    if constexpr (W == 1) {
        r.w[0] = __brevll(a.w[0]);
    } else {
        r.w[0] = __brevll(a.w[1]);
        r.w[1] = __brevll(a.w[0]);
    }
#pragma unroll
    for (int k = 0; k < W; ++k) r.w[k] = ((r.w[k] & kEven) << 1) | ((r.w[k] >> 1) & kEven);
    r = shr<W>(r, 64 * W - 2 * len);
    Bits<W> neg;
#pragma unroll
    for (int k = 0; k < W; ++k) neg.w[k] = kOdd;
    return r ^ (neg & low_mask<W>(2 * len));
}

// true iff no adjacent inverse pair (the word is freely reduced)
template <int W>
__device__ __forceinline__ bool is_freely_reduced(const Rel<W>& r) {
    if (r.len < 2) return true;
    Bits<W> y = r.b ^ shr<W>(r.b, 2);  // group t = code[t] ^ code[t+1]
    Bits<W> z;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        uint64_t q = y.w[k] ^ kOdd;               // group == 0  <=>  adjacent inverse pair
        z.w[k] = ~(q | (q >> 1)) & kEven;         // bit 2t set <=> group t is zero
    }
    z = z & low_mask<W>(2 * (r.len - 1));
    return is_zero(z);
}

// general free reduction (utils.py:207-217).  Rare path: only caller-supplied words
// can be non-reduced, every word this library produces already is.
template <int W>
__device__ __noinline__ void free_reduce_slow(Rel<W>& r) {
    Bits<W> out = bz<W>();
    int top = 0;
    uint32_t last = 0;
    for (int t = 0; t < r.len; ++t) {
        uint32_t c = get_code<W>(r.b, t);
        if (top > 0 && (last ^ c) == 2u) {
            --top;
            out = out & low_mask<W>(2 * top);
            last = top > 0 ? get_code<W>(out, top - 1) : 0u;
        } else {
            Bits<W> cb = bz<W>();
            cb.w[0] = c;
            out = out | shl<W>(cb, 2 * top);
            ++top;
            last = c;
        }
    }
    r.b = out;
    r.len = top;
}

// cyclic reduction of a freely reduced word (utils.py:220-229)
template <int W>
__device__ __forceinline__ void cyclic_reduce(Rel<W>& r) {
    if (r.len < 2) return;
    Bits<W> x = r.b ^ inverse_bits<W>(r.b, r.len);
    // w[p] == -w[L-1-p] for p < c  <=>  the first c groups of w and inverse(w) agree.
    // A non-empty freely reduced word differs from its inverse before the middle.
    int c = ctz<W>(x) >> 1;
    c = min(c, r.len >> 1);
    if (c) {
        r.len -= 2 * c;
        r.b = shr<W>(r.b, 2 * c) & low_mask<W>(2 * r.len);
    }
}

// utils.py:175-240 on one relator
template <int W>
__device__ __forceinline__ void simplify(Rel<W>& r, bool cyclical) {
    if (!is_freely_reduced<W>(r)) free_reduce_slow<W>(r);
    if (cyclical) cyclic_reduce<W>(r);
}

// Decoded move id (table checked against the reference on all 12 ids):
//   ids 0..3  concat:  target = (id+1)&1, sign = + for ids 0,3 and - for ids 1,2
//   ids 4..11 conj:    target = (id+1)&1, generator code below
__device__ __forceinline__ uint32_t conj_code(int id) {
    // id:    4    5    6    7    8    9    10   11
    // g:    -1   -2   -2   +1   +1   +2   +2   -1
    // code:  3    2    2    1    1    0    0    3
    return (0xC16Bu >> (2 * (id - 4))) & 3u;
}

// r_i <- r_i r_j^{sign} with junction-only cancellation, accepted iff <= mrl letters
// (ac_moves.py:53-74).  Returns true if the relator was rewritten.
template <int W>
__device__ __forceinline__ bool concat(Rel<W>& u, const Rel<W>& w, bool invert, int mrl) {
    Rel<W> v;
    v.len = w.len;
    v.b = invert ? inverse_bits<W>(w.b, w.len) : w.b;
    // u[lu-1-t] == -v[t]  <=>  inverse(u)[t] == v[t]
    Bits<W> x = inverse_bits<W>(u.b, u.len) ^ v.b;
    int k = min(ctz<W>(x) >> 1, min(u.len, v.len));
    int ns = u.len + v.len - 2 * k;
    if (ns > mrl) return false;
    int keep = u.len - k;
    u.b = (u.b & low_mask<W>(2 * keep)) | shl<W>(shr<W>(v.b, 2 * k), 2 * keep);
    u.len = ns;
    return true;
}

// r_i <- g r_i g^{-1}, at most one letter cancelled at each end (ac_moves.py:113-154).
// Requires u.len > 0.  Returns true if rewritten.
template <int W>
__device__ __forceinline__ bool conjugate(Rel<W>& u, uint32_t g, int mrl) {
    uint32_t first = get_code<W>(u.b, 0);
    uint32_t last = get_code<W>(u.b, u.len - 1);
    int s = first == (g ^ 2u);
    int e = last == g;
    int ns = u.len + 2 - 2 * (s + e);
    if (ns > mrl) return false;
    int nmid = u.len - s - e;  // letters kept: u[s : len-e]  (>= 0; -1 impossible: s,e need 2 letters... see below)
    if (nmid < 0) nmid = 0;    // len==1 cannot have s and e both set, kept as a guard
    Bits<W> mid = shr<W>(u.b, 2 * s) & low_mask<W>(2 * nmid);
    int pos = nmid;
    if (!s) {
        mid = shl<W>(mid, 2);
        mid.w[0] |= g;
        pos += 1;
    }
    if (!e) {
        Bits<W> gb = bz<W>();
        gb.w[0] = g ^ 2u;
        mid = mid | shl<W>(gb, 2 * pos);
    }
    u.b = mid;
    u.len = ns;
    return true;
}

// Full ACMove on packed relators.  Returns ST_*.  `changed_other` reports whether the
// relator NOT targeted by the move was altered by the trailing simplification (only
// possible for caller-supplied, not-yet-normalised states).
template <int W>
__device__ __forceinline__ int apply_move(Rel<W>& r0, Rel<W>& r1, int id, int mrl, bool cyclical,
                                          bool& changed_other) {
    const bool tgt1 = ((id + 1) & 1) != 0;
    Rel<W> u = tgt1 ? r1 : r0;
    Rel<W> w = tgt1 ? r0 : r1;
    if (id < 4) {
        concat<W>(u, w, id == 1 || id == 2, mrl);
    } else {
        if (u.len == 0) return ST_INDEX;  // relator_nonzero[0] on an empty array
        conjugate<W>(u, conj_code(id), mrl);
    }
    if (u.len == 0 || w.len == 0) return ST_ASSERT;  // utils.py:261-263
    simplify<W>(u, cyclical);
    const Rel<W> w_in = w;
    simplify<W>(w, cyclical);
    changed_other = !(w.len == w_in.len && w.b == w_in.b);
    if (tgt1) {
        r1 = u;
        r0 = w;
    } else {
        r0 = u;
        r1 = w;
    }
    return ST_OK;
}

// ---- byte <-> packed, generic alignment (any mrl) ---------------------------------
template <int W>
__device__ __forceinline__ Rel<W> pack_bytes(const int8_t* p, int mrl) {
    Rel<W> r;
    r.b = bz<W>();
    int len = 0;
    for (int t = 0; t < mrl; ++t) {
        int8_t v = p[t];
        if (v != 0) {
            uint64_t c = code_of(v);
            if (W == 1 || t < 32) r.b.w[0] |= c << (2 * t);
            else r.b.w[W - 1] |= c << (2 * t - 64);
            ++len;
        }
    }
    r.len = len;
    return r;
}
template <int W>
__device__ __forceinline__ void unpack_bytes(int8_t* p, const Rel<W>& r, int mrl) {
    for (int t = 0; t < mrl; ++t) p[t] = t < r.len ? letter_of(get_code<W>(r.b, t)) : (int8_t)0;
}

}  // namespace acs
