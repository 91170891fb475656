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
//
// The kernels that use this are bound by the ALU pipe, not by HBM, so the helpers are
// written for instruction count: strings are N 32-bit words (16 letters each, N = 3 at
// max_relator_length 36), variable shifts are a binary word-select stage + one funnel
// shift per word, masks come from clamped funnel shifts, and left shifts / adds are
// phrased so that ptxas can place them on the FMA pipe (IMAD) beside the ALU pipe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace acs {

// per-row status, shared with the C ABI (include/acsolver_b200.h)
enum : int { ST_OK = 0, ST_ASSERT = 1, ST_INDEX = 2 };

constexpr uint32_t kEven32 = 0x55555555u;  // bit 0 of every 2-bit group
constexpr uint32_t kOdd32 = 0xAAAAAAAAu;   // bit 1 of every group == the "negate" mask

// words needed for max_relator_length letters
__host__ __device__ constexpr int words_for(int mrl) { return (mrl + 15) / 16; }

template <int N>
struct Bits {
    uint32_t w[N];
};

template <int N>
__device__ __forceinline__ Bits<N> bz() {
    Bits<N> r;
#pragma unroll
    for (int k = 0; k < N; ++k) r.w[k] = 0;
    return r;
}
template <int N>
__device__ __forceinline__ bool is_zero(const Bits<N>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) o |= a.w[k];
    return o == 0;
}
template <int N>
__device__ __forceinline__ Bits<N> operator^(const Bits<N>& a, const Bits<N>& b) {
    Bits<N> r;
#pragma unroll
    for (int k = 0; k < N; ++k) r.w[k] = a.w[k] ^ b.w[k];
    return r;
}
template <int N>
__device__ __forceinline__ Bits<N> operator|(const Bits<N>& a, const Bits<N>& b) {
    Bits<N> r;
#pragma unroll
    for (int k = 0; k < N; ++k) r.w[k] = a.w[k] | b.w[k];
    return r;
}
template <int N>
__device__ __forceinline__ Bits<N> operator&(const Bits<N>& a, const Bits<N>& b) {
    Bits<N> r;
#pragma unroll
    for (int k = 0; k < N; ++k) r.w[k] = a.w[k] & b.w[k];
    return r;
}
template <int N>
__device__ __forceinline__ bool operator==(const Bits<N>& a, const Bits<N>& b) {
    return is_zero(a ^ b);
}

// ---- shifts -------------------------------------------------------------------------
// word-granular moves: x[j] <- x[j+S] (down) / x[j-S] (up), zero fill
template <int N, int S>
__device__ __forceinline__ void words_down_if(Bits<N>& x, bool p) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t src = (j + S < N) ? x.w[j + S] : 0u;
        x.w[j] = p ? src : x.w[j];
    }
}
template <int N, int S>
__device__ __forceinline__ void words_up_if(Bits<N>& x, bool p) {
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        const uint32_t src = (j - S >= 0) ? x.w[j - S] : 0u;
        x.w[j] = p ? src : x.w[j];
    }
}
// logical right shift by n bits, 0 <= n <= 32*N
template <int N>
__device__ __forceinline__ Bits<N> shr(Bits<N> x, int n) {
    const int q = n >> 5;
    if constexpr (N >= 4) words_down_if<N, 4>(x, (q & 4) != 0);
    if constexpr (N >= 2) words_down_if<N, 2>(x, (q & 2) != 0);
    words_down_if<N, 1>(x, (q & 1) != 0);
    Bits<N> r;
#pragma unroll
    for (int j = 0; j < N; ++j) r.w[j] = __funnelshift_r(x.w[j], (j + 1 < N) ? x.w[j + 1] : 0u, n);
    return r;
}
// logical left shift by n bits, 0 <= n <= 32*N
template <int N>
__device__ __forceinline__ Bits<N> shl(Bits<N> x, int n) {
    const int q = n >> 5;
    if constexpr (N >= 4) words_up_if<N, 4>(x, (q & 4) != 0);
    if constexpr (N >= 2) words_up_if<N, 2>(x, (q & 2) != 0);
    words_up_if<N, 1>(x, (q & 1) != 0);
    Bits<N> r;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) r.w[j] = __funnelshift_l((j > 0) ? x.w[j - 1] : 0u, x.w[j], n);
    return r;
}
template <int N, int K>
__device__ __forceinline__ Bits<N> shr_const(const Bits<N>& x) {
    Bits<N> r;
#pragma unroll
    for (int j = 0; j < N; ++j) r.w[j] = __funnelshift_r(x.w[j], (j + 1 < N) ? x.w[j + 1] : 0u, K);
    return r;
}
template <int N, int K>
__device__ __forceinline__ Bits<N> shl_const(const Bits<N>& x) {
    Bits<N> r;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) r.w[j] = __funnelshift_l((j > 0) ? x.w[j - 1] : 0u, x.w[j], K);
    return r;
}
// low n bits set, any n >= 0 (clamped funnel shift: min(n - 32j, 32) ones in word j)
template <int N>
__device__ __forceinline__ Bits<N> low_mask(int n) {
    Bits<N> r;
#pragma unroll
    for (int j = 0; j < N; ++j) r.w[j] = __funnelshift_lc(0xFFFFFFFFu, 0u, max(n - 32 * j, 0));
    return r;
}
// index of the lowest set bit; 32*N if none
template <int N>
__device__ __forceinline__ int ctz(const Bits<N>& a) {
    int r = 32 * N;
#pragma unroll
    for (int j = N - 1; j >= 0; --j) r = a.w[j] ? 32 * j + (__ffs((int)a.w[j]) - 1) : r;
    return r;
}

// A relator: len letters, codes beyond len are zero (canonical, so equality of
// (bits,len) pairs is equality of padded int8 rows).
template <int N>
struct Rel {
    Bits<N> b;
    int len;
};

// 2-bit code of an int8 letter in {+-1,+-2} and back
__device__ __forceinline__ uint32_t code_of(int8_t v) {
    return (((uint32_t)(uint8_t)v >> 6) & 2u) | ((uint32_t)v & 1u);
}
__device__ __forceinline__ int8_t letter_of(uint32_t c) {
    // 0 -> +2, 1 -> +1, 2 -> -2, 3 -> -1
    return (int8_t)((0xFFFE0102u >> (8 * c)) & 0xFFu);
}

// code of letter t (0 <= t < 16*N)
template <int N>
__device__ __forceinline__ uint32_t get_code(const Bits<N>& b, int t) {
    uint32_t w = b.w[0];
    const int q = t >> 4;
#pragma unroll
    for (int j = 1; j < N; ++j) w = (q == j) ? b.w[j] : w;
    return (w >> (2 * (t & 15))) & 3u;
}
// OR a 2-bit code into letter position t (the position must currently hold 00)
template <int N>
__device__ __forceinline__ void or_code(Bits<N>& b, int t, uint32_t c) {
    const uint32_t v = c << (2 * (t & 15));
    const int q = t >> 4;
#pragma unroll
    for (int j = 0; j < N; ++j) b.w[j] |= (q == j) ? v : 0u;
}

// reverse the order of all 16*N groups (group t -> 16N-1-t), each group negated
template <int N>
__device__ __forceinline__ Bits<N> reverse_negate_all(const Bits<N>& a) {
    Bits<N> r;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const uint32_t v = __brev(a.w[N - 1 - j]);  // groups reversed, bits inside a group swapped
        // swap back and flip the sign bit: new bit1 = ~old bit0 ... of the swapped pair
        r.w[j] = (((v & kEven32) * 2u) | ((v >> 1) & kEven32)) ^ kOdd32;
    }
    return r;
}
// inverse word: letters reversed and negated (ac_moves.py:43-48)
template <int N>
__device__ __forceinline__ Bits<N> inverse_bits(const Bits<N>& a, int len) {
    return shr<N>(reverse_negate_all<N>(a), 32 * N - 2 * len);  // the shift also clears >= len
}

// true iff no adjacent inverse pair (the word is freely reduced)
template <int N>
__device__ __forceinline__ bool is_freely_reduced(const Rel<N>& r) {
    if (r.len < 2) return true;
    const Bits<N> s = shr_const<N, 2>(r.b);
    const Bits<N> m = low_mask<N>(2 * (r.len - 1));
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint32_t q = r.b.w[k] ^ s.w[k] ^ kOdd32;        // group == 0  <=>  adjacent inverse pair
        any |= ~(q | (q >> 1)) & kEven32 & m.w[k];            // bit 2t set <=> letters t,t+1 cancel
    }
    return any == 0;
}

// general free reduction (utils.py:207-217).  Rare path: only caller-supplied words
// can be non-reduced, every word this library produces already is.
template <int N>
__device__ __noinline__ void free_reduce_slow(Rel<N>& r) {
    Bits<N> out = bz<N>();
    int top = 0;
    uint32_t last = 0;
    for (int t = 0; t < r.len; ++t) {
        const uint32_t c = get_code<N>(r.b, t);
        if (top > 0 && (last ^ c) == 2u) {
            --top;
            out = out & low_mask<N>(2 * top);
            last = top > 0 ? get_code<N>(out, top - 1) : 0u;
        } else {
            or_code<N>(out, top, c);
            ++top;
            last = c;
        }
    }
    r.b = out;
    r.len = top;
}

// cyclic reduction of a freely reduced word (utils.py:220-229)
template <int N>
__device__ __forceinline__ void cyclic_reduce(Rel<N>& r) {
    if (r.len < 2) return;
    // cheap exit: first and last letter are not inverse of each other
    if ((get_code<N>(r.b, 0) ^ get_code<N>(r.b, r.len - 1)) != 2u) return;
    const Bits<N> x = r.b ^ inverse_bits<N>(r.b, r.len);
    // w[p] == -w[L-1-p] for p < c  <=>  the first c groups of w and inverse(w) agree.
    // A non-empty freely reduced word differs from its inverse before the middle.
    int c = ctz<N>(x) >> 1;
    c = min(c, r.len >> 1);
    r.len -= 2 * c;
    r.b = shr<N>(r.b, 2 * c) & low_mask<N>(2 * r.len);
}

// utils.py:175-240 on one relator
template <int N>
__device__ __forceinline__ void simplify(Rel<N>& r, bool cyclical) {
    if (!is_freely_reduced<N>(r)) free_reduce_slow<N>(r);
    if (cyclical) cyclic_reduce<N>(r);
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
template <int N>
__device__ __forceinline__ bool concat(Rel<N>& u, const Rel<N>& w, bool invert, int mrl) {
    // junction: u[lu-1-t] == -v[t]  <=>  inverse(u)[t] == v[t]
    const Bits<N> iu = inverse_bits<N>(u.b, u.len);
    Bits<N> v = w.b;
    if (invert) v = inverse_bits<N>(w.b, w.len);
    const int k = min(ctz<N>(iu ^ v) >> 1, min(u.len, w.len));
    const int ns = u.len + w.len - 2 * k;
    if (ns > mrl) return false;
    const int keep = u.len - k;
    u.b = (u.b & low_mask<N>(2 * keep)) | shl<N>(shr<N>(v, 2 * k), 2 * keep);
    u.len = ns;
    return true;
}

// r_i <- g r_i g^{-1}, at most one letter cancelled at each end (ac_moves.py:113-154).
// Requires u.len > 0.  Returns true if rewritten.
//   s = r[0] == -g, e = r[-1] == g;  cand = [g]*(1-s) ++ r[s : len-e] ++ [-g]*(1-e)
template <int N>
__device__ __forceinline__ bool conjugate(Rel<N>& u, uint32_t g, int mrl) {
    const uint32_t first = u.b.w[0] & 3u;
    const uint32_t last = get_code<N>(u.b, u.len - 1);
    const bool s = first == (g ^ 2u);
    const bool e = last == g;
    const int ns = u.len + 2 - 2 * ((int)s + (int)e);
    if (ns > mrl) return false;
    Bits<N> t = u.b;
    if (e) t = t & low_mask<N>(2 * (u.len - 1));  // drop the last letter
    const Bits<N> down = shr_const<N, 2>(t);      // drop the first letter
    Bits<N> up = shl_const<N, 2>(t);              // make room for g
    up.w[0] |= g;
#pragma unroll
    for (int j = 0; j < N; ++j) t.w[j] = s ? down.w[j] : up.w[j];
    if (!e) or_code<N>(t, u.len + 1 - 2 * (int)s, g ^ 2u);
    u.b = t;
    u.len = ns;
    return true;
}

// Full ACMove on packed relators.  Returns ST_*.
//   TRUSTED = false: r0, r1 are arbitrary right-padded words (possibly non-reduced, possibly
//             empty): the reference's full validate + simplify of BOTH relators is evaluated.
//   TRUSTED = true : both words are already normal forms for this `cyclical` flag (freely
//             reduced, and cyclically reduced if cyclical) -- true for every state this
//             library produced with the same flag.  The trailing simplification is then the
//             identity on the untouched relator, a conjugation is a rotation or a no-op
//             under cyclic reduction, and only a concatenation needs the cyclic strip.
// `changed_other` reports whether the relator NOT targeted by the move was altered.
template <int N, bool TRUSTED>
__device__ __forceinline__ int apply_move(Rel<N>& r0, Rel<N>& r1, int id, int mrl, bool cyclical,
                                          bool& changed_other) {
    const bool tgt1 = ((id + 1) & 1) != 0;
    Rel<N> u, w;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        u.b.w[j] = tgt1 ? r1.b.w[j] : r0.b.w[j];
        w.b.w[j] = tgt1 ? r0.b.w[j] : r1.b.w[j];
    }
    u.len = tgt1 ? r1.len : r0.len;
    w.len = tgt1 ? r0.len : r1.len;
    changed_other = false;
    bool need_cyc = cyclical;
    if (id < 4) {
        concat<N>(u, w, id == 1 || id == 2, mrl);
    } else {
        if (u.len == 0) return ST_INDEX;  // relator_nonzero[0] on an empty array
        if (TRUSTED && cyclical) {
            // u is cyclically reduced: g u g^-1 reduces back to u unless an end cancels
            const uint32_t g = conj_code(id);
            const bool s = (u.b.w[0] & 3u) == (g ^ 2u);
            const bool e = get_code<N>(u.b, u.len - 1) == g;
            if (s | e) conjugate<N>(u, g, mrl);  // a rotation by one letter, same length
            need_cyc = false;
        } else {
            conjugate<N>(u, conj_code(id), mrl);
        }
    }
    if (u.len == 0 || w.len == 0) return ST_ASSERT;  // utils.py:261-263
    if (TRUSTED) {
        if (need_cyc) cyclic_reduce<N>(u);
    } else {
        simplify<N>(u, cyclical);
        const Rel<N> w_in = w;
        simplify<N>(w, cyclical);
        changed_other = !(w.len == w_in.len && w.b == w_in.b);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        r0.b.w[j] = tgt1 ? w.b.w[j] : u.b.w[j];
        r1.b.w[j] = tgt1 ? u.b.w[j] : w.b.w[j];
    }
    r0.len = tgt1 ? w.len : u.len;
    r1.len = tgt1 ? u.len : w.len;
    return ST_OK;
}

// true iff the pair is a normal form for `cyclical` (what TRUSTED assumes)
template <int N>
__device__ __forceinline__ bool is_normal_form(const Rel<N>& r, bool cyclical) {
    if (!is_freely_reduced<N>(r)) return false;
    if (cyclical && r.len >= 2 && (get_code<N>(r.b, 0) ^ get_code<N>(r.b, r.len - 1)) == 2u) return false;
    return true;
}

// ---- byte <-> packed, generic alignment (any mrl) ---------------------------------
template <int N>
__device__ __forceinline__ Rel<N> pack_bytes(const int8_t* p, int mrl) {
    Rel<N> r;
    r.b = bz<N>();
    int len = 0;
    for (int t = 0; t < mrl; ++t) {
        const int8_t v = p[t];
        if (v != 0) {
            or_code<N>(r.b, t, code_of(v));
            ++len;
        }
    }
    r.len = len;
    return r;
}
template <int N>
__device__ __forceinline__ void unpack_bytes(int8_t* p, const Rel<N>& r, int mrl) {
    for (int t = 0; t < mrl; ++t) p[t] = t < r.len ? letter_of(get_code<N>(r.b, t)) : (int8_t)0;
}

}  // namespace acs
