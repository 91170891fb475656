// core_host.cpp -- TEST HARNESS: compiles the device headers csrc/ac_core.cuh and
// csrc/ac_pack.cuh for the HOST (g++; the CUDA intrinsics they use are shimmed below) so the
// packed move logic can be checked against the oracle on a machine without a GPU.  This is
// test infrastructure only; the product never runs this code path.
#include <algorithm>
#include <cstdint>
#include <cstring>

static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t n) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (n & 31));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (n & 31)) >> 32);
}
static inline uint32_t __funnelshift_lc(uint32_t lo, uint32_t hi, uint32_t n) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    const uint32_t s = n > 32 ? 32 : n;
    return (uint32_t)((s == 32 ? (v << 16) << 16 : (v << s)) >> 32);
}
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline int __ffs(int x) { return x ? __builtin_ctz((unsigned)x) + 1 : 0; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
using std::max;
using std::min;
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

#include "ac_core.cuh"
#include "ac_pack.cuh"

using namespace acs;

template <int N, bool T>
static void run_bytes(const int8_t* in, const uint8_t* act, int8_t* out, uint8_t* lens, uint8_t* status, int64_t n,
                      int mrl, int cyc) {
    for (int64_t r = 0; r < n; ++r) {
        const int8_t* p = in + r * 2 * mrl;
        int8_t* q = out + r * 2 * mrl;
        std::memcpy(q, p, 2 * mrl);
        Rel<N> r0 = pack_bytes<N>(p, mrl), r1 = pack_bytes<N>(p + mrl, mrl);
        bool co = false;
        int st = act[r] > 11 ? (int)ST_ASSERT : apply_move<N, T>(r0, r1, act[r], mrl, cyc != 0, co);
        status[r] = (uint8_t)st;
        if (st == ST_OK) {
            unpack_bytes<N>(q, r0, mrl);
            unpack_bytes<N>(q + mrl, r1, mrl);
            lens[2 * r] = (uint8_t)r0.len;
            lens[2 * r + 1] = (uint8_t)r1.len;
        }
    }
}

// word path: mirrors ac_step_words_kernel's per-thread body (pack_words / store of the target)
template <int NW, bool T>
static void run_words(const int8_t* in, const uint8_t* act, int8_t* out, uint8_t* lens, uint8_t* status, int64_t n,
                      int cyc) {
    constexpr int N = (NW + 3) / 4;
    const int mrl = 4 * NW;
    for (int64_t r = 0; r < n; ++r) {
        uint32_t w0[NW], w1[NW];
        std::memcpy(w0, in + r * 2 * mrl, 4 * NW);
        std::memcpy(w1, in + r * 2 * mrl + mrl, 4 * NW);
        std::memcpy(out + r * 2 * mrl, in + r * 2 * mrl, 2 * mrl);
        if (T && cyc && act[r] >= 4 && act[r] <= 11) {  // the kernel's byte-domain rotation path
            const bool tgt1 = ((act[r] + 1) & 1) != 0;
            uint32_t(&u)[NW] = tgt1 ? w1 : w0;
            uint32_t(&w)[NW] = tgt1 ? w0 : w1;
            const int lu = count_letters<NW>(u), lw = count_letters<NW>(w);
            int st = ST_OK;
            if (lu == 0) st = ST_INDEX;
            else if (lw == 0) st = ST_ASSERT;
            status[r] = (uint8_t)st;
            if (st != ST_OK) continue;
            lens[2 * r] = (uint8_t)(tgt1 ? lw : lu);
            lens[2 * r + 1] = (uint8_t)(tgt1 ? lu : lw);
            uint8_t* ub = reinterpret_cast<uint8_t*>(out + r * 2 * mrl + (tgt1 ? mrl : 0));
            int fix_pos;
            uint32_t fix_val;
            if (conj_rotate_words<NW>(u, lu, u[0] & 0xFFu, ub[lu - 1], conj_letter_byte(act[r]), fix_pos, fix_val)) {
                std::memcpy(ub, u, 4 * NW);
                if (fix_pos >= 0) ub[fix_pos] = (uint8_t)fix_val;
            }
            continue;
        }
        Rel<N> r0 = pack_words<NW, N>(w0), r1 = pack_words<NW, N>(w1);
        bool co = false;
        int st = act[r] > 11 ? (int)ST_ASSERT : apply_move<N, T>(r0, r1, act[r], mrl, cyc != 0, co);
        status[r] = (uint8_t)st;
        if (st != ST_OK) continue;
        lens[2 * r] = (uint8_t)r0.len;
        lens[2 * r + 1] = (uint8_t)r1.len;
        const bool tgt1 = ((act[r] + 1) & 1) != 0;
        for (int h = 0; h < 2; ++h) {
            if (!(h == (tgt1 ? 1 : 0) || co)) continue;
            const Rel<N>& t = h ? r1 : r0;
            uint32_t ow[NW + 1];
            for (int i = 0; i < (NW + 1) / 2; ++i) unpack_pair<N>(t, i, ow[2 * i], ow[2 * i + 1]);
            std::memcpy(out + r * 2 * mrl + h * mrl, ow, 4 * NW);
        }
    }
}

extern "C" int hostsim_moves(const int8_t* in, const uint8_t* act, int8_t* out, uint8_t* lens, uint8_t* status,
                             int64_t n, int mrl, int cyc, int trusted, int use_words) {
    if (use_words && mrl % 4 == 0) {
        switch (mrl / 4) {
#define C(k)                                                                   \
    case k:                                                                    \
        if (trusted) run_words<k, true>(in, act, out, lens, status, n, cyc);   \
        else run_words<k, false>(in, act, out, lens, status, n, cyc);          \
        return 0;
            C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15) C(16)
#undef C
        }
    }
    switch (words_for(mrl)) {
#define D(k)                                                                        \
    case k:                                                                         \
        if (trusted) run_bytes<k, true>(in, act, out, lens, status, n, mrl, cyc);   \
        else run_bytes<k, false>(in, act, out, lens, status, n, mrl, cyc);          \
        return 0;
        D(1) D(2) D(3) D(4)
#undef D
    }
    return -1;
}
