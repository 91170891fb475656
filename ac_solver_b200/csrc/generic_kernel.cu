// generic_kernel.cu -- byte-domain AC moves for ANY int8 alphabet, one thread per row.
//
// This is the general path behind the reference's single-call API (ACMove,
// concatenate_relators, conjugate, simplify_relator, simplify_presentation) where the
// reference accepts arbitrary integer letters (its unit tests use +-3, +-4:
// /root/reference/tests/test_ac_env.py:18-83,184-477).  It follows the reference's array
// semantics literally -- "filter the non-zero letters, rewrite one half, validate, reduce
// both halves, pad" (ac_moves.py:36-74,108-154, utils.py:175-280) -- so it also defines
// the behaviour for rows the packed fast path refuses (letters outside {+-1,+-2}).
// It is not a performance path: rows live in local memory and are walked letter by letter.
#include <cstdint>
#include <cuda_runtime.h>

#include "acs_internal.h"

namespace acs {

constexpr int kMaxWidth = 256;

__device__ static int gather_nz(const int8_t* p, int n, int8_t* dst) {
    int c = 0;
    for (int t = 0; t < n; ++t)
        if (p[t] != 0) dst[c++] = p[t];
    return c;
}
__device__ static int count_nz(const int8_t* p, int n) {
    int c = 0;
    for (int t = 0; t < n; ++t) c += p[t] != 0;
    return c;
}
// utils.py:13-54 on an array of 2*mrl letters
__device__ static bool valid_presentation(const int8_t* p, int mrl) {
    int l0 = count_nz(p, mrl), l1 = count_nz(p + mrl, mrl);
    if (l0 == 0 || l1 == 0) return false;
    for (int t = l0; t < mrl; ++t)
        if (p[t] != 0) return false;
    for (int t = l1; t < mrl; ++t)
        if (p[mrl + t] != 0) return false;
    return true;
}
// utils.py:207-229: stack free reduction (unique normal form) + optional cyclic strip
__device__ static int simplify_word(int8_t* w, int n, bool cyclical) {
    int top = 0;
    for (int t = 0; t < n; ++t) {
        int8_t c = w[t];
        if (top > 0 && w[top - 1] == (int8_t)(-c)) --top;
        else w[top++] = c;
    }
    n = top;
    if (cyclical && n > 0) {
        int pos = 0;
        while (pos < n && w[pos] == (int8_t)(-w[n - pos - 1])) ++pos;
        if (pos) {
            for (int t = 0; t < n - 2 * pos; ++t) w[t] = w[t + pos];
            n -= 2 * pos;
        }
    }
    return n;
}
// ac_moves.py:36-74; returns new size or -1 (rejected)
__device__ static int concat_raw(int8_t* P, int mrl, int i, int sign) {
    const int j = 1 - i;
    int8_t u[kMaxWidth / 2], v[kMaxWidth / 2];
    int lu = gather_nz(P + i * mrl, mrl, u);
    int lv = 0;
    if (sign == 1) lv = gather_nz(P + j * mrl, mrl, v);
    else
        for (int t = mrl - 1; t >= 0; --t)
            if (P[j * mrl + t] != 0) v[lv++] = (int8_t)(-P[j * mrl + t]);
    int acc = 0;
    const int lim = min(lu, lv);
    while (acc < lim && u[lu - 1 - acc] == (int8_t)(-v[acc])) ++acc;
    const int ns = lu + lv - 2 * acc;
    if (ns > mrl) return -1;
    int8_t* d = P + i * mrl;
    int q = 0;
    for (int t = 0; t < lu - acc; ++t) d[q++] = u[t];
    for (int t = acc; t < lv; ++t) d[q++] = v[t];
    for (; q < mrl; ++q) d[q] = 0;
    return ns;
}
// ac_moves.py:108-154; returns new size, -1 (rejected) or -2 (IndexError on an empty relator)
__device__ static int conj_raw(int8_t* P, int mrl, int i, int g) {
    int8_t rel[kMaxWidth / 2];
    const int size = gather_nz(P + i * mrl, mrl, rel);
    if (size == 0) return -2;
    const int s = rel[0] == (int8_t)(-g), e = rel[size - 1] == (int8_t)g;
    const int ns = size + 2 - 2 * (s + e);
    if (ns > mrl) return -1;
    int8_t* d = P + i * mrl;
    for (int t = s; t < size - e; ++t) d[1 - s + (t - s)] = rel[t];
    if (!s) d[0] = (int8_t)g;
    if (!e) d[size + 1 - 2 * s] = (int8_t)(-g);
    if (s && e)
        for (int t = ns; t < ns + 2 && t < mrl; ++t) d[t] = 0;
    return ns;
}

__global__ void __launch_bounds__(128) ac_generic_kernel(const GenericParams G) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= G.n) return;
    int8_t P[kMaxWidth];
    const int Wd = G.width;
    for (int t = 0; t < Wd; ++t) P[t] = G.in[row * Wd + t];
    int status = 0;
    if (G.op == OP_SIMPLIFY_RELATOR) {
        const int n = count_nz(P, Wd);
        for (int t = n; t < Wd; ++t)
            if (P[t] != 0) status = 1;  // "expect all zeros to be at the right end"
        int m = 0;
        if (!status) {
            m = simplify_word(P, n, G.cyclical != 0);
            for (int t = m; t < Wd; ++t) P[t] = 0;
        }
        G.aux[row] = m;
    } else {
        const int mrl = Wd / 2;
        int ns = 0;
        if (G.op == OP_ACMOVE) {
            const int id = G.action[row];
            if (id > 11) status = 1;
            else {
                const int m = id + 1, i = m & 1;
                if (id < 4) {
                    concat_raw(P, mrl, i, (((m - i) / 2) & 1) ? -1 : 1);  // ac_moves.py:192-198
                } else {
                    const int jp = ((m - i) / 2) & 1;                       // ac_moves.py:199-206
                    const int sp = ((m - i - 2 * jp) / 4) & 1;
                    if (conj_raw(P, mrl, i, (sp ? -1 : 1) * (jp + 1)) == -2) status = 2;
                }
            }
        } else if (G.op == OP_CONCAT_RAW) {
            ns = concat_raw(P, mrl, G.i, G.sign);
            G.aux[row] = ns;
        } else if (G.op == OP_CONJ_RAW) {
            ns = conj_raw(P, mrl, G.i, G.sign * G.j);
            G.aux[row] = ns;
        }
        if (G.op == OP_ACMOVE || G.op == OP_SIMPLIFY_PRESENTATION) {
            if (!status && !valid_presentation(P, mrl)) status = 1;  // utils.py:261-263
            if (!status) {
                for (int k = 0; k < 2; ++k) {
                    int8_t* r = P + k * mrl;
                    const int n = count_nz(r, mrl);
                    const int m = simplify_word(r, n, G.cyclical != 0);
                    for (int t = m; t < mrl; ++t) r[t] = 0;
                    G.aux[2 * row + k] = m;
                }
            }
        }
    }
    if (G.status) G.status[row] = (uint8_t)status;
    if (!status)
        for (int t = 0; t < Wd; ++t) G.out[row * Wd + t] = P[t];
    else
        for (int t = 0; t < Wd; ++t) G.out[row * Wd + t] = G.in[row * Wd + t];
}

// flags: bit0 = is_array_valid_presentation, bit1 = every letter in {0,+-1,+-2},
//        bit2 = zeros only on the right of each half (empty relators allowed)
__global__ void __launch_bounds__(128) ac_validate_kernel(const int8_t* in, uint8_t* flags, int64_t n, int mrl) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int8_t* p = in + row * 2 * mrl;
    bool alpha = true, padded = true;
    int len[2];
    for (int k = 0; k < 2; ++k) {
        int l = 0;
        bool seen_zero = false;
        for (int t = 0; t < mrl; ++t) {
            const int v = p[k * mrl + t];
            if (v < -2 || v > 2) alpha = false;
            if (v == 0) seen_zero = true;
            else {
                ++l;
                if (seen_zero) padded = false;
            }
        }
        len[k] = l;
    }
    const bool valid = padded && len[0] > 0 && len[1] > 0;
    flags[row] = (uint8_t)((valid ? 1 : 0) | (alpha ? 2 : 0) | (padded ? 4 : 0));
}

cudaError_t launch_generic(const GenericParams& P, cudaStream_t s) {
    if (P.n <= 0) return cudaSuccess;
    if (P.width < 1 || P.width > kMaxWidth) return cudaErrorInvalidValue;
    const unsigned blocks = (unsigned)((P.n + 127) / 128);
    ac_generic_kernel<<<blocks, 128, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_validate(const int8_t* in, uint8_t* flags, int64_t n, int mrl, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    ac_validate_kernel<<<blocks, 128, 0, s>>>(in, flags, n, mrl);
    return cudaGetLastError();
}

}  // namespace acs
