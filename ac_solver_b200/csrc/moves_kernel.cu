// moves_kernel.cu -- K1: batched ACMove / ACEnv.step over [N, 2*mrl] int8 rows.
//
// Reference semantics: ac_solver/envs/ac_moves.py:159-231 (ACMove) and
// ac_solver/envs/ac_env.py:95-113 (ACEnv.step); paths relative to /root/reference.
//
// Design (HBM-bound integer rewriting, no tensor cores):
//   * a tile of 128 rows (128 * 2*mrl bytes, 9216 B at mrl 36) is moved HBM -> shared
//     memory by ONE TMA bulk copy (cp.async.bulk, mbarrier completion) and back by ONE
//     bulk store, so global traffic is fully coalesced 16-byte bursts although a row is
//     72 bytes (8-byte aligned only);
//   * one thread owns one row: 64-bit conflict-free LDS of the row, int8 -> 2-bit codes
//     with a multiply-gather, the move on registers (ac_core.cuh), 2-bit -> int8 with a
//     PRMT table lookup, and only the REWRITTEN relator is stored back to the tile;
//   * reward / done / truncated / step counter are fused into the same pass.
// Budget at the measured 6.46 TB/s and 150 B per move: 26 warp-instructions per row per
// SM; this thread-per-row, loop-free formulation needs about 12.
#include <cstdint>
#include <cuda_runtime.h>

#include "ac_core.cuh"
#include "acs_internal.h"

namespace acs {

constexpr int kTileRows = 128;  // rows per CTA tile == threads per CTA

// ---- mbarrier / TMA bulk helpers (sm_90+ PTX, SASS: UBLKCP / SYNCS) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- int8 words <-> 2-bit codes -----------------------------------------------------
// four letters (one 32-bit word) -> their four codes in the TOP byte of the result
__device__ __forceinline__ uint32_t codes_top8(uint32_t w) {
    uint32_t t = (w & 0x01010101u) | ((w >> 6) & 0x02020202u);
    return t * 0x01041040u;  // 2^24 + 2^18 + 2^12 + 2^6: gathers the 2-bit fields, no carries
}
// top bytes of four products -> one 32-bit word of 16 codes
__device__ __forceinline__ uint32_t gather4(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
    uint32_t a = __byte_perm(p0, p1, 0x0073);
    uint32_t b = __byte_perm(p2, p3, 0x0073);
    return __byte_perm(a, b, 0x5410);
}

template <int NW, int W>
__device__ __forceinline__ Rel<W> pack_words(const uint32_t (&w)[NW]) {
    Rel<W> r;
    uint32_t nz = 0;
    uint32_t piece[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < (NW + 3) / 4; ++q) {
        uint32_t p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = 4 * q + k;
            if (j < NW) {
                p[k] = codes_top8(w[j]);
                nz += (w[j] | (w[j] >> 1)) & 0x01010101u;
            } else {
                p[k] = 0;
            }
        }
        piece[q] = gather4(p[0], p[1], p[2], p[3]);
    }
    r.b.w[0] = (uint64_t)piece[0] | ((uint64_t)piece[1] << 32);
    if constexpr (W == 2) r.b.w[1] = (uint64_t)piece[2] | ((uint64_t)piece[3] << 32);
    r.len = (int)((nz * 0x01010101u) >> 24);
    return r;
}

// word j (letters 4j..4j+3) of a packed relator back to int8, zero beyond len
template <int W>
__device__ __forceinline__ uint32_t unpack_word(const Rel<W>& r, int j) {
    uint32_t c8;
    if (W == 1 || j < 8) c8 = (uint32_t)(r.b.w[0] >> (8 * j)) & 0xFFu;
    else c8 = (uint32_t)(r.b.w[W - 1] >> (8 * (j - 8))) & 0xFFu;
    uint32_t t = (c8 | (c8 << 4)) & 0x0F0Fu;
    t = (t | (t << 2)) & 0x3333u;                       // one code per selector nibble
    uint32_t bytes = __byte_perm(0xFFFE0102u, 0u, t);   // code -> letter table lookup
    int n = max(8 * r.len - 32 * j, 0);
    return bytes & __funnelshift_lc(0xFFFFFFFFu, 0u, n);  // low min(n,32) bits kept
}

// ---- per-row epilogue shared by both kernels ---------------------------------------
struct RowOut {
    int len0, len1, status;
};

__device__ __forceinline__ void write_row_outputs(const StepParams& P, int64_t row, const RowOut& o) {
    if (o.status != ST_OK) {
        if (P.status) P.status[row] = (uint8_t)o.status;
        if (P.err) {
            atomicAdd((unsigned long long*)&P.err[0], 1ull);
            atomicMin((unsigned long long*)&P.err[1], (unsigned long long)row);
        }
        return;  // the reference raised: state, counters and rewards stay untouched
    }
    if (P.status) P.status[row] = 0;
    if (P.lens) reinterpret_cast<uchar2*>(P.lens)[row] = make_uchar2((uint8_t)o.len0, (uint8_t)o.len1);
    if (P.reward) {  // ac_env.py:101-105
        const int tot = o.len0 + o.len1;
        const bool d = tot == 2;
        P.reward[row] = d ? P.max_reward : -tot;
        P.done[row] = (uint8_t)d;
        int sc = P.step_count[row] + 1;
        P.step_count[row] = sc;
        P.truncated[row] = (uint8_t)(sc >= P.horizon);
    }
}

// ---- tile movement -------------------------------------------------------------------
__device__ __forceinline__ void tile_load(uint8_t* tile, uint64_t* bar, const StepParams& P, int64_t row0,
                                          int nrows, bool bulk) {
    const uint32_t bytes = (uint32_t)nrows * 2u * (uint32_t)P.mrl;
    const int8_t* src = P.in + row0 * 2 * P.mrl;
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(tile, src, bytes, bar);
        }
    } else {
        for (uint32_t b = threadIdx.x; b < bytes; b += blockDim.x) tile[b] = (uint8_t)src[b];
    }
}
__device__ __forceinline__ void tile_store(const uint8_t* tile, const StepParams& P, int64_t row0, int nrows,
                                           bool bulk) {
    const uint32_t bytes = (uint32_t)nrows * 2u * (uint32_t)P.mrl;
    int8_t* dst = P.out + row0 * 2 * P.mrl;
    if (bulk) {
        fence_async_proxy();  // make the generic-proxy STS visible to the async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store_1d(dst, tile, bytes);
            tma_store_commit_and_wait_read();
        }
    } else {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < bytes; b += blockDim.x) dst[b] = (int8_t)tile[b];
    }
}

// ---- fast kernel: mrl == 4*NW, rows read/written as whole words ----------------------
template <int NW, int W>
__global__ void __launch_bounds__(kTileRows, 8) ac_step_words_kernel(const StepParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    constexpr int ROWB = 8 * NW;  // bytes per row
    const int64_t row0 = (int64_t)blockIdx.x * kTileRows;
    const int nrows = (int)min((int64_t)kTileRows, P.n - row0);
    const bool bulk = P.bulk_ok && nrows == kTileRows;

    if (bulk && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    tile_load(smem, &bar, P, row0, nrows, bulk);

    const int64_t row = row0 + threadIdx.x;
    const bool active = threadIdx.x < nrows;
    int action = 0;
    if (active) action = P.action[row];  // overlaps the bulk copy

    if (bulk) mbar_wait(&bar, 0);
    else __syncthreads();

    if (active) {
        uint32_t w0[NW], w1[NW];
        const uint2* rp = reinterpret_cast<const uint2*>(smem + (size_t)threadIdx.x * ROWB);
        // 64-bit loads; relator 1 starts at word NW (a 64-bit boundary only when NW is even)
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            uint2 v = rp[j];
            const int a = 2 * j, b = 2 * j + 1;
            if (a < NW) w0[a] = v.x; else w1[a - NW] = v.x;
            if (b < NW) w0[b] = v.y; else w1[b - NW] = v.y;
        }
        Rel<W> r0 = pack_words<NW, W>(w0);
        Rel<W> r1 = pack_words<NW, W>(w1);
        RowOut o;
        if (action > 11) {
            o.status = ST_ASSERT;  // ac_moves.py:188-190
            o.len0 = o.len1 = 0;
        } else {
            bool changed_other = false;
            o.status = apply_move<W>(r0, r1, action, 4 * NW, P.cyclical != 0, changed_other);
            o.len0 = r0.len;
            o.len1 = r1.len;
            if (o.status == ST_OK) {
                const bool tgt1 = ((action + 1) & 1) != 0;
                uint32_t* wp = reinterpret_cast<uint32_t*>(smem + (size_t)threadIdx.x * ROWB) + (tgt1 ? NW : 0);
                const Rel<W>& t = tgt1 ? r1 : r0;
#pragma unroll
                for (int j = 0; j < NW; ++j) wp[j] = unpack_word<W>(t, j);
                if (changed_other) {  // only for caller-supplied, not yet normalised rows
                    uint32_t* op = reinterpret_cast<uint32_t*>(smem + (size_t)threadIdx.x * ROWB) + (tgt1 ? 0 : NW);
                    const Rel<W>& q = tgt1 ? r0 : r1;
#pragma unroll
                    for (int j = 0; j < NW; ++j) op[j] = unpack_word<W>(q, j);
                }
            }
        }
        write_row_outputs(P, row, o);
    }
    tile_store(smem, P, row0, nrows, bulk);
}

// ---- generic kernel: any mrl <= 64, byte accesses to the staged tile -----------------
template <int W>
__global__ void __launch_bounds__(kTileRows, 8) ac_step_bytes_kernel(const StepParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const int rowb = 2 * P.mrl;
    const int64_t row0 = (int64_t)blockIdx.x * kTileRows;
    const int nrows = (int)min((int64_t)kTileRows, P.n - row0);
    const bool bulk = P.bulk_ok && nrows == kTileRows;
    if (bulk && threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    tile_load(smem, &bar, P, row0, nrows, bulk);
    const int64_t row = row0 + threadIdx.x;
    const bool active = threadIdx.x < nrows;
    int action = 0;
    if (active) action = P.action[row];
    if (bulk) mbar_wait(&bar, 0);
    else __syncthreads();
    if (active) {
        int8_t* rp = reinterpret_cast<int8_t*>(smem) + (size_t)threadIdx.x * rowb;
        Rel<W> r0 = pack_bytes<W>(rp, P.mrl);
        Rel<W> r1 = pack_bytes<W>(rp + P.mrl, P.mrl);
        RowOut o;
        if (action > 11) {
            o.status = ST_ASSERT;
            o.len0 = o.len1 = 0;
        } else {
            bool changed_other = false;
            o.status = apply_move<W>(r0, r1, action, P.mrl, P.cyclical != 0, changed_other);
            o.len0 = r0.len;
            o.len1 = r1.len;
            if (o.status == ST_OK) {
                const bool tgt1 = ((action + 1) & 1) != 0;
                if (tgt1 || changed_other) unpack_bytes<W>(rp + P.mrl, r1, P.mrl);
                if (!tgt1 || changed_other) unpack_bytes<W>(rp, r0, P.mrl);
            }
        }
        write_row_outputs(P, row, o);
    }
    tile_store(smem, P, row0, nrows, bulk);
}

template <int NW>
static cudaError_t launch_words(const StepParams& P, cudaStream_t s) {
    constexpr int W = NW <= 8 ? 1 : 2;
    const int64_t tiles = (P.n + kTileRows - 1) / kTileRows;
    const size_t smem = (size_t)kTileRows * 8 * NW;
    ac_step_words_kernel<NW, W><<<(unsigned)tiles, kTileRows, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_step(const StepParams& P, cudaStream_t s) {
    if (P.n <= 0) return cudaSuccess;
    if (P.mrl < 1 || P.mrl > 64) return cudaErrorInvalidValue;
    if (P.mrl % 4 == 0) {
        switch (P.mrl / 4) {
#define ACS_CASE(k) case k: return launch_words<k>(P, s);
            ACS_CASE(1) ACS_CASE(2) ACS_CASE(3) ACS_CASE(4) ACS_CASE(5) ACS_CASE(6) ACS_CASE(7) ACS_CASE(8)
            ACS_CASE(9) ACS_CASE(10) ACS_CASE(11) ACS_CASE(12) ACS_CASE(13) ACS_CASE(14) ACS_CASE(15)
            ACS_CASE(16)
#undef ACS_CASE
        }
    }
    const int64_t tiles = (P.n + kTileRows - 1) / kTileRows;
    const size_t smem = (size_t)kTileRows * 2 * P.mrl;
    if (P.mrl <= 32) ac_step_bytes_kernel<1><<<(unsigned)tiles, kTileRows, smem, s>>>(P);
    else ac_step_bytes_kernel<2><<<(unsigned)tiles, kTileRows, smem, s>>>(P);
    return cudaGetLastError();
}

}  // namespace acs
