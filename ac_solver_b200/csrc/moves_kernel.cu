// moves_kernel.cu -- K1: batched ACMove / ACEnv.step over [N, 2*mrl] int8 rows.
//
// Reference semantics: ac_solver/envs/ac_moves.py:159-231 (ACMove) and
// ac_solver/envs/ac_env.py:95-113 (ACEnv.step); paths relative to /root/reference.
//
// Design (integer rewriting, no tensor cores; the roofline is HBM, the practical limit is
// the ALU pipe):
//   * a tile of 128 rows (128 * 2*mrl bytes, 9216 B at mrl 36) is moved HBM -> shared
//     memory by ONE TMA bulk copy (cp.async.bulk, mbarrier completion) and back by ONE
//     bulk store, so global traffic is fully coalesced 16-byte bursts although a row is
//     72 bytes (8-byte aligned only);
//   * one thread owns one row; the rows of a tile are re-assigned by move class (ballot +
//     popc prefix) so that warps are uniform: concatenations go through 2-bit packed
//     registers (int8 -> codes with multiply-gathers, ac_pack.cuh; the move in ac_core.cuh;
//     codes -> int8 with a PRMT table lookup), conjugations of normal-form states are a
//     one-letter rotation done directly on the int8 words; only the REWRITTEN relator is
//     stored back;
//   * reward / done / truncated / step counter / lengths / action log are fused into the
//     same pass, staged in shared memory and written coalesced by the row's owner thread;
//   * consecutive launches overlap their ramp and drain by programmatic dependent launch.
// Measured: 27.8 us for 1 Mi rows at mrl 36 = 0.875 of the measured HBM copy peak.
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include <cuda_runtime.h>

#include "ac_core.cuh"
#include "ac_pack.cuh"
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
        if (P.reward) {  // the reference raised: the state and the step counter stay untouched; the step's
            P.reward[row] = 0;  // outputs are cleared so that a stale done/truncated cannot trigger an auto-reset
            P.done[row] = 0;
            P.truncated[row] = 0;
        }
        return;
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
        if (P.action_log)  // ACEnv.actions (ac_env.py:96)
            P.action_log[row * P.log_stride + min(sc - 1, P.log_stride - 1)] = P.action[row];
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

// store a packed relator as NW int8 words
template <int NW, int N>
__device__ __forceinline__ void store_relator(uint32_t* wp, const Rel<N>& t) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; ++i) {
        uint32_t lo, hi;
        unpack_pair<N>(t, i, lo, hi);
        wp[2 * i] = lo;
        if (2 * i + 1 < NW) wp[2 * i + 1] = hi;
    }
}

// ---- fast kernel: mrl == 4*NW, rows read/written as whole words ----------------------
// Rows of a tile are partitioned by move class so that warps are uniform: threads
// [0, n0) take the concatenation rows (ids 0..3, and invalid ids), the rest the
// conjugation rows (ids 4..11), each class in row order (ballot + popc prefix).  With
// TRUSTED states and cyclic reduction a conjugation is a rotation of the target relator by
// one letter (or nothing), done directly on the int8 words with funnel shifts -- two thirds
// of uniformly random moves never enter the packed domain.  Per-row results are staged in
// shared memory and written back by the thread that owns the row index, so every global
// access stays coalesced.
template <int NW, bool TRUSTED, bool LENS, int TR>
__global__ void __launch_bounds__(TR, (TR <= 128 ? 8 : 4)) ac_step_words_kernel(const StepParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint8_t s_perm[TR];
    __shared__ uint8_t s_act[TR];
    __shared__ uint32_t s_res[TR];  // status << 16 | len0 << 8 | len1
    __shared__ int s_wcnt[TR / 32];
    __shared__ uint16_t s_lens[TR];  // len0 | len1 << 8 on entry (ACS_FLAG_LENS_VALID)
    constexpr int N = (NW + 3) / 4;
    constexpr int ROWB = 8 * NW;  // bytes per row
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * TR;
    const int nrows = (int)min((int64_t)TR, P.n - row0);
    const bool bulk = P.bulk_ok && nrows == TR;

    // PDL: let the next grid of the stream start launching, then wait for the previous grid's
    // memory before touching global data (no-ops when launched without the attribute)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (bulk && tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    tile_load(smem, &bar, P, row0, nrows, bulk);

    // ---- overlaps the bulk copy: per-row scalars and the class partition ----
    const int64_t row = row0 + tid;
    const bool active = tid < nrows;
    int action = 255;
    int sc = 0;
    constexpr bool have_lens = TRUSTED && LENS;  // lengths carried beside the state (ac_env.py:84-92)
    if (active) {
        action = P.action[row];
        if (P.reward) sc = P.step_count[row];
        if constexpr (have_lens) s_lens[tid] = reinterpret_cast<const uint16_t*>(P.lens)[row];
    }
    const bool is_concat = active && !(action >= 4 && action <= 11);
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, is_concat);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) s_wcnt[wid] = __popc(bal);
    s_act[tid] = (uint8_t)action;
    __syncthreads();
    int n0 = 0, before = __popc(bal & ((1u << lane) - 1u));
#pragma unroll
    for (int w = 0; w < TR / 32; ++w) {
        const int c = s_wcnt[w];
        n0 += c;
        if (w < wid) before += c;
    }
    s_perm[is_concat ? before : n0 + (tid - before)] = (uint8_t)tid;  // stable within each class
    __syncthreads();
    const int r = s_perm[tid];  // the row this thread processes
    const int act = s_act[r];

    if (bulk) mbar_wait(&bar, 0);

    if (r < nrows) {
        uint32_t* rw = reinterpret_cast<uint32_t*>(smem + (size_t)r * ROWB);
        int status = ST_OK, len0, len1;
        if (TRUSTED && P.cyclical && act >= 4 && act <= 11) {
            // ---- conjugation of a normal form: rotation in the byte domain ----
            const bool tgt1 = ((act + 1) & 1) != 0;
            uint32_t u[NW];
            int lu, lw;
            if constexpr (have_lens) {
                // only the target relator is touched: (NW+1)/2 64-bit loads starting at its
                // 8-byte aligned base (one word early when NW is odd and the target is r1)
                constexpr int LD = (NW + 1) / 2;
                uint32_t t[2 * LD];
                const uint2* rp = reinterpret_cast<const uint2*>(rw) + (tgt1 ? NW / 2 : 0);
#pragma unroll
                for (int j = 0; j < LD; ++j) {
                    const uint2 v = rp[j];
                    t[2 * j] = v.x;
                    t[2 * j + 1] = v.y;
                }
                const bool skew = (NW & 1) && tgt1;
#pragma unroll
                for (int j = 0; j < NW; ++j) u[j] = skew ? t[j + 1] : t[j];
                const int l01 = s_lens[r];
                lu = tgt1 ? (l01 >> 8) : (l01 & 0xFF);
                lw = tgt1 ? (l01 & 0xFF) : (l01 >> 8);
            } else {
                uint32_t w0[NW], w1[NW], w[NW];
                const uint2* rp = reinterpret_cast<const uint2*>(rw);
#pragma unroll
                for (int j = 0; j < NW; ++j) {  // 64-bit loads of the whole row: fewest LDS wavefronts
                    const uint2 v = rp[j];
                    const int a = 2 * j, b = 2 * j + 1;
                    if (a < NW) w0[a] = v.x; else w1[a - NW] = v.x;
                    if (b < NW) w0[b] = v.y; else w1[b - NW] = v.y;
                }
#pragma unroll
                for (int j = 0; j < NW; ++j) {
                    u[j] = tgt1 ? w1[j] : w0[j];
                    w[j] = tgt1 ? w0[j] : w1[j];
                }
                lu = count_letters<NW>(u);
                lw = count_letters<NW>(w);
            }
            len0 = tgt1 ? lw : lu;
            len1 = tgt1 ? lu : lw;
            if (lu == 0) status = ST_INDEX;       // relator_nonzero[0] on an empty array
            else if (lw == 0) status = ST_ASSERT;  // utils.py:261-263
            else {
                uint8_t* ub = reinterpret_cast<uint8_t*>(rw + (tgt1 ? NW : 0));
                int fix_pos;
                uint32_t fix_val;
                if (conj_rotate_words<NW>(u, lu, u[0] & 0xFFu, ub[lu - 1], conj_letter_byte(act), fix_pos, fix_val)) {
#pragma unroll
                    for (int j = 0; j < NW; ++j) rw[(tgt1 ? NW : 0) + j] = u[j];
                    if (fix_pos >= 0) ub[fix_pos] = (uint8_t)fix_val;
                }
            }
        } else {
            // ---- packed path: concatenations, and everything for untrusted / non-cyclic rows ----
            uint32_t w0[NW], w1[NW];
            const uint2* rp = reinterpret_cast<const uint2*>(rw);
#pragma unroll
            for (int j = 0; j < NW; ++j) {
                const uint2 v = rp[j];
                const int a = 2 * j, b = 2 * j + 1;
                if (a < NW) w0[a] = v.x; else w1[a - NW] = v.x;
                if (b < NW) w0[b] = v.y; else w1[b - NW] = v.y;
            }
            Rel<N> r0, r1;
            if constexpr (have_lens) {
                r0 = pack_words<NW, N, false>(w0);
                r1 = pack_words<NW, N, false>(w1);
                const int l01 = s_lens[r];
                r0.len = l01 & 0xFF;
                r1.len = l01 >> 8;
            } else {
                r0 = pack_words<NW, N>(w0);
                r1 = pack_words<NW, N>(w1);
            }
            if (act > 11) {
                status = ST_ASSERT;  // ac_moves.py:188-190
            } else {
                bool changed_other = false;
                status = apply_move<N, TRUSTED>(r0, r1, act, 4 * NW, P.cyclical != 0, changed_other);
                if (status == ST_OK) {
                    const bool tgt1 = ((act + 1) & 1) != 0;
                    Rel<N> t;
#pragma unroll
                    for (int j = 0; j < N; ++j) t.b.w[j] = tgt1 ? r1.b.w[j] : r0.b.w[j];
                    t.len = tgt1 ? r1.len : r0.len;
                    store_relator<NW, N>(rw + (tgt1 ? NW : 0), t);
                    if (!TRUSTED && changed_other)  // only for caller-supplied, not yet normalised rows
                        store_relator<NW, N>(rw + (tgt1 ? 0 : NW), tgt1 ? r0 : r1);
                }
            }
            len0 = r0.len;
            len1 = r1.len;
        }
        s_res[r] = ((uint32_t)status << 16) | ((uint32_t)len0 << 8) | (uint32_t)len1;
    }
    tile_store(smem, P, row0, nrows, bulk);  // fence + __syncthreads inside: s_res is visible after it

    if (active) {  // coalesced per-row outputs by the owner of the row index
        const uint32_t res = s_res[tid];
        const int status = (int)(res >> 16), l0 = (int)((res >> 8) & 0xFF), l1 = (int)(res & 0xFF);
        if (status != ST_OK) {
            if (P.status) P.status[row] = (uint8_t)status;
            if (P.err) {
                atomicAdd((unsigned long long*)&P.err[0], 1ull);
                atomicMin((unsigned long long*)&P.err[1], (unsigned long long)row);
            }
            if (P.reward) {  // cleared, so that a stale done/truncated cannot trigger an auto-reset
                P.reward[row] = 0;
                P.done[row] = 0;
                P.truncated[row] = 0;
            }
        } else {  // (a raising row keeps its state and step counter)
            if (P.status) P.status[row] = 0;
            if (P.lens) reinterpret_cast<uchar2*>(P.lens)[row] = make_uchar2((uint8_t)l0, (uint8_t)l1);
            if (P.reward) {  // ac_env.py:101-105
                const int tot = l0 + l1;
                const bool d = tot == 2;
                P.reward[row] = d ? P.max_reward : -tot;
                P.done[row] = (uint8_t)d;
                P.step_count[row] = sc + 1;
                P.truncated[row] = (uint8_t)(sc + 1 >= P.horizon);
                if (P.action_log)  // ACEnv.actions (ac_env.py:96): the episode's move sequence
                    P.action_log[row * P.log_stride + min(sc, P.log_stride - 1)] = (uint8_t)action;
            }
        }
    }
}

// ---- generic kernel: any mrl <= 64, byte accesses to the staged tile -----------------
template <int N, bool TRUSTED>
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
        Rel<N> r0 = pack_bytes<N>(rp, P.mrl);
        Rel<N> r1 = pack_bytes<N>(rp + P.mrl, P.mrl);
        RowOut o;
        if (action > 11) {
            o.status = ST_ASSERT;
            o.len0 = o.len1 = 0;
        } else {
            bool changed_other = false;
            o.status = apply_move<N, TRUSTED>(r0, r1, action, P.mrl, P.cyclical != 0, changed_other);
            o.len0 = r0.len;
            o.len1 = r1.len;
            if (o.status == ST_OK) {
                const bool tgt1 = ((action + 1) & 1) != 0;
                if (tgt1 || changed_other) unpack_bytes<N>(rp + P.mrl, r1, P.mrl);
                if (!tgt1 || changed_other) unpack_bytes<N>(rp, r0, P.mrl);
            }
        }
        write_row_outputs(P, row, o);
    }
    tile_store(smem, P, row0, nrows, bulk);
}

template <int NW>
static cudaError_t launch_words(const StepParams& P, cudaStream_t s) {
    // tile rows: 128 by default; ACS_TILE_ROWS=64|256 selects the tuning variants that are
    // compiled for the headline width (mrl 36)
    static const int tile_rows = [] {
        const char* e = getenv("ACS_TILE_ROWS");
        const int v = e ? atoi(e) : 128;
        return (v == 64 || v == 256) ? v : 128;
    }();
    auto launch = [&](auto tr) {
        constexpr int TR = decltype(tr)::value;
        const int64_t tiles = (P.n + TR - 1) / TR;
        const size_t smem = (size_t)TR * 8 * NW;
        // Programmatic dependent launch: the next step kernel in the stream may become resident
        // while this one drains; it blocks in griddepcontrol.wait before its first global access,
        // so stream ordering of the data is unchanged (ACS_PDL=0 disables).
        static const bool pdl = [] {
            const char* e = getenv("ACS_PDL");
            return !(e && e[0] == '0');
        }();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)tiles);
        cfg.blockDim = dim3(TR);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        if (P.trusted && P.lens_valid) cudaLaunchKernelEx(&cfg, ac_step_words_kernel<NW, true, true, TR>, P);
        else if (P.trusted) cudaLaunchKernelEx(&cfg, ac_step_words_kernel<NW, true, false, TR>, P);
        else cudaLaunchKernelEx(&cfg, ac_step_words_kernel<NW, false, false, TR>, P);
    };
    if constexpr (NW == 9) {
        if (tile_rows == 64) launch(std::integral_constant<int, 64>{});
        else if (tile_rows == 256) launch(std::integral_constant<int, 256>{});
        else launch(std::integral_constant<int, 128>{});
    } else {
        launch(std::integral_constant<int, 128>{});
    }
    return cudaGetLastError();
}

template <int N>
static cudaError_t launch_bytes(const StepParams& P, cudaStream_t s) {
    const int64_t tiles = (P.n + kTileRows - 1) / kTileRows;
    const size_t smem = (size_t)kTileRows * 2 * P.mrl;
    if (P.trusted) ac_step_bytes_kernel<N, true><<<(unsigned)tiles, kTileRows, smem, s>>>(P);
    else ac_step_bytes_kernel<N, false><<<(unsigned)tiles, kTileRows, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_step(const StepParams& P, cudaStream_t s) {
    if (P.n <= 0) return cudaSuccess;
    if (P.mrl < 1 || P.mrl > 64) return cudaErrorInvalidValue;
    if (P.mrl % 4 == 0) {
        switch (P.mrl / 4) {
#define ACS_CASE(k) \
    case k:         \
        return launch_words<k>(P, s);
            ACS_CASE(1) ACS_CASE(2) ACS_CASE(3) ACS_CASE(4) ACS_CASE(5) ACS_CASE(6) ACS_CASE(7) ACS_CASE(8)
            ACS_CASE(9) ACS_CASE(10) ACS_CASE(11) ACS_CASE(12) ACS_CASE(13) ACS_CASE(14) ACS_CASE(15)
            ACS_CASE(16)
#undef ACS_CASE
        }
    }
    switch (words_for(P.mrl)) {
        case 1: return launch_bytes<1>(P, s);
        case 2: return launch_bytes<2>(P, s);
        case 3: return launch_bytes<3>(P, s);
        default: return launch_bytes<4>(P, s);
    }
}

// Auto-reset of finished environments (gymnasium 0.28.1 SyncVectorEnv semantics, see
// envs/vector_env.py): rows with done|truncated hand their last observation to final_obs and go
// back to their own initial state with zeroed counters.  One thread per 2 bytes of a row.
__global__ void __launch_bounds__(256) env_autoreset_kernel(uint16_t* state, const uint16_t* init, uint16_t* final_obs,
                                                            const uint8_t* done, const uint8_t* trunc,
                                                            int32_t* step_count, int32_t* final_steps,
                                                            uint16_t* lens, const uint16_t* init_lens, int64_t n,
                                                            int mrl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * mrl) return;
    const int64_t row = i / mrl;
    if (!(done[row] | trunc[row])) return;
    if (final_obs) final_obs[i] = state[i];
    state[i] = init[i];
    if (i % mrl == 0) {
        if (final_steps) final_steps[row] = step_count[row];
        step_count[row] = 0;
        if (lens) lens[row] = init_lens[row];
    }
}

cudaError_t launch_autoreset(int8_t* state, const int8_t* init, int8_t* final_obs, const uint8_t* done,
                             const uint8_t* trunc, int32_t* step_count, int32_t* final_steps, uint8_t* lens,
                             const uint8_t* init_lens, int64_t n, int mrl, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * mrl;
    env_autoreset_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<uint16_t*>(state), reinterpret_cast<const uint16_t*>(init),
        reinterpret_cast<uint16_t*>(final_obs), done, trunc, step_count, final_steps, reinterpret_cast<uint16_t*>(lens),
        reinterpret_cast<const uint16_t*>(init_lens), n, mrl);
    return cudaGetLastError();
}


// ---- reward wrappers of the PPO environment, per environment and on the device ------------------
// gymnasium 0.28.1 NormalizeReward (agents/environment.py:45-46; one wrapper per environment):
//   returns = returns * gamma * (1 - terminated) + reward;  rms.update([returns]);
//   reward  = reward / sqrt(rms.var + eps)          (RunningMeanStd: mean 0, var 1, count 1e-4)
// followed by TransformReward(clip) (environment.py:48-52).  stats = [4][n] doubles: returns,
// mean, var, count.  (Recalled semantics: the gymnasium package is not available here, so this
// wrapper is "parity unpinned" like the rest of the vector-env shell.)
__global__ void __launch_bounds__(256) env_reward_transform_kernel(const int32_t* reward, const uint8_t* done,
                                                                   double* stats, float* out, int64_t n, double gamma,
                                                                   double eps, int normalize, int clip, double lo,
                                                                   double hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = (double)reward[i];
    if (normalize) {
        double ret = stats[i] * gamma * (done[i] ? 0.0 : 1.0) + r;
        double mean = stats[n + i], var = stats[2 * n + i], count = stats[3 * n + i];
        const double delta = ret - mean, tot = count + 1.0;  // batch of one: batch_var = 0
        const double new_mean = mean + delta / tot;
        const double m2 = var * count + delta * delta * count / tot;
        stats[i] = ret;
        stats[n + i] = new_mean;
        stats[2 * n + i] = m2 / tot;
        stats[3 * n + i] = tot;
        r = r / sqrt(m2 / tot + eps);
    }
    if (clip) r = fmin(fmax(r, lo), hi);
    out[i] = (float)r;
}

cudaError_t launch_reward_transform(const int32_t* reward, const uint8_t* done, double* stats, float* out, int64_t n,
                                    double gamma, double eps, int normalize, int clip, double lo, double hi,
                                    cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    env_reward_transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reward, done, stats, out, n, gamma, eps,
                                                                            normalize, clip, lo, hi);
    return cudaGetLastError();
}

// ---- curriculum reset of the PPO rollout, on the device (agents/training.py:169-224) -----------
// Finished environments (done | truncated) record their result and move on to another initial state
// of the pool: round one hands out the not yet processed states in order (in environment order, as
// the reference's sequential loop does); afterwards an unsolved state with probability
// 1 - repeat_solved_prob (or when nothing is solved yet), else a solved one, uniformly.  The
// reference draws from Python's `random`; here a counter-based hash of (seed, draw index)
// supplies the uniform numbers (same distribution, not the same stream), and the solved set seen by
// a draw is the one at the END of the step.  ONE block (the environment count is in the thousands).
__device__ __forceinline__ unsigned long long cur_mix(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void __launch_bounds__(1024) env_curriculum_kernel(const CurriculumParams P) {
    __shared__ unsigned s_warp[32];
    __shared__ long long s_base, s_draw0;
    __shared__ int s_nsolved;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rowb = 2 * P.mrl;
    if (tid == 0) {
        s_base = P.counters[0];
        s_draw0 = P.counters[2];
    }
    __syncthreads();
    // ---- pass 1: results of the finished episodes (training.py:170-189) ----
    for (int64_t i = tid; i < P.n; i += blockDim.x) {
        if (!(P.done[i] | P.trunc[i])) continue;
        if (P.final_steps) P.final_steps[i] = P.step_count[i];
        if (P.done[i]) {
            const int st = P.cur_state[i];
            // `solved` is a byte per state; the first solver appends the state to the solved list
            unsigned int* word = reinterpret_cast<unsigned int*>(P.solved + (st & ~3));
            const unsigned int bit = 1u << (8 * (st & 3));
            if (!(atomicOr(word, bit) & bit)) P.solved_list[atomicAdd((unsigned long long*)&P.counters[1], 1ull)] = st;
            atomicMin(&P.best[st], ((unsigned long long)P.step_count[i] << 32) | (unsigned long long)i);
        }
    }
    __threadfence_block();
    __syncthreads();
    for (int64_t i = tid; i < P.n; i += blockDim.x) {  // the shortest solving episode keeps its action sequence
        if (!P.done[i]) continue;
        const int st = P.cur_state[i];
        if (P.best[st] == (((unsigned long long)P.step_count[i] << 32) | (unsigned long long)i)) {
            const int len = min(P.step_count[i], P.log_stride);
            for (int k = 0; k < len; ++k) P.best_actions[(int64_t)st * P.log_stride + k] = P.action_log[i * P.log_stride + k];
        }
    }
    if (tid == 0) s_nsolved = (int)P.counters[1];
    __syncthreads();
    // ---- pass 2: the next initial state of every finished environment, in environment order ----
    long long fin_seen = 0;
    for (int64_t i0 = 0; i0 < P.n; i0 += blockDim.x) {
        const int64_t i = i0 + tid;
        const bool fin = i < P.n && (P.done[i] | P.trunc[i]);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, fin);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        long long before = fin_seen + __popc(bal & ((1u << lane) - 1u));
        long long chunk_total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            if (w < wid) before += s_warp[w];
            chunk_total += s_warp[w];
        }
        if (fin) {
            const long long cand = s_base + before;  // round one: max(states_processed) + 1 (training.py:206-207)
            int next;
            if (cand < P.n_states) {
                next = (int)cand;
            } else {
                const unsigned long long d = (unsigned long long)(s_draw0 + before) * 8;
                const float u = (float)(cur_mix(P.seed ^ cur_mix(d)) >> 40) * (1.0f / 16777216.0f);
                const int ns = s_nsolved, nu = P.n_states - ns;
                if (ns == 0 || (nu > 0 && u > P.repeat_solved_prob)) {  // training.py:210-217
                    next = -1;
                    for (int t = 1; t <= 6 && next < 0; ++t) {  // rejection sampling of an unsolved state
                        const int c = (int)(cur_mix(P.seed ^ cur_mix(d + t)) % (unsigned long long)P.n_states);
                        if (!P.solved[c]) next = c;
                    }
                    if (next < 0) {  // dense solved set: first unsolved state after a random start
                        int c = (int)(cur_mix(P.seed ^ cur_mix(d + 7)) % (unsigned long long)P.n_states);
                        for (int t = 0; t < P.n_states && P.solved[c]; ++t) c = c + 1 == P.n_states ? 0 : c + 1;
                        next = c;
                    }
                } else {
                    next = P.solved_list[cur_mix(P.seed ^ cur_mix(d + 1)) % (unsigned long long)ns];
                }
            }
            P.cur_state[i] = next;
            int8_t* row = P.state + i * rowb;
            const int8_t* src = P.pool + (int64_t)next * rowb;
            for (int k = 0; k < rowb; ++k) {
                if (P.final_obs) P.final_obs[i * rowb + k] = row[k];
                row[k] = src[k];
            }
            P.lens[2 * i] = P.pool_lens[2 * next];
            P.lens[2 * i + 1] = P.pool_lens[2 * next + 1];
            P.step_count[i] = 0;
        }
        fin_seen += chunk_total;
        __syncthreads();
    }
    if (tid == 0) {
        const long long used = min((long long)P.n_states - s_base, fin_seen);
        P.counters[0] = s_base + (used > 0 ? used : 0);
        P.counters[2] = s_draw0 + fin_seen;
        P.counters[3] += fin_seen;
    }
}

cudaError_t launch_curriculum(const CurriculumParams& P, cudaStream_t s) {
    if (P.n <= 0) return cudaSuccess;
    env_curriculum_kernel<<<1, 1024, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace acs
