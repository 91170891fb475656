// bfs.cu -- breadth-first search of the AC graph on the device, bit-exact with the
// reference's sequential bfs() (ac_solver/search/breadth_first.py:15-97, paths relative to
// /root/reference): same (solved, path), same visited set in the same insertion order at any
// node budget, same "New minimal length" sequence.
//
// Formulation (generalised from levels to chunks).  The FIFO queue of
// the reference IS the node array in insertion order, so the search processes consecutive
// chunks [head, head+F) of already committed nodes:
//   expand : one thread per candidate (parent p, action a), global id g = 12*p + a.  The
//            child is computed on packed relators (ac_core.cuh) and inserted into an exact
//            open-addressing table with "smallest id wins": a slot is one 64-bit word
//            (24-bit fingerprint | 40-bit index+1).  Index < n_nodes denotes a committed
//            node, n_nodes + c a tentative candidate of this chunk, so one atomicMin both
//            lets committed nodes beat candidates and keeps the earliest candidate.  Keys
//            are never trusted to the fingerprint: on a fingerprint match the stored key is
//            read (committed) or recomputed from its parent (tentative) and compared.
//   mark   : a candidate is a winner iff its slot still holds its own id; 384-thread blocks
//            cover 32 parents x 12 actions, block winner counts feed a scan.
//   cut    : (last chunk only) first parent after which |visited| >= budget.
//   commit : winners with id below the limit (solving child / cut) are appended in id order
//            -- ballot/popc block scan + scanned block offsets -- which is exactly the
//            reference's FIFO order; their table slots are re-pointed at the new node index.
// Later chunks see earlier chunks' children as committed, so chunking is sequentially exact.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/acsolver_b200.h"
#include "ac_core.cuh"
#include "ac_keys.cuh"
#include "acs_internal.h"

namespace acs {

constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kParentsPerBlock = 32;
constexpr int kBfsThreads = kParentsPerBlock * 12;  // 384
constexpr int64_t kMaxChunkParents = 1ll << 22;

struct BfsCtrl {
    unsigned long long sol;        // min global candidate id with total length 2
    unsigned long long err;        // min (global candidate id << 2 | status)
    unsigned long long total;      // winners in the chunk
    unsigned long long committed;  // winners appended by the commit kernel
    unsigned long long cut;        // first chunk-local parent after which |visited| >= budget
    unsigned long long first_len[128];  // first global candidate id producing total length L
};

struct BfsArgs {
    uint64_t* keys;
    uint64_t* parent;
    uint64_t* table;
    uint64_t tmask;
    uint32_t* cand_slot;
    uint32_t* block_cnt;
    BfsCtrl* ctrl;
    uint64_t head;     // first parent of the chunk
    uint64_t nparents; // F
    uint64_t n_nodes;  // committed nodes at chunk start
    uint64_t budget;
    uint64_t limit;    // commit: chunk-local candidate limit
    int mrl;
    int cyclical;
    int min_len;       // smallest total length seen before this chunk
    int trusted;       // all parents of the chunk are normal forms (every chunk but the root's)
};

template <int W>
__device__ __forceinline__ int child_of(const BfsArgs& A, uint64_t parent, int action, Key<W>& child, Key<W>& pk,
                                        int& total_len) {
    pk = load_key<W>(A.keys, parent);
    Rel<2 * W> r0, r1;
    split_key<W>(pk, r0, r1);
    bool co;
    // every node after the root was produced by apply_move with this cyclical flag, so it is
    // a normal form; only the caller-supplied root needs the reference's full simplification
    const int st = A.trusted ? apply_move<2 * W, true>(r0, r1, action, A.mrl, A.cyclical != 0, co)
                             : apply_move<2 * W, false>(r0, r1, action, A.mrl, A.cyclical != 0, co);
    child = make_key<W>(r0, r1);
    total_len = r0.len + r1.len;
    return st;
}

template <int W>
__global__ void __launch_bounds__(kBfsThreads) bfs_expand_kernel(const BfsArgs A) {
    const uint64_t c = (uint64_t)blockIdx.x * kBfsThreads + threadIdx.x;
    if (c >= A.nparents * 12) return;
    const uint64_t p = A.head + c / 12;
    const int a = (int)(c % 12);
    const uint64_t gid = p * 12 + a;
    Key<W> child, pk;
    int L;
    const int st = child_of<W>(A, p, a, child, pk, L);
    uint32_t my_slot = kNoSlot;
    if (st != ST_OK) {
        atomicMin(&A.ctrl->err, (gid << 2) | (unsigned)st);
    } else {
        if (L < A.min_len) atomicMin(&A.ctrl->first_len[L], gid);
        if (L == 2) atomicMin(&A.ctrl->sol, gid);  // breadth_first.py:84-85, before the visited test
        if (!key_eq<W>(child, pk)) {
            const uint64_t h = key_hash<W>(child);
            const uint64_t fp = h >> 40;
            const uint64_t mine = (fp << 40) | (A.n_nodes + c + 1);
            uint64_t s = h & A.tmask;
            uint64_t probes = 0;
            for (;;) {
                uint64_t cur = __ldcg(&A.table[s]);  // L2 read: other SMs update slots atomically
                if (cur == 0) {
                    cur = atomicCAS((unsigned long long*)&A.table[s], 0ull, (unsigned long long)mine);
                    if (cur == 0) {
                        my_slot = (uint32_t)s;
                        break;
                    }
                }
                if ((cur >> 40) == fp) {
                    const uint64_t idx = (cur & kIdxMask) - 1;
                    Key<W> other;
                    if (idx < A.n_nodes) {
                        other = load_key<W>(A.keys, idx);
                    } else {
                        const uint64_t oc = idx - A.n_nodes;
                        Key<W> opk;
                        int oL;
                        child_of<W>(A, A.head + oc / 12, (int)(oc % 12), other, opk, oL);
                    }
                    if (key_eq<W>(other, child)) {
                        if (idx >= A.n_nodes) {  // same state generated twice in this chunk
                            atomicMin((unsigned long long*)&A.table[s], (unsigned long long)mine);
                            my_slot = (uint32_t)s;
                        }
                        break;  // committed: already visited
                    }
                }
                s = (s + 1) & A.tmask;
                if (++probes > A.tmask) {  // table full: cannot happen with the host's chunk sizing
                    atomicMin(&A.ctrl->err, (gid << 2) | 3u);
                    break;
                }
            }
        }
    }
    A.cand_slot[c] = my_slot;
}

// block-wide exclusive scan of 0/1 flags with ballots; returns this thread's rank and the total
__device__ __forceinline__ uint32_t block_rank(bool flag, uint32_t& total) {
    __shared__ uint32_t warp_cnt[kBfsThreads / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, flag);
    const uint32_t within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kBfsThreads / 32; ++w) {
        const uint32_t v = warp_cnt[w];
        if (w < (int)wid) before += v;
        tot += v;
    }
    __syncthreads();
    total = tot;
    return before + within;
}

__global__ void __launch_bounds__(kBfsThreads) bfs_mark_kernel(const BfsArgs A) {
    const uint64_t c = (uint64_t)blockIdx.x * kBfsThreads + threadIdx.x;
    bool win = false;
    if (c < A.nparents * 12) {
        const uint32_t s = A.cand_slot[c];
        if (s != kNoSlot) {
            win = ((A.table[s] & kIdxMask) - 1) == A.n_nodes + c;
            if (!win) A.cand_slot[c] = kNoSlot;
        }
    }
    uint32_t total;
    block_rank(win, total);
    if (threadIdx.x == 0) A.block_cnt[blockIdx.x] = total;
}

// exclusive scan of block_cnt[0..n) in place, total to ctrl->total (single block)
__global__ void __launch_bounds__(1024) bfs_scan_kernel(uint32_t* block_cnt, uint64_t n, BfsCtrl* ctrl) {
    __shared__ uint64_t part[1024];
    const uint64_t per = (n + 1023) / 1024;
    const uint64_t lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
    uint64_t s = 0;
    for (uint64_t i = lo; i < hi; ++i) s += block_cnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint64_t v = part[i];
            part[i] = run;
            run += v;
        }
        ctrl->total = run;
    }
    __syncthreads();
    uint64_t run = part[threadIdx.x];
    for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t v = block_cnt[i];
        block_cnt[i] = (uint32_t)run;  // chunk winners < 2^32 (<= 12 * 2^22)
        run += v;
    }
}

// last chunk only: first parent p (chunk-local) with n_nodes + #winners(parents <= p) >= budget
__global__ void __launch_bounds__(kBfsThreads) bfs_cut_kernel(const BfsArgs A) {
    const uint64_t c = (uint64_t)blockIdx.x * kBfsThreads + threadIdx.x;
    const bool win = c < A.nparents * 12 && A.cand_slot[c] != kNoSlot;
    uint32_t total;
    const uint32_t rank = block_rank(win, total);
    // the last candidate (a == 11) of each parent knows the inclusive count up to its parent
    const uint64_t after = A.n_nodes + A.block_cnt[blockIdx.x] + rank + (win ? 1 : 0);
    if (c < A.nparents * 12 && (c % 12) == 11 && after >= A.budget) atomicMin(&A.ctrl->cut, c / 12);
}

template <int W>
__global__ void __launch_bounds__(kBfsThreads) bfs_commit_kernel(const BfsArgs A) {
    const uint64_t c = (uint64_t)blockIdx.x * kBfsThreads + threadIdx.x;
    uint32_t s = kNoSlot;
    if (c < A.nparents * 12 && c < A.limit) s = A.cand_slot[c];
    const bool win = s != kNoSlot;
    uint32_t total;
    const uint32_t rank = block_rank(win, total);
    if (win) {
        const uint64_t p = A.head + c / 12;
        const int a = (int)(c % 12);
        Key<W> child, pk;
        int L;
        child_of<W>(A, p, a, child, pk, L);
        const uint64_t idx = A.n_nodes + A.block_cnt[blockIdx.x] + rank;
        store_key<W>(A.keys, idx, child);
        A.parent[idx] = (p << 4) | (uint64_t)a;
        A.table[s] = (A.table[s] & ~kIdxMask) | (idx + 1);
    }
    if (threadIdx.x == 0 && total) atomicAdd(&A.ctrl->committed, (unsigned long long)total);
}

// path of node `node` from the root followed by (extra_action, extra_len); one thread
template <int W>
__global__ void bfs_path_kernel(const uint64_t* keys, const uint64_t* parent, uint64_t node, int extra_action,
                                int extra_len, int32_t* path, int path_cap, int32_t* path_len) {
    int depth = 0;
    for (uint64_t q = node;; q = parent[q] >> 4) {
        ++depth;
        if (parent[q] == kNone) break;
    }
    *path_len = depth + 1;
    int pos = depth - 1;
    for (uint64_t q = node;; q = parent[q] >> 4, --pos) {
        const Key<W> k = load_key<W>(keys, q);
        const int L = (int)(k.k[W - 1] >> 58) + (int)(k.k[2 * W - 1] >> 58);
        const bool root = parent[q] == kNone;
        if (pos < path_cap) {
            path[2 * pos] = root ? -1 : (int)(parent[q] & 15);
            path[2 * pos + 1] = L;
        }
        if (root) break;
    }
    if (depth < path_cap) {
        path[2 * depth] = extra_action;
        path[2 * depth + 1] = extra_len;
    }
}

}  // namespace acs

// ---------------------------------------------------------------------------------------
using namespace acs;

namespace {
// error text goes to the library-wide thread-local slot read by acs_last_error()
struct ErrProxy {
    ErrProxy& operator=(const std::string& m) {
        acs::set_last_error(m.c_str());
        return *this;
    }
    ErrProxy& operator=(const char* m) {
        acs::set_last_error(m);
        return *this;
    }
} g_bfs_err;
}  // namespace

struct acs_bfs {
    int device = 0;
    int mrl = 0, W = 1, cyclical = 0;
    int64_t budget = 0;
    uint64_t cap = 0, tcap = 0, n_nodes = 0;
    uint64_t* keys = nullptr;
    uint64_t* parent = nullptr;
    uint64_t* table = nullptr;
    uint32_t* cand_slot = nullptr;
    uint32_t* block_cnt = nullptr;
    BfsCtrl* ctrl = nullptr;
    BfsCtrl* h_ctrl = nullptr;  // pinned
    int32_t* d_path = nullptr;
    int path_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t chunk_cap = 0;
};

#define BFS_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            g_bfs_err = std::string(#call) + ": " + cudaGetErrorString(e__);        \
            return e__ == cudaErrorMemoryAllocation ? ACS_ERR_NOMEM : ACS_ERR_CUDA; \
        }                                                                           \
    } while (0)

extern "C" int acs_bfs_create(acs_ctx* /*ctx*/, int device, int mrl, int64_t max_nodes, int cyclical, acs_bfs** out) {
    if (!out) return ACS_ERR_INVALID;
    *out = nullptr;
    if (mrl < 1 || mrl > 61) {
        g_bfs_err = "bfs needs 1 <= max_relator_length <= 61";
        return ACS_ERR_UNSUPPORTED;
    }
    if (max_nodes < 0) return ACS_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        g_bfs_err = "no CUDA device visible; there is no CPU fallback";
        return ACS_ERR_NO_DEVICE;
    }
    BFS_CUDA(cudaSetDevice(device));
    acs_bfs* b = new acs_bfs();
    b->device = device;
    b->mrl = mrl;
    b->W = mrl <= 29 ? 1 : 2;
    b->cyclical = cyclical ? 1 : 0;
    b->budget = max_nodes;
    b->cap = (uint64_t)max_nodes + 16;  // the budget test runs after all 12 children: overshoot <= 11
    uint64_t t = 1024;
    while (t < 2 * b->cap) t <<= 1;
    b->tcap = t;
    if (t > (1ull << 31)) {  // slot indices are kept in 32 bits (0xFFFFFFFF = "no slot")
        delete b;
        g_bfs_err = "bfs: node budget above 2^30 needs the sharded search (acs_sbfs_*, search/sharded.py)";
        return ACS_ERR_UNSUPPORTED;
    }
    b->chunk_cap = std::min<uint64_t>((uint64_t)kMaxChunkParents, b->cap);
    const uint64_t nblocks = (b->chunk_cap + kParentsPerBlock - 1) / kParentsPerBlock;
    b->path_cap = 1 << 16;
    auto cleanup = [&](int rc) {
        acs_bfs_destroy(b);
        return rc;
    };
#define BFS_ALLOC(ptr, bytes)                                                      \
    do {                                                                           \
        cudaError_t e__ = cudaMalloc((void**)&(ptr), (bytes));                     \
        if (e__ != cudaSuccess) {                                                  \
            g_bfs_err = std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(e__); \
            cudaGetLastError();                                                    \
            return cleanup(ACS_ERR_NOMEM);                                         \
        }                                                                          \
    } while (0)
    BFS_ALLOC(b->keys, b->cap * 2 * b->W * sizeof(uint64_t));
    BFS_ALLOC(b->parent, b->cap * sizeof(uint64_t));
    BFS_ALLOC(b->table, b->tcap * sizeof(uint64_t));
    BFS_ALLOC(b->cand_slot, b->chunk_cap * 12 * sizeof(uint32_t));
    BFS_ALLOC(b->block_cnt, (nblocks + 1) * sizeof(uint32_t));
    BFS_ALLOC(b->ctrl, sizeof(BfsCtrl));
    BFS_ALLOC(b->d_path, (size_t)b->path_cap * 2 * sizeof(int32_t) + 16);
#undef BFS_ALLOC
    if (cudaMallocHost((void**)&b->h_ctrl, sizeof(BfsCtrl)) != cudaSuccess) return cleanup(ACS_ERR_NOMEM);
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(ACS_ERR_CUDA);
    cudaEventCreate(&b->ev0);
    cudaEventCreate(&b->ev1);
    *out = b;
    return ACS_OK;
}

extern "C" void acs_bfs_destroy(acs_bfs* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) {
        cudaStreamSynchronize(b->stream);
        cudaStreamDestroy(b->stream);
    }
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    cudaFree(b->keys);
    cudaFree(b->parent);
    cudaFree(b->table);
    cudaFree(b->cand_slot);
    cudaFree(b->block_cnt);
    cudaFree(b->ctrl);
    cudaFree(b->d_path);
    if (b->h_ctrl) cudaFreeHost(b->h_ctrl);
    delete b;
}

namespace {

template <int W>
int bfs_run_impl(acs_bfs* b, const int8_t* h_presentation, int32_t* h_path, int path_cap, acs_search_result* res) {
    std::memset(res, 0, sizeof(*res));
    BFS_CUDA(cudaSetDevice(b->device));
    Key<W> root;
    int lens[2];
    bool valid;
    if (!pack_root<W>(h_presentation, b->mrl, root, lens, valid)) {
        g_bfs_err = "bfs: letters outside {+-1,+-2} are not supported by the packed search";
        return ACS_ERR_UNSUPPORTED;
    }
    if (!valid) {  // breadth_first.py:36-38 asserts is_array_valid_presentation
        res->status = ACS_ROW_ASSERT;
        return ACS_OK;
    }
    cudaStream_t s = b->stream;
    BFS_CUDA(cudaMemsetAsync(b->table, 0, b->tcap * sizeof(uint64_t), s));
    const uint64_t h = key_hash<W>(root);
    const uint64_t slot_val = ((h >> 40) << 40) | 1ull;  // node index 0, stored +1
    const uint64_t none = kNone;
    BFS_CUDA(cudaMemcpyAsync(b->keys, &root, sizeof(root), cudaMemcpyHostToDevice, s));
    BFS_CUDA(cudaMemcpyAsync(b->parent, &none, sizeof(none), cudaMemcpyHostToDevice, s));
    BFS_CUDA(cudaMemcpyAsync(b->table + (h & (b->tcap - 1)), &slot_val, sizeof(slot_val), cudaMemcpyHostToDevice, s));
    BFS_CUDA(cudaEventRecord(b->ev0, s));

    BfsArgs A{};
    A.keys = b->keys;
    A.parent = b->parent;
    A.table = b->table;
    A.tmask = b->tcap - 1;
    A.cand_slot = b->cand_slot;
    A.block_cnt = b->block_cnt;
    A.ctrl = b->ctrl;
    A.budget = (uint64_t)b->budget;
    A.mrl = b->mrl;
    A.cyclical = b->cyclical;

    uint64_t n_nodes = 1, head = 0, level_end = 1;
    int min_len = lens[0] + lens[1];
    int levels = 0;
    bool solved = false, budget_hit = false;
    int err_status = 0;
    uint64_t n_expanded = 0, sol_gid = kNone;
    std::vector<std::pair<uint64_t, int>> minlen_events;  // (global candidate id, length)

    while (head < n_nodes && !solved && !budget_hit && !err_status) {
        if (head == level_end) {
            level_end = n_nodes;
            ++levels;
        }
        // A chunk plants up to 12*F tentative entries beside the n_nodes committed ones: keep
        // the table at most 3/4 full (tcap >= 2*cap, so at least cap/24 parents always fit).
        const uint64_t room = (3 * (b->tcap / 4) > n_nodes) ? (3 * (b->tcap / 4) - n_nodes) / 12 : 0;
        const uint64_t F = std::min<uint64_t>(std::min<uint64_t>(level_end - head, b->chunk_cap),
                                              std::max<uint64_t>(room, 1));
        const unsigned nblocks = (unsigned)((F + kParentsPerBlock - 1) / kParentsPerBlock);
        // reset the control block
        for (auto& v : b->h_ctrl->first_len) v = kNone;
        b->h_ctrl->sol = kNone;
        b->h_ctrl->err = kNone;
        b->h_ctrl->total = 0;
        b->h_ctrl->committed = 0;
        b->h_ctrl->cut = kNone;
        BFS_CUDA(cudaMemcpyAsync(b->ctrl, b->h_ctrl, sizeof(BfsCtrl), cudaMemcpyHostToDevice, s));
        A.head = head;
        A.nparents = F;
        A.n_nodes = n_nodes;
        A.min_len = min_len;
        A.trusted = head > 0 ? 1 : 0;
        A.limit = F * 12;
        bfs_expand_kernel<W><<<nblocks, kBfsThreads, 0, s>>>(A);
        bfs_mark_kernel<<<nblocks, kBfsThreads, 0, s>>>(A);
        bfs_scan_kernel<<<1, 1024, 0, s>>>(b->block_cnt, nblocks, b->ctrl);
        BFS_CUDA(cudaGetLastError());
        BFS_CUDA(cudaMemcpyAsync(b->h_ctrl, b->ctrl, sizeof(BfsCtrl), cudaMemcpyDeviceToHost, s));
        BFS_CUDA(cudaStreamSynchronize(s));
        const uint64_t total = b->h_ctrl->total;
        uint64_t limit = F * 12;
        bool cut = false;
        if (n_nodes + total >= (uint64_t)b->budget) {
            bfs_cut_kernel<<<nblocks, kBfsThreads, 0, s>>>(A);
            BFS_CUDA(cudaGetLastError());
            BFS_CUDA(cudaMemcpyAsync(&b->h_ctrl->cut, &b->ctrl->cut, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BFS_CUDA(cudaStreamSynchronize(s));
            if (b->h_ctrl->cut != kNone) {
                cut = true;
                limit = std::min<uint64_t>(limit, 12 * (b->h_ctrl->cut + 1));
            }
        }
        bool sol_here = false;
        if (b->h_ctrl->sol != kNone) {
            const uint64_t sc = b->h_ctrl->sol - head * 12;
            if (sc < limit) {  // solved before (or at) the cut parent: Appendix B step 5
                limit = sc;
                sol_here = true;
                cut = false;
            }
        }
        if (b->h_ctrl->err != kNone && (b->h_ctrl->err & 3) == 3) {
            g_bfs_err = "bfs: internal error, visited table overflow";
            return ACS_ERR_CUDA;
        }
        if (b->h_ctrl->err != kNone) {
            const uint64_t ec = (b->h_ctrl->err >> 2) - head * 12;
            if (ec < limit) {  // the reference raises here, before anything later happens
                err_status = (int)(b->h_ctrl->err & 3);
                limit = ec;
                sol_here = false;
                cut = false;
            }
        }
        // "New minimal length found" events in reference order
        const uint64_t gid_limit = head * 12 + limit + (sol_here ? 1 : 0);
        {
            std::vector<std::pair<uint64_t, int>> ev;
            for (int L = 0; L < 128; ++L)
                if (b->h_ctrl->first_len[L] != kNone && b->h_ctrl->first_len[L] < gid_limit)
                    ev.emplace_back(b->h_ctrl->first_len[L], L);
            std::sort(ev.begin(), ev.end());
            for (auto& e : ev)
                if (e.second < min_len) {
                    min_len = e.second;
                    minlen_events.push_back(e);
                }
        }
        A.limit = limit;
        bfs_commit_kernel<W><<<nblocks, kBfsThreads, 0, s>>>(A);
        BFS_CUDA(cudaGetLastError());
        uint64_t committed = total;
        if (limit < F * 12) {
            BFS_CUDA(cudaMemcpyAsync(&b->h_ctrl->committed, &b->ctrl->committed, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            BFS_CUDA(cudaStreamSynchronize(s));
            committed = b->h_ctrl->committed;
        }
        n_nodes += committed;
        if (sol_here) {
            solved = true;
            sol_gid = b->h_ctrl->sol;
            n_expanded = sol_gid / 12 + 1;
        } else if (err_status) {
            n_expanded = (b->h_ctrl->err >> 2) / 12 + 1;
        } else if (cut) {
            budget_hit = true;
            n_expanded = head + b->h_ctrl->cut + 1;
        } else {
            n_expanded = head + F;
        }
        head += F;
    }
    BFS_CUDA(cudaEventRecord(b->ev1, s));
    b->n_nodes = n_nodes;
    res->solved = solved ? 1 : 0;
    res->status = err_status;
    res->budget_hit = budget_hit ? 1 : 0;
    res->n_visited = (int64_t)n_nodes;
    res->n_expanded = (int64_t)n_expanded;
    res->n_moves = solved ? (int64_t)(sol_gid + 1)
                          : (err_status ? (int64_t)((b->h_ctrl->err >> 2) + 1) : (int64_t)(n_expanded * 12));
    res->frontier_left = (int64_t)(n_nodes - n_expanded);
    res->n_levels = levels;
    res->n_minlen = (int32_t)std::min<size_t>(minlen_events.size(), 128);
    for (int i = 0; i < res->n_minlen; ++i) res->minlen_log[i] = minlen_events[i].second;
    if (solved) {
        const int cap = std::min(path_cap, b->path_cap);
        int32_t* d_len = b->d_path + 2 * (size_t)b->path_cap;
        bfs_path_kernel<W><<<1, 1, 0, s>>>(b->keys, b->parent, sol_gid / 12, (int)(sol_gid % 12), 2, b->d_path, cap, d_len);
        BFS_CUDA(cudaGetLastError());
        int32_t plen = 0;
        BFS_CUDA(cudaMemcpyAsync(&plen, d_len, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        BFS_CUDA(cudaStreamSynchronize(s));
        res->path_len = plen;
        if (h_path && cap > 0)
            BFS_CUDA(cudaMemcpy(h_path, b->d_path, (size_t)std::min(plen, cap) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    BFS_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    res->seconds_device = ms * 1e-3;
    return ACS_OK;
}

template <int W>
int bfs_visited_impl(acs_bfs* b, int8_t* h_out, int64_t cap_rows, int64_t* n_out) {
    BFS_CUDA(cudaSetDevice(b->device));
    const uint64_t n = std::min<uint64_t>(b->n_nodes, (uint64_t)std::max<int64_t>(cap_rows, 0));
    if (n_out) *n_out = (int64_t)n;
    if (n == 0) return ACS_OK;
    int8_t* d = nullptr;
    BFS_CUDA(cudaMalloc((void**)&d, n * 2 * b->mrl));
    keys_unpack_kernel<W><<<(unsigned)((n + 255) / 256), 256, 0, b->stream>>>(b->keys, d, n, b->mrl);
    cudaError_t e = cudaMemcpyAsync(h_out, d, n * 2 * b->mrl, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    cudaFree(d);
    BFS_CUDA(e);
    return ACS_OK;
}

}  // namespace

extern "C" int acs_bfs_run(acs_bfs* b, const int8_t* h_presentation, int32_t* h_path, int path_cap,
                           acs_search_result* res) {
    if (!b || !h_presentation || !res) return ACS_ERR_INVALID;
    return b->W == 1 ? bfs_run_impl<1>(b, h_presentation, h_path, path_cap, res)
                     : bfs_run_impl<2>(b, h_presentation, h_path, path_cap, res);
}

extern "C" int acs_bfs_visited(acs_bfs* b, int8_t* h_out, int64_t cap_rows, int64_t* n_out) {
    if (!b || (!h_out && cap_rows > 0)) return ACS_ERR_INVALID;
    return b->W == 1 ? bfs_visited_impl<1>(b, h_out, cap_rows, n_out) : bfs_visited_impl<2>(b, h_out, cap_rows, n_out);
}
