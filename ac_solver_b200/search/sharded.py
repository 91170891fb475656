"""Hash-partitioned breadth-first search over several GPUs of one node -- the NCCL all-to-all
variant with a host-driven chunk loop (round 1).  The default multi-GPU search is now the native
driver in ``search/partitioned.py`` / ``csrc/pbfs.cu`` (no host synchronisation per chunk, the
exchange fused into the expansion kernel by peer stores); this module stays as the
``torch.distributed`` collective baseline it is measured against, and for its gloo CPU test.

One process per GPU (``torchrun``); ``torch.distributed`` (NCCL over NVLink) carries the only
data-path collective -- an all-to-all of newly generated (state key, candidate id) records to
their owner rank -- plus three small all-reduces per chunk (minima of the control block, the
winner bitmap, the table room).  The visited set and the node store are partitioned by
``owner(state) = hash(key) mod world``.

The result is bit-identical to the single-GPU search (csrc/bfs.cu) and hence to the reference's
sequential ``bfs`` (ac_solver/search/breadth_first.py:15-97) for every world size: nodes carry
GLOBAL ids that equal their FIFO position, candidates are labelled ``12*(gid-head)+action`` and
the smallest label wins every tie (duplicate states, solving child, budget cut, raising move),
see csrc/sbfs.cu for the per-rank kernels.

The chunk loop below is pure host logic on small scalars; everything touching node data goes
through a ``ShardOps`` object.  ``GpuShardOps`` drives the CUDA kernels through the C ABI; the
CPU test-suite injects a numpy implementation to exercise this loop under gloo with
world_size 2 (tests/test_sharded_cpu.py).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib

I64_MAX = np.iinfo(np.int64).max


# --------------------------------------------------------------------------------------------
# collectives (no-ops for a single process)
# --------------------------------------------------------------------------------------------
class _Comm:
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1

    def allreduce(self, t, op="sum"):
        if self.on and self.world > 1:
            ops = {"sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN, "max": self.dist.ReduceOp.MAX}
            self.dist.all_reduce(t, op=ops[op], group=self.group)
        return t

    def alltoall_counts(self, counts, device):
        t = self.torch.as_tensor(counts, dtype=self.torch.int64, device=device)
        if not (self.on and self.world > 1):
            return t.clone()
        out = self.torch.empty_like(t)
        self.dist.all_to_all_single(out, t, group=self.group)
        return out

    def alltoall_v(self, send, send_counts, recv_counts):
        """Variable all-to-all of the rows of ``send`` (rows grouped by destination rank)."""
        if not (self.on and self.world > 1):
            return send
        n_recv = int(sum(recv_counts))
        out = self.torch.empty((n_recv,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(out, send, output_split_sizes=[int(x) for x in recv_counts],
                                    input_split_sizes=[int(x) for x in send_counts], group=self.group)
        return out


# --------------------------------------------------------------------------------------------
# per-rank operations on the GPU
# --------------------------------------------------------------------------------------------
class GpuShardOps:
    """Owns this rank's shard (torch CUDA tensors) and calls the csrc/sbfs.cu kernels."""

    def __init__(self, rank, world, mrl, cyclical, budget, device=None, cap_local=None):
        import torch

        self.torch = torch
        self.L = _lib.lib()
        self.rank, self.world, self.mrl, self.cyclical, self.budget = rank, world, mrl, int(bool(cyclical)), int(budget)
        self.W = 1 if mrl <= 29 else 2
        if not torch.cuda.is_available():
            raise _lib.AcsError("no CUDA device visible; the sharded search has no CPU fallback")
        self.dev = torch.device("cuda", _lib.default_device() if device is None else device)
        torch.cuda.set_device(self.dev)
        if cap_local is None:  # hash partitioning is balanced to a few sigma of sqrt(n/world)
            share = (self.budget + 16 + world - 1) // world
            cap_local = int(share * 1.1) + 200_000 if world > 1 else self.budget + 16
        self.cap = int(cap_local)
        tcap = 1024
        while tcap < 2 * self.cap:
            tcap <<= 1
        self.tcap = tcap
        if tcap > (1 << 31):  # slot indices travel in 32 bits (csrc/sbfs.cu rec_slot)
            raise _lib.AcsError(f"shard of {self.cap} nodes per rank is too large (table > 2^31 slots): use more GPUs")
        i64 = dict(dtype=torch.int64, device=self.dev)
        self.keys = torch.empty((self.cap, 2 * self.W), **i64)
        self.parent = torch.empty(self.cap, **i64)
        self.gid = torch.empty(self.cap, **i64)
        self.table = torch.zeros(self.tcap, **i64)
        self.n_local = 0
        self.ctrl = torch.empty(130, **i64)
        self.small = torch.zeros(8, **i64)
        self.args = _lib.SbfsArgs()
        self._bufs = {}

    # -- helpers --
    def _buf(self, name, n, dtype, cols=None):
        """Grow-only scratch tensor reused across chunks (no allocator traffic in the loop)."""
        n = max(int(n), 1)
        cur = self._bufs.get(name)
        if cur is None or cur.shape[0] < n:
            shape = (int(n * 1.25) + 16,) if cols is None else (int(n * 1.25) + 16, cols)
            cur = self.torch.empty(shape, dtype=dtype, device=self.dev)
            self._bufs[name] = cur
        return cur[:n]

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def _fill(self, **kw):
        a = self.args
        a.keys, a.parent, a.gid, a.table = (self.keys.data_ptr(), self.parent.data_ptr(), self.gid.data_ptr(),
                                            self.table.data_ptr())
        a.tmask = self.tcap - 1
        a.n_local = self.n_local
        a.mrl, a.cyclical, a.world, a.rank, a.W = self.mrl, self.cyclical, self.world, self.rank, self.W
        a.budget = self.budget
        a.ctrl = self.ctrl.data_ptr()
        for k, v in kw.items():
            setattr(a, k, v)
        return C.byref(a)

    def room(self):
        """Parents per chunk this rank's table can absorb (kept <= 3/4 full, with slack for skew)."""
        free = (3 * (self.tcap // 4)) - self.n_local
        return max(int(free * self.world // 24), 1)

    def add_root(self, presentation):
        p8 = np.ascontiguousarray(presentation, dtype=np.int8)
        key = (C.c_uint64 * 4)()
        h = C.c_uint64()
        total, valid = C.c_int(), C.c_int()
        rc = self.L.acs_sbfs_pack_root(p8.ctypes.data, self.mrl, key, C.byref(h), C.byref(total), C.byref(valid))
        if rc == -3:
            raise ValueError("the GPU search supports the two-generator alphabet {+-1, +-2} only")
        _lib.check(rc)
        if valid.value and self.L.acs_sbfs_owner(h.value, self.world) == self.rank:
            t = self.torch
            k = np.array([key[i] for i in range(2 * self.W)], dtype=np.uint64).view(np.int64)
            self.keys[0] = t.from_numpy(k).to(self.dev)
            self.parent[0] = -1
            self.gid[0] = 0
            slot_val = np.array([((h.value >> 41) << 40) | 1], dtype=np.uint64).view(np.int64)
            self.table[h.value & (self.tcap - 1)] = int(slot_val[0])
            self.n_local = 1
        return total.value, bool(valid.value)

    def begin_chunk(self, head, nparents, n_nodes, min_len, trusted):
        t = self.torch
        self.head, self.F, self.n_nodes = int(head), int(nparents), int(n_nodes)
        self.min_len, self.trusted = int(min_len), int(bool(trusted))
        # local parents of the chunk: owned nodes with head <= gid < head+F (gid is increasing)
        for k, v in enumerate((self.head, self.head + self.F)):
            _lib.check(self.L.acs_sbfs_lower_bound(self.gid.data_ptr(), self.n_local, v,
                                                   self.small[k : k + 1].data_ptr(), self._stream()))
        lo_hi = self.small[:2].cpu()
        self.l0, self.l1 = int(lo_hi[0]), int(lo_hi[1])
        self.nwords = (12 * self.F + 31) // 32
        self.bitmap_local = self._buf("bitmap_local", self.nwords, t.int32).zero_()
        nscr = self.nwords + 2 + (self.nwords + 2047) // 2048  # prefixes + total + scan scratch
        self.prefix_local = self._buf("prefix_local", nscr, t.int32)
        self.prefix_global = self._buf("prefix_global", nscr, t.int32)

    def _chunk_args(self, **kw):
        return self._fill(l0=self.l0, l1=self.l1, head=self.head, nparents=self.F, n_nodes=self.n_nodes,
                          min_len=self.min_len, trusted=self.trusted, **kw)

    def expand_count(self):
        t = self.torch
        self.ctrl.fill_(-1)  # all-ones == "none" for the unsigned atomicMin
        self.dest_count = t.zeros(self.world, dtype=t.int64, device=self.dev)
        _lib.check(self.L.acs_sbfs_expand(self._chunk_args(dest_count=self.dest_count.data_ptr()), 0, self._stream()))
        counts = self.dest_count.cpu().numpy().astype(np.int64)
        ctrl = self.ctrl.clone()
        ctrl[ctrl < 0] = I64_MAX  # signed MIN all-reduce: "none" must be the largest value
        return counts, ctrl

    def expand_scatter(self, counts):
        t = self.torch
        n = int(counts.sum())
        offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        cursor = t.from_numpy(offs).to(self.dev)
        send_keys = self._buf("send_keys", n, t.int64, cols=2 * self.W)
        send_c = self._buf("send_c", n, t.int32)
        _lib.check(self.L.acs_sbfs_expand(self._chunk_args(dest_cursor=cursor.data_ptr(),
                                                           send_keys=send_keys.data_ptr(),
                                                           send_c=send_c.data_ptr()), 1, self._stream()))
        return send_keys[:n], send_c[:n]

    def insert_mark(self, recv_keys, recv_c):
        t = self.torch
        self.recv_keys, self.recv_c = recv_keys.contiguous(), recv_c.contiguous()
        self.n_recv = int(self.recv_c.shape[0])
        self.rec_slot = self._buf("rec_slot", self.n_recv, t.int32)
        _lib.check(self.L.acs_sbfs_insert_mark(self._recv_args(), self._stream()))
        out = self._buf("bitmap_global", self.nwords, t.int32)
        out.copy_(self.bitmap_local)
        return out

    def _recv_args(self, **kw):
        return self._chunk_args(recv_keys=self.recv_keys.data_ptr(), recv_c=self.recv_c.data_ptr(), n_recv=self.n_recv,
                                rec_slot=self.rec_slot.data_ptr(), bitmap_local=self.bitmap_local.data_ptr(),
                                prefix_local=self.prefix_local.data_ptr(), **kw)

    def finish(self, bitmap_global):
        """Scan both bitmaps; returns (global winners, table-overflow flag)."""
        self.bitmap_global = bitmap_global.contiguous()
        s = self._stream()
        _lib.check(self.L.acs_sbfs_scan(self.bitmap_local.data_ptr(), self.prefix_local.data_ptr(), self.nwords, s))
        _lib.check(self.L.acs_sbfs_scan(self.bitmap_global.data_ptr(), self.prefix_global.data_ptr(), self.nwords, s))
        return int(self.prefix_global[self.nwords].item())

    def overflow_flag(self):
        """1 if this rank's insert kernel ran out of table slots in the current chunk (it reports
        through ctrl[1] AFTER the control block was cloned for the all-reduce), or the shard is full."""
        t = self.torch
        over = (self.ctrl[1] == 3).to(t.int64).reshape(1)
        return over

    def find_cut(self):
        self.small[2] = -1
        a = self._recv_args(bitmap_global=self.bitmap_global.data_ptr(), prefix_global=self.prefix_global.data_ptr(),
                            cut=self.small[2:3].data_ptr())
        _lib.check(self.L.acs_sbfs_cut(a, self._stream()))
        v = int(self.small[2].item())
        return None if v < 0 else v

    def commit(self, limit):
        a = self._recv_args(bitmap_global=self.bitmap_global.data_ptr(), prefix_global=self.prefix_global.data_ptr(),
                            limit=int(limit))
        s = self._stream()
        _lib.check(self.L.acs_sbfs_rank_at(a, self.small[4:6].data_ptr(), s))
        cg, cl = (int(x) for x in self.small[4:6].cpu())
        full = self.torch.tensor([1 if self.n_local + cl > self.cap else 0], dtype=self.torch.int64, device=self.dev)
        if self.world > 1 and self.torch.distributed.is_initialized():  # raise on every rank together
            self.torch.distributed.all_reduce(full, op=self.torch.distributed.ReduceOp.MAX)
        if int(full.item()):
            raise _lib.AcsError(f"shard capacity {self.cap} exceeded on a rank (skewed partition); pass a larger cap_local")
        _lib.check(self.L.acs_sbfs_commit(a, s))
        self.n_local += cl
        return cg, cl

    def lookup(self, gid):
        out = self.small[:4]
        _lib.check(self.L.acs_sbfs_lookup(self._fill(), int(gid), out.data_ptr(), self._stream()))
        return out.clone()

    def visited(self):
        t = self.torch
        rows = t.empty((max(self.n_local, 1), 2 * self.mrl), dtype=t.int8, device=self.dev)
        _lib.check(self.L.acs_sbfs_unpack(self.keys.data_ptr(), rows.data_ptr(), self.n_local, self.mrl, self._stream()))
        return self.gid[: self.n_local].cpu().numpy(), rows[: self.n_local].cpu().numpy()

    @property
    def device(self):
        return self.dev


# --------------------------------------------------------------------------------------------
# the chunk loop (host logic shared by every ShardOps implementation)
# --------------------------------------------------------------------------------------------
def bfs_sharded(presentation, max_nodes_to_explore=10000, cyclically_reduce_after_moves=False, group=None,
                ops_factory=None, want_visited=False, chunk_parents=1 << 22, verbose=False):
    """Sharded ``bfs``: call on every rank of the process group with the same arguments.

    Returns ``(solved, path|None, info)`` on every rank (identical values).  ``info["visited"]``
    (rank 0, on request) holds all visited states ordered by global id == the reference's
    insertion order."""
    import torch

    comm = _Comm(group)
    p = np.asarray(presentation)
    if p.ndim != 1 or p.size % 2 or p.size == 0:
        raise AssertionError(f"{presentation} is not a valid presentation")
    mrl = p.size // 2
    budget = int(max_nodes_to_explore)
    factory = ops_factory or GpuShardOps
    ops = factory(comm.rank, comm.world, mrl, bool(cyclically_reduce_after_moves), budget)
    L0, valid = ops.add_root(p)
    if not valid:
        raise AssertionError(f"{presentation} is not a valid presentation")  # breadth_first.py:36-38
    dev = ops.device

    n_nodes, head, level_end, levels = 1, 0, 1, 0
    min_len, minlen_log = L0, []
    solved = budget_hit = False
    status, n_expanded, sol_gid = 0, 0, None
    while head < n_nodes and not (solved or budget_hit or status):
        if head == level_end:
            level_end, levels = n_nodes, levels + 1
        room = int(comm.allreduce(torch.tensor([ops.room()], dtype=torch.int64, device=dev), "min").item())
        F = max(1, min(level_end - head, int(chunk_parents), room, (1 << 26) // 12 - 1))
        ops.begin_chunk(head, F, n_nodes, min_len, trusted=head > 0)
        counts, ctrl = ops.expand_count()
        ctrl = comm.allreduce(ctrl, "min").cpu().numpy()
        recv_counts = comm.alltoall_counts(counts, dev).cpu().numpy()
        send_keys, send_c = ops.expand_scatter(counts)
        recv_keys = comm.alltoall_v(send_keys, counts, recv_counts)
        recv_c = comm.alltoall_v(send_c, counts, recv_counts)
        bitmap = comm.allreduce(ops.insert_mark(recv_keys, recv_c), "sum")
        total = ops.finish(bitmap)
        if hasattr(ops, "overflow_flag"):  # every rank must stop together (a lone raise would hang the others)
            if int(comm.allreduce(ops.overflow_flag(), "max").item()):
                raise _lib.AcsError("sharded bfs: visited table overflow on one rank (skewed shard)")
        sol = None if ctrl[0] == I64_MAX else int(ctrl[0])
        err = None if ctrl[1] == I64_MAX else int(ctrl[1])
        if err is not None and (err & 3) == 3:
            raise _lib.AcsError("sharded bfs: visited table overflow (internal error)")
        limit, cut = 12 * F, None
        if n_nodes + total >= budget:
            cut = ops.find_cut()
            if cut is not None:
                limit = min(limit, 12 * (cut + 1))
        sol_here = False
        if sol is not None and sol - 12 * head < limit:  # solved at or before the cut parent
            limit, sol_here, cut = sol - 12 * head, True, None
        if err is not None and (err >> 2) - 12 * head < limit:  # the reference raises here
            limit, status, sol_here, cut = (err >> 2) - 12 * head, err & 3, False, None
        gid_limit = 12 * head + limit + (1 if sol_here else 0)
        events = sorted((int(ctrl[2 + ln]), ln) for ln in range(128) if ctrl[2 + ln] != I64_MAX and ctrl[2 + ln] < gid_limit)
        for _, ln in events:
            if ln < min_len:
                min_len = ln
                minlen_log.append(ln)
        committed, _ = ops.commit(limit)
        n_nodes += committed
        if sol_here:
            solved, sol_gid, n_expanded = True, sol, sol // 12 + 1
        elif status:
            n_expanded = (err >> 2) // 12 + 1
        elif cut is not None:
            budget_hit, n_expanded = True, head + cut + 1
        else:
            n_expanded = head + F
        head += F

    info = {
        "n_visited": n_nodes, "n_expanded": n_expanded, "budget_hit": budget_hit, "n_levels": levels,
        "frontier_left": n_nodes - n_expanded, "status": status, "minlen_log": minlen_log,
        "n_moves": (sol_gid + 1) if solved else ((err >> 2) + 1 if status else 12 * n_expanded),
        "world": comm.world, "n_local": ops.n_local,
    }
    path = None
    if solved:  # walk the parent chain; exactly one rank owns each global id
        chain, g = [], sol_gid // 12
        while g >= 0:
            rec = comm.allreduce(ops.lookup(g), "sum").cpu().numpy()
            assert rec[0] == 1, "global id not found on exactly one rank"
            chain.append((int(rec[2]), int(rec[3])))
            g = int(rec[1])
        path = chain[::-1] + [(sol_gid % 12, 2)]
    if want_visited and not status:
        gids, rows = ops.visited()
        if comm.on and comm.world > 1:
            gathered = [None] * comm.world if comm.rank == 0 else None
            comm.dist.gather_object((gids, rows), gathered, dst=0, group=group)
            if comm.rank == 0:
                gids = np.concatenate([g for g, _ in gathered])
                rows = np.concatenate([r for _, r in gathered])
        if comm.rank == 0:
            order = np.argsort(gids, kind="stable")
            assert np.array_equal(gids[order], np.arange(n_nodes)), "global ids are not a permutation"
            info["visited"] = rows[order]
    if status == _lib.ROW_ASSERT:
        raise AssertionError("a move produced an invalid presentation (empty relator)")
    if status == _lib.ROW_INDEX:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    if verbose and comm.rank == 0:
        for m in minlen_log:
            print(f"New minimal length found: {m}")
    if budget_hit and comm.rank == 0:
        print(f"Exiting search as number of explored nodes = {n_nodes} has exceeded the limit {budget}")
    return solved, path, info
