"""TEST-ONLY numpy implementation of the ShardOps interface (see ac_solver_b200/search/sharded.py).

It lets the CPU test-suite run the real chunk loop of the sharded BFS -- partitioning, the
all-to-all exchange, the bitmap all-reduce, global ranking, budget cut, path walk -- under gloo
with world_size 2.  Children are computed with the CPU oracle; nothing here is product code."""

import zlib

import numpy as np
import torch

from oracle import oracle as O

I64_MAX = np.iinfo(np.int64).max


class NumpyShardOps:
    def __init__(self, rank, world, mrl, cyclical, budget):
        self.rank, self.world, self.mrl, self.cyclical, self.budget = rank, world, mrl, bool(cyclical), budget
        self.states, self.parent, self.gid, self.table = [], [], [], {}
        self.device = torch.device("cpu")

    @property
    def n_local(self):
        return len(self.states)

    def owner(self, row):
        return zlib.crc32(np.ascontiguousarray(row, np.int8).tobytes()) % self.world

    def room(self):
        return 5 + 3 * self.rank  # tiny, rank-dependent: forces many chunks and the min all-reduce

    def add_root(self, p):
        p8 = np.ascontiguousarray(p, np.int8)
        if np.abs(p8).max() > 2:
            raise ValueError("alphabet")
        valid = bool(O.lib().aco_is_valid_presentation(p8.ctypes.data, self.mrl))
        if valid and self.owner(p8) == self.rank:
            self.states.append(p8.copy())
            self.parent.append(-1)
            self.gid.append(0)
            self.table[p8.tobytes()] = 0
        return int(np.count_nonzero(p8)), valid

    def begin_chunk(self, head, F, n_nodes, min_len, trusted):
        self.head, self.F, self.n_nodes, self.min_len = head, F, n_nodes, min_len
        self.local = [j for j, g in enumerate(self.gid) if head <= g < head + F]

    def expand_count(self):
        ctrl = np.full(130, I64_MAX, np.int64)
        self.records = [[] for _ in range(self.world)]
        for j in self.local:
            for a in range(12):
                gidc = self.gid[j] * 12 + a
                c = (self.gid[j] - self.head) * 12 + a
                try:
                    child, lens = O.acmove(a, self.states[j], self.mrl, cyclical=self.cyclical)
                except AssertionError:
                    ctrl[1] = min(ctrl[1], (gidc << 2) | 1)
                    continue
                except IndexError:
                    ctrl[1] = min(ctrl[1], (gidc << 2) | 2)
                    continue
                L = sum(lens)
                if L < self.min_len:
                    ctrl[2 + L] = min(ctrl[2 + L], gidc)
                if L == 2:
                    ctrl[0] = min(ctrl[0], gidc)
                if not np.array_equal(child, self.states[j]):
                    self.records[self.owner(child)].append((child, c))
        counts = np.array([len(r) for r in self.records], np.int64)
        return counts, torch.from_numpy(ctrl)

    def expand_scatter(self, counts):
        flat = [rc for r in self.records for rc in r]
        keys = np.zeros((len(flat), 2 * self.mrl), np.int8)
        cs = np.zeros(len(flat), np.int32)
        for i, (k, c) in enumerate(flat):
            keys[i], cs[i] = k, c
        return torch.from_numpy(keys), torch.from_numpy(cs)

    def insert_mark(self, recv_keys, recv_c):
        self.recv_keys, self.recv_c = recv_keys.numpy(), recv_c.numpy()
        best = {}
        for i in range(len(self.recv_c)):
            kb = self.recv_keys[i].tobytes()
            if kb in self.table:
                continue
            c = int(self.recv_c[i])
            if kb not in best or c < best[kb][0]:
                best[kb] = (c, i)
        self.winners = sorted(best.values())
        nwords = (12 * self.F + 31) // 32
        bm = np.zeros(nwords, np.uint32)
        for c, _ in self.winners:
            bm[c >> 5] |= np.uint32(1 << (c & 31))
        return torch.from_numpy(bm.view(np.int32).copy())

    def _bits(self, bm):
        w = bm.numpy().view(np.uint32)
        return np.unpackbits(w.view(np.uint8), bitorder="little")

    def finish(self, bitmap_global):
        self.gbits = self._bits(bitmap_global)
        self.gprefix = np.concatenate([[0], np.cumsum(self.gbits)])
        return int(self.gbits.sum())

    def find_cut(self):
        for p in range(self.F):
            if self.n_nodes + self.gprefix[12 * (p + 1)] >= self.budget:
                return p
        return None

    def commit(self, limit):
        cl = 0
        for c, i in self.winners:
            if c >= limit:
                continue
            row = self.recv_keys[i].copy()
            self.table[row.tobytes()] = len(self.states)
            self.states.append(row)
            self.parent.append(((self.head + c // 12) << 4) | (c % 12))
            self.gid.append(self.n_nodes + int(self.gprefix[c]))
            cl += 1
        return int(self.gprefix[limit]), cl

    def lookup(self, g):
        out = np.zeros(4, np.int64)
        if g in self.gid:
            j = self.gid.index(g)
            out[0] = 1
            out[1] = -1 if self.parent[j] < 0 else self.parent[j] >> 4
            out[2] = -1 if self.parent[j] < 0 else self.parent[j] & 15
            out[3] = np.count_nonzero(self.states[j])
        return torch.from_numpy(out)

    def visited(self):
        return np.array(self.gid, np.int64), (np.stack(self.states) if self.states else np.zeros((0, 2 * self.mrl), np.int8))
