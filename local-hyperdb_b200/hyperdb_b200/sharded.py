"""Row-sharded ranking over the GPUs of one box: one process per GPU (torch.distributed / NCCL over
NVLink 5), rank r holds the contiguous row block [r*N/G, (r+1)*N/G) in a DeviceMatrix.

Per query batch: every rank runs the fused sweep + certify on its shard and emits its local top-k as
(float64 score, GLOBAL row id); ONE exchange step -- an all-gather of 2*B*k + 2*B 8-byte words per
rank -- and a final G*k -> k merge with the same (score desc, id asc) key.  Contiguous shards + global
ids keep the lower-index tie rule exact across shards.  The only other cross-shard dependency is the
time-decay reference max(ts) over kept rows (hyperdb/hyperdb.py:1334-1344): an 8-byte all-reduce(MAX)
whenever timestamps or the row subset change.

The compute engine is injectable so the host logic (bounds, packing, exchange, flag handling) is
covered by world_size-2 gloo tests on CPU with a test-side engine; the product engine is CudaEngine
(no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

from . import _native as N


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous, balanced: the first n_total % world shards get one extra row."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def packed_len(b: int, k: int) -> int:
    """Words (int64) of one rank's message: scores[b*k] | ids[b*k] | counts[b] | flags (b uint32, two per word)."""
    return 2 * b * k + b + (b + 1) // 2


class CudaEngine:
    """Local shard on the current CUDA device, results kept on the device."""

    def __init__(self, matrix):
        import torch
        self.m = matrix
        self.torch = torch
        self.device = torch.device("cuda", matrix.device)
        self.post = None                      # torch stream of the pipelined certify / exchange / merge
        self._path = 0
        self._qdt = {torch.float16: N.HDB_F16, torch.float32: N.HDB_F32, torch.float64: N.HDB_F64}

    def enable_pipeline(self, on=True):
        """Run certify / exchange / merge of query i on a second stream while the sweep of query i+1 streams the matrix."""
        if on and self.post is None:
            # high priority: when a sweep's CTAs leave the SMs, certify / exchange / merge are scheduled before the next
            # query's pending sweep CTAs
            self.post = self.torch.cuda.Stream(device=self.device, priority=-1)
            self.m.set_post_stream(self.post.cuda_stream)
        elif not on and self.post is not None:
            self.torch.cuda.current_stream(self.device).wait_stream(self.post)
            self.m.set_post_stream(0)
            self.post = None

    def kept_ts_max(self):
        return self.m.kept_ts_max()

    def set_decay_reference(self, ts_max):
        self.m.set_decay_reference(ts_max)

    def n_kept(self):
        return self.m.n_kept

    # mutation of the local shard (ShardedMatrix.append / remove_rows)
    def n_rows(self):
        return self.m.shape[0]

    def row_offset(self):
        return self.m.row_offset

    def append(self, rows):
        self.m.append(rows)

    def remove_local(self, local_rows):
        self.m.remove_rows(local_rows)

    def set_row_offset(self, off):
        self.m.set_row_offset(off)

    def attach_exchange(self, xchg):
        """Device-output queries also deliver their results to every rank (fused into the certify kernel)."""
        N.check(N.lib().hdb_matrix_attach_exchange(self.m._h, xchg._h if xchg is not None else None))

    def set_path(self, mode):
        if mode != self._path:
            self.m.set_path(mode)
            self._path = mode

    def local_topk(self, queries, k, metric, bias, exact=False):
        """-> packed int64 CUDA tensor [packed_len(B, k)] (scores bit-cast).  exact: False = automatic path,
        True = exact full-vector path, or a path mode of hdb_matrix_set_path (4 = wide candidate class)."""
        torch = self.torch
        q = queries if torch.is_tensor(queries) else torch.as_tensor(queries)
        if q.device != self.device or not q.is_contiguous():
            q = q.to(self.device).contiguous()
        b = 1 if q.dim() == 1 else q.shape[0]
        buf = torch.empty(packed_len(b, k), dtype=torch.int64, device=self.device)
        # [scores b*k | ids b*k | counts b | flags]: addresses by arithmetic (this runs once per query: no tensor views)
        base = buf.data_ptr()
        self.set_path(int(exact))
        N.check(N.lib().hdb_query(self.m._h, N.METRIC_IDS[metric], q.data_ptr(), self._qdt[q.dtype], N.HDB_DEVICE, b, int(k), float(bias),
                                  base + 8 * b * k, base, base + 16 * b * k, base + 16 * b * k + 8 * b, N.HDB_DEVICE))
        return buf

    def merge(self, gathered, b, k):
        """gathered: int64 CUDA tensor [G, packed_len]; -> (idx [b,k], score [b,k], count [b], flags int32 [G,b])."""
        torch = self.torch
        g, ln = gathered.shape
        flags = gathered[:, 2 * b * k + b:].view(torch.int32)[:, :b]
        if g == 1:           # single shard: the local result is the result
            row = gathered[0]
            return (row[b * k: 2 * b * k].view(b, k), row[: b * k].view(torch.float64).view(b, k),
                    row[2 * b * k: 2 * b * k + b], flags)
        out = torch.empty(2 * b * k + b, dtype=torch.int64, device=self.device)      # one block: one D2H copy later
        idx = out[: b * k].view(b, k)
        sc = out[b * k: 2 * b * k].view(torch.float64).view(b, k)
        cnt = out[2 * b * k:]
        base = gathered.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(N.lib().hdb_merge_topk(self.device.index, C.c_void_p(stream), g, b, k, ln,
                                       C.c_void_p(base), C.c_void_p(base + 8 * b * k), C.c_void_p(base + 16 * b * k),
                                       N.HDB_DEVICE, C.c_void_p(idx.data_ptr()), C.c_void_p(sc.data_ptr()),
                                       C.c_void_p(cnt.data_ptr()), N.HDB_DEVICE))
        return idx, sc, cnt, flags


class PeerExchange:
    """This rank's end of the NVLink peer-memory candidate exchange (include/hyperdb_b200.h, hdb_exchange_*): every
    rank's result block is stored straight into every rank's buffer and merged there -- two tiny kernels per step and
    no collective-library call on the query path (the NCCL all-gather cost more host time than a sharded sweep takes)."""

    def __init__(self, device_index, world, rank, max_words):
        self._h = C.c_void_p()
        self.world, self.rank, self.max_words = int(world), int(rank), int(max_words) + (int(max_words) & 1)
        N.check(N.lib().hdb_exchange_create(int(device_index), self.world, self.rank, self.max_words, C.byref(self._h)))

    def handle(self) -> bytes:
        buf = C.create_string_buffer(N.lib().hdb_exchange_handle_bytes())
        N.check(N.lib().hdb_exchange_local_handle(self._h, buf))
        return buf.raw

    def connect(self, handles):
        """handles: the `handle()` bytes of every rank, indexed by rank (CUDA IPC: one process per GPU)."""
        blob = b"".join(handles)
        N.check(N.lib().hdb_exchange_connect(self._h, C.c_char_p(blob)))

    def local_buffer(self) -> int:
        p = C.c_void_p()
        N.check(N.lib().hdb_exchange_local_buffer(self._h, C.byref(p)))
        return p.value

    def connect_pointers(self, buffers):
        """Same-process form: `local_buffer()` of every rank (several shards driven by one process)."""
        arr = (C.c_void_p * self.world)(*[C.c_void_p(b) for b in buffers])
        N.check(N.lib().hdb_exchange_connect_pointers(self._h, arr))

    def step(self, stream_ptr, mine, b, k, out_idx, out_score, out_count, out_flags):
        N.check(N.lib().hdb_exchange_step(self._h, C.c_void_p(stream_ptr), C.c_void_p(mine.data_ptr()), mine.numel(), b, k,
                                          C.c_void_p(out_idx.data_ptr()), C.c_void_p(out_score.data_ptr()),
                                          C.c_void_p(out_count.data_ptr()), C.c_void_p(out_flags.data_ptr())))

    def push(self, stream_ptr, mine):
        N.check(N.lib().hdb_exchange_push(self._h, C.c_void_p(stream_ptr), C.c_void_p(mine.data_ptr()), mine.numel()))

    def wait_merge(self, stream_ptr, b, k, out_idx, out_score, out_count, out_flags):
        N.check(N.lib().hdb_exchange_wait_merge(self._h, C.c_void_p(stream_ptr), b, k, C.c_void_p(out_idx.data_ptr()),
                                                C.c_void_p(out_score.data_ptr()), C.c_void_p(out_count.data_ptr()),
                                                C.c_void_p(out_flags.data_ptr())))

    def collect_async(self, b, k, out_ptr):
        """Second half of a step whose first half was done by an attached DeviceMatrix (hdb_matrix_attach_exchange): wait +
        merge on the exchange's own stream into the block [idx b*k | score b*k | count b | flags world*b] at out_ptr."""
        N.check(N.lib().hdb_exchange_collect_async(self._h, b, k, out_ptr, out_ptr + 8 * b * k, out_ptr + 16 * b * k,
                                                   out_ptr + 16 * b * k + 8 * b))

    def set_stream(self, stream_ptr):
        """Run wait + merge on a caller-owned stream (it must outlive this exchange)."""
        N.check(N.lib().hdb_exchange_set_stream(self._h, C.c_void_p(stream_ptr)))

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        N.check(N.lib().hdb_exchange_stream(self._h, C.byref(p)))
        return p.value

    def error(self) -> bool:
        e = C.c_int()
        N.check(N.lib().hdb_exchange_error(self._h, C.byref(e)))
        return bool(e.value)

    def close(self):
        if self._h.value:
            N.lib().hdb_exchange_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StepResult:
    """(idx [b,k], score [b,k], count [b], per-shard flags [world,b]) of one sharded query step as views of ONE device
    block [idx | score | count | flags]; the views are only built when somebody looks (a pipelined loop that keeps the
    results on the device pays one allocation per step).  Unpacks like the 4-tuple it replaces."""
    __slots__ = ("block", "b", "k", "w", "_views")

    def __init__(self, block, b, k, w):
        self.block, self.b, self.k, self.w, self._views = block, b, k, w, None

    def views(self):
        if self._views is None:
            import torch
            o, b, k, w = self.block, self.b, self.k, self.w
            self._views = (o[: b * k].view(b, k), o[b * k: 2 * b * k].view(torch.float64).view(b, k), o[2 * b * k: 2 * b * k + b],
                           o[2 * b * k + b:].view(torch.int32)[: w * b].view(w, b))
        return self._views

    def __iter__(self):
        return iter(self.views())

    def __getitem__(self, i):
        return self.views()[i]

    def __len__(self):
        return 4


class LocalStepResult(StepResult):
    """Single shard: the engine's packed block [scores | ids | counts | flags] IS the result; nothing is launched or
    allocated to present it (a 100 microsecond sweep does not wait for 100 microseconds of Python)."""
    __slots__ = ()

    def views(self):
        if self._views is None:
            import torch
            o, b, k = self.block, self.b, self.k
            self._views = (o[b * k: 2 * b * k].view(b, k), o[: b * k].view(torch.float64).view(b, k), o[2 * b * k: 2 * b * k + b],
                           o[2 * b * k + b:].view(torch.int32)[:b].view(1, b))
        return self._views


class ShardedMatrix:
    """engine: the local shard (CudaEngine in production); group: torch.distributed process group."""

    def __init__(self, engine, n_total, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.exchanges = 0
        self.xchg = None                      # PeerExchange (enable_peer_exchange), else torch.distributed all-gather

    def enable_peer_exchange(self, max_batch=64, max_k=128):
        """Exchange candidates through NVLink peer memory instead of an NCCL all-gather (CUDA engines, world > 1).
        Collective: the IPC handles travel once through torch.distributed.  The buffer holds one step of max_batch queries
        x max_k results (packed_len words); hdb_query refuses a larger step, so `query` cuts a larger batch into pieces that fit
        (`exchange_batch_limit`) and callers of `submit` / `query_async` must do the same."""
        if self.world == 1 or self.xchg is not None:
            return self.xchg is not None
        import torch
        dev = self.engine.device
        xchg, err = None, None
        try:
            xchg = PeerExchange(dev.index, self.world, self.rank, packed_len(max_batch, max_k))
            mine = xchg.handle()
        except Exception as e:                                   # noqa: BLE001
            mine, err = None, e
        handles = [None] * self.world
        self.dist.all_gather_object(handles, mine, group=self.group)
        if err is None and all(h is not None for h in handles):
            try:
                xchg.connect(handles)
            except Exception as e:                               # noqa: BLE001
                err = e
        # every rank must take the same decision: one failed mapping switches the whole group back to the all-gather
        ok = torch.tensor([0 if (err is not None or any(h is None for h in handles)) else 1], dtype=torch.int32, device=self._comm_device())
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            if xchg is not None:
                xchg.close()
            if err is not None:
                raise RuntimeError(f"peer-memory exchange unavailable on rank {self.rank}: {err}")
            raise RuntimeError("peer-memory exchange unavailable on another rank")
        self.xchg = xchg
        if hasattr(self.engine, "attach_exchange"):
            # wait + merge run on a torch-owned highest-priority stream: result blocks record their use of it
            # (record_stream) and may be freed after the exchange is closed, so the stream must outlive the exchange
            self._xs = torch.cuda.Stream(device=dev, priority=-1)
            xchg.set_stream(self._xs.cuda_stream)
            self.engine.attach_exchange(xchg)   # the certify kernel now pushes; this side only collects
        return True

    def _comm_device(self):
        """Where the tiny control collectives (decay reference, row counts, set-up handshakes) live: on the GPU under NCCL,
        on the host under gloo (e.g. several ranks sharing one GPU in the tests)."""
        if self.dist.is_initialized() and "nccl" not in str(self.dist.get_backend(self.group)):
            return "cpu"
        return getattr(self.engine, "device", "cpu")

    def refresh_decay(self):
        """All-reduce(MAX) of the kept-row timestamp maximum, then rebuild every shard's decay column."""
        import torch
        mx, cnt = self.engine.kept_ts_max()
        t = torch.tensor([mx if cnt > 0 else float("-inf")], dtype=torch.float64, device=self._comm_device())
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        if t.item() != float("-inf"):
            self.engine.set_decay_reference(t.item())
        return t.item()

    def total_kept(self):
        import torch
        t = torch.tensor([self.engine.n_kept()], dtype=torch.int64, device=self._comm_device())
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    # -- mutation: rows stay contiguous per rank, global ids stay 0..n_total-1 ---------------------------------------
    def _renumber(self):
        """All-gather the local row counts and give every shard its new global offset (exclusive prefix sum)."""
        import torch
        counts = torch.zeros(self.world, dtype=torch.int64, device=self._comm_device())
        counts[self.rank] = self.engine.n_rows()
        if self.world > 1:
            self.dist.all_reduce(counts, group=self.group)
        counts = counts.cpu()
        self.engine.set_row_offset(int(counts[: self.rank].sum()))
        self.n_total = int(counts.sum())
        return self.n_total

    def append(self, rows):
        """Collective.  New documents get the next global ids (hyperdb/hyperdb.py:504-509 appends at the end), so they
        go to the LAST rank's shard; the other ranks pass `rows` too (ignored) or None.  Returns the new total."""
        if self.rank == self.world - 1 and rows is not None and len(rows):
            self.engine.append(rows)
        return self._renumber()

    def remove_rows(self, global_rows):
        """Collective: every rank passes the same GLOBAL row ids (hyperdb/hyperdb.py:718-728); each removes the ones
        it owns, survivors keep their order, and the shards are renumbered.  Returns the new total."""
        import numpy as np
        ids = np.unique(np.asarray(global_rows, dtype=np.int64))
        if len(ids) and (ids[0] < 0 or ids[-1] >= self.n_total):
            raise IndexError("row id out of range")
        off, n = self.engine.row_offset(), self.engine.n_rows()
        mine = ids[(ids >= off) & (ids < off + n)] - off
        if len(mine):
            self.engine.remove_local(mine)
        return self._renumber()

    def query_async(self, queries, top_k, metric, recency_bias=0.0, exact=False):
        """Enqueue sweep -> all-gather -> merge; returns device tensors (idx, score, count, per-shard flags)."""
        import torch
        k = max(int(top_k), 0)
        b = 1 if getattr(queries, "ndim", 1) == 1 else queries.shape[0]
        mine = self.engine.local_topk(queries, k, metric, recency_bias, exact=exact)
        post = getattr(self.engine, "post", None)
        if self.xchg is not None:
            # hdb_query already delivered this rank's records to every rank (fused push); wait + merge run on the exchange's
            # own stream, ordered by the arrival flags.  Batches beyond the exchange buffer fail inside hdb_query.
            w = self.world
            if post is not None:
                mine.record_stream(post)
            out = torch.empty(2 * b * k + b + (w * b + 1) // 2, dtype=torch.int64, device=mine.device)   # [idx | score | count | flags]
            out.record_stream(self._xs)
            self.xchg.collect_async(b, k, out.data_ptr())
            self.exchanges += 1
            return StepResult(out, b, k, w)
        if self.world == 1 and hasattr(self.engine, "m"):
            if post is not None:
                mine.record_stream(post)                 # pipelined: `mine` is completed on the post stream
            return LocalStepResult(mine, b, k, 1)
        if post is not None:
            # pipelined: `mine` is completed on the post stream; exchange and merge follow it there
            mine.record_stream(post)
            with torch.cuda.stream(post):       # allocations made here belong to the post stream
                return self._exchange_and_merge(mine, b, k, post)
        return self._exchange_and_merge(mine, b, k, None)

    def _exchange_and_merge(self, mine, b, k, post):
        import torch
        if self.world > 1:
            gathered = torch.empty(self.world * mine.numel(), dtype=mine.dtype, device=mine.device)
            self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
            self.exchanges += 1
            gathered = gathered.view(self.world, mine.numel())
        else:
            gathered = mine.view(1, -1)
        return self.engine.merge(gathered, b, k)

    def wait_results(self):
        """Make the current stream wait for everything enqueued on the pipelined post stream."""
        post = getattr(self.engine, "post", None)
        if post is None and self.xchg is None:
            return
        import torch
        cur = torch.cuda.current_stream(self.engine.device)
        if post is not None:
            cur.wait_stream(post)
        if self.xchg is not None:
            cur.wait_stream(self._xs)

    # -- asynchronous host API: up to 4 steps in flight, ONE pinned result block per step ---------------------------
    def submit(self, queries, top_k, metric, recency_bias=0.0, _path=0):
        """Enqueue one step (host or CUDA queries) and return a ticket; nothing is synchronised.  Collective: every rank
        submits the same batch in the same order.  Needs the CUDA engine with the peer exchange (or a single rank)."""
        import numpy as np
        eng = self.engine
        if not hasattr(eng, "m") or (self.world > 1 and self.xchg is None):
            raise RuntimeError("submit/collect need the CUDA engine and, on several ranks, the peer-memory exchange")
        from .device_matrix import as_float_array, _NP2HDB, _is_torch
        if _is_torch(queries):
            q = queries.contiguous()
            qdt, space = eng._qdt[q.dtype], (N.HDB_DEVICE if q.is_cuda else N.HDB_HOST)
            ptr, shape = q.data_ptr(), tuple(q.shape)
        else:
            q = as_float_array(queries)
            qdt, space, ptr, shape = _NP2HDB[q.dtype], N.HDB_HOST, q.ctypes.data, q.shape
        b = 1 if len(shape) == 1 else shape[0]
        k = max(int(top_k), 0)
        eng.set_path(_path)
        t = C.c_int64()
        N.check(N.lib().hdb_query_submit(eng.m._h, N.METRIC_IDS[metric], C.c_void_p(ptr), qdt, space, b, k, float(recency_bias),
                                         self.world, C.byref(t)))
        self.exchanges += 1
        return (t.value, b, k, q, metric, float(recency_bias), _path)

    def collect(self, ticket):
        """Wait for one ticket: (idx [B,k], scores [B,k], counts [B]) on the host, identical on every rank.  A step some shard
        could not certify is repeated -- by every rank alike, they all see every shard's flags -- with the wide candidate
        class and then on the exact path."""
        import numpy as np
        tid, b, k, q, metric, bias, path = ticket
        idx = np.empty((b, k), np.int64)
        sc = np.empty((b, k), np.float64)
        cnt = np.empty(b, np.int64)
        flags = np.empty((self.world, b), np.uint32)
        N.check(N.lib().hdb_query_collect(self.engine.m._h, tid, C.c_void_p(idx.ctypes.data), C.c_void_p(sc.ctypes.data),
                                          C.c_void_p(cnt.ctypes.data), C.c_void_p(flags.ctypes.data)))
        if (flags & N.FLAG_EXCHANGE_ERROR).any():
            raise RuntimeError("peer-memory exchange: a rank did not deliver its candidates within 10 s")
        if (flags & N.FLAG_QUERY_NAN).any():
            raise ValueError("Vectors and query_vector should not contain NaN values.")
        bad = np.nonzero((flags & N.FLAG_UNCERTIFIED).any(axis=0))[0] if k > 0 else ()
        if len(bad):
            # only the queries some shard could not certify are repeated (a batch of 4096 typically has a handful)
            nxt = 4 if (path == 0 and k <= 100) else (1 if path != 1 else None)
            if nxt is None:
                raise RuntimeError("the exact path reported an uncertified result")
            if len(bad) == b:
                sub = q
            elif isinstance(q, np.ndarray):
                sub = np.ascontiguousarray(q[bad])
            else:
                sub = q[[int(i) for i in bad]].contiguous()
            try:
                r_idx, r_sc, r_cnt = self.collect(self.submit(sub, k, metric, bias, _path=nxt))
            finally:
                self.engine.set_path(0)
            idx[bad], sc[bad], cnt[bad] = r_idx, r_sc, r_cnt
        return idx, sc, cnt

    def exchange_batch_limit(self, top_k):
        """Largest batch one step can carry through the attached peer exchange for this top_k (None: no exchange, no limit;
        0: not even one query fits).  hdb_query refuses a batch whose message exceeds the buffer (csrc/api.cu)."""
        if self.xchg is None:
            return None
        k = max(int(top_k), 0)
        words = int(self.xchg.max_words)
        b = max(0, (2 * words) // (4 * k + 3))                   # packed_len(b, k) = 2bk + b + ceil(b/2) <= words
        while b > 0 and packed_len(b, k) > words:
            b -= 1
        return b

    def query(self, queries, top_k, metric, recency_bias=0.0):
        """Host results, identical on every rank: (idx [B,k], scores [B,k], counts [B]).  A batch larger than the peer
        exchange's buffer is cut into the largest pieces that fit (every rank cuts alike), three of them in flight."""
        if hasattr(self.engine, "m") and (self.world == 1 or self.xchg is not None):
            shape = tuple(getattr(queries, "shape", ())) or (len(queries),)
            b = 1 if len(shape) == 1 else int(shape[0])
            fit = self.exchange_batch_limit(top_k)
            if fit is not None and b > fit:
                if fit < 1:
                    raise ValueError(f"top_k = {top_k} does not fit the peer exchange buffer (enable_peer_exchange(max_batch, max_k))")
                import numpy as np
                tickets, parts = [], []
                for i in range(0, b, fit):
                    tickets.append(self.submit(queries[i:i + fit], top_k, metric, recency_bias))
                    if len(tickets) == 3:
                        parts.append(self.collect(tickets.pop(0)))
                while tickets:
                    parts.append(self.collect(tickets.pop(0)))
                return tuple(np.concatenate([p[j] for p in parts]) for j in range(3))
            return self.collect(self.submit(queries, top_k, metric, recency_bias))
        idx, sc, cnt, flags = self.query_async(queries, top_k, metric, recency_bias)
        self.wait_results()
        flags = flags.cpu().numpy()
        if (flags & N.FLAG_QUERY_NAN).any():
            raise ValueError("Vectors and query_vector should not contain NaN values.")
        if (flags & N.FLAG_UNCERTIFIED).any():      # every rank sees every flag after the all-gather: same branch everywhere
            idx, sc, cnt, flags = self.query_async(queries, top_k, metric, recency_bias, exact=True)
            self.wait_results()
        b, k = idx.shape
        if idx.data_ptr() + 8 * b * k == sc.data_ptr():          # merged block [idx | score | count]: one copy
            import torch
            blk = torch.as_strided(idx, (2 * b * k + b,), (1,)).cpu()
            return (blk[: b * k].view(b, k).numpy(), blk[b * k: 2 * b * k].view(torch.float64).view(b, k).numpy(),
                    blk[2 * b * k:].numpy())
        return idx.cpu().numpy(), sc.cpu().numpy(), cnt.cpu().numpy()


class GraphedQuery:
    """A captured CUDA graph of one query step (prepare -> sweep/contraction -> certify -> all-gather -> merge) for a
    fixed (batch, top_k, metric, recency_bias): replaying it costs one launch instead of a dozen host calls, which is
    what bounds the latency-dominated configurations (small shards at 8 GPUs, bit-packed hamming).

    The query is copied into a static device buffer before each replay; the outputs are static device tensors that the
    next replay overwrites."""

    def __init__(self, sharded, queries_like, top_k, metric, recency_bias=0.0):
        import torch
        self.sm = sharded
        eng = sharded.engine
        if getattr(eng, "post", None) is not None:
            # pipelined mode spreads one step over internal streams whose cross-step events cannot be captured
            raise RuntimeError("GraphedQuery needs the un-pipelined engine: call enable_pipeline(False) first")
        self.q_static = queries_like.clone()
        side = torch.cuda.Stream(device=eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):
            eng.m.set_stream(side.cuda_stream)
            for _ in range(2):                                   # warm-up: workspaces, NCCL channels, allocator
                sharded.query_async(self.q_static, top_k, metric, recency_bias)
            side.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                eng.m.set_stream(torch.cuda.current_stream(eng.device).cuda_stream)
                self.out = sharded.query_async(self.q_static, top_k, metric, recency_bias)
        eng.m.set_stream(0)
        torch.cuda.current_stream(eng.device).wait_stream(side)

    def replay(self, queries):
        """Enqueue one step on the current stream; returns the static output tensors (idx, score, count, flags)."""
        self.q_static.copy_(queries, non_blocking=True)
        self.graph.replay()
        return self.out
