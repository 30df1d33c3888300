"""Row-sharded collection over torch.distributed (one process per GPU).

The reference is single-process (no collectives anywhere); this is the
partitioning BASELINE.json's north_star asks for.  Every rank owns a shard of the rows and a
full ``B200Collection`` over it.  All public methods are *collective*: every rank calls them
with the same arguments and gets the same result.

  add/upsert   new rows of a batch are split into contiguous slices, one per rank; every row
               gets a global insertion sequence number (the oracle's tie-break key).
  query        queries are replicated; each rank answers from its shard with the exact engine
               (fp64 distances kept), the per-rank [nq, k] candidate lists are exchanged with ONE
               all_gather (NCCL over NVLink on GPUs; 12-20 B per candidate), and every rank runs
               the same merge on (fp64 distance, global sequence) -- the exchange is the only
               data-path collective because top-k is a decomposable reduction.
  where        each shard evaluates the clause on its own metadata tables (no global mask).

``shard_factory`` / ``merge_fn`` exist so the host-side logic can be exercised with gloo on a
machine without a GPU (tests inject CPU stand-ins); the defaults are the CUDA engine and the
CUDA merge kernel, and there is no CPU fallback in the product path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def cuda_merge(rows, d64, cnt, k):
    """[R,nq,k] int64 / fp64, [R,nq] int32 CUDA tensors -> merged (rows, dist fp32, count)."""
    R, nq, _ = rows.shape
    out_rows = torch.empty((nq, k), dtype=torch.int64, device=rows.device)
    out_dist = torch.empty((nq, k), dtype=torch.float32, device=rows.device)
    out_cnt = torch.empty((nq,), dtype=torch.int32, device=rows.device)
    lib = _lib.load()
    _lib.check(lib.b2r_merge_shards(rows.data_ptr(), d64.data_ptr(), cnt.data_ptr(), R, nq, k,
                                    out_rows.data_ptr(), out_dist.data_ptr(), out_cnt.data_ptr(),
                                    rows.device.index or 0,
                                    torch.cuda.current_stream(rows.device).cuda_stream), "b2r_merge_shards")
    return out_rows, out_dist, out_cnt


class ShardedCollection:
    def __init__(self, name="multimodal_rag", metadata=None, *, group=None, device=None,
                 shard_factory=None, merge_fn=None, **shard_kw):
        if not dist.is_initialized():
            raise RuntimeError("ShardedCollection needs an initialised torch.distributed process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.name, self.metadata = name, metadata
        if shard_factory is None:
            from .collection import B200Collection
            if device is None:
                device = torch.cuda.current_device()
            shard_factory = lambda: B200Collection(name, metadata, device=device, **shard_kw)
            self.comm_device = torch.device("cuda", device)
        else:
            self.comm_device = torch.device("cpu") if device is None else torch.device(device)
        self.shard = shard_factory()
        self.merge_fn = merge_fn or cuda_merge
        self._next_seq = 0
        self._gseq = np.zeros(0, dtype=np.int64)          # local row -> global sequence number
        self._gseq_dev = None

    # ---- helpers -------------------------------------------------------------------
    def _all_gather_obj(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def _slice(self, m):
        """contiguous, balanced slice of m new rows owned by this rank"""
        lo = (m * self.rank) // self.world
        hi = (m * (self.rank + 1)) // self.world
        return lo, hi

    def _ingest(self, ids, embeddings, metadatas, documents, upsert):
        n = len(ids)
        # the WHOLE batch is validated on EVERY rank before any rank-local work: a bad batch raises everywhere, so no
        # rank is left waiting in the next collective for one that bailed out
        if len(set(ids)) != n:
            raise ValueError("Expected IDs to be unique")
        if not all(isinstance(i, str) and i for i in ids):
            raise ValueError("Expected IDs to be non-empty strings")
        emb = embeddings if isinstance(embeddings, (np.ndarray, torch.Tensor)) else np.asarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or len(emb) != n:
            raise ValueError(f"Number of embeddings {len(emb)} must match number of ids {n}")
        dim = getattr(self.shard, "dimension", None)
        if dim is not None and emb.shape[1] != dim:
            raise ValueError(f"Embedding dimension {emb.shape[1]} does not match collection dimensionality {dim}")
        for name, lst in (("metadatas", metadatas), ("documents", documents)):
            if lst is not None and len(lst) != n:
                raise ValueError(f"Number of {name} {len(lst)} must match number of ids {n}")
        if metadatas is not None:
            from .where import MetaTable
            MetaTable.validate_batch(metadatas)
        have_local = self._held(ids)
        have = set().union(*self._all_gather_obj(have_local))
        if upsert:
            new = list(range(n))
        else:
            new = [j for j, i in enumerate(ids) if i not in have]
        lo, hi = self._slice(len(new))
        mine = new[lo:hi]
        if upsert and have_local:
            # the old versions go only after the batch was accepted everywhere (validated above); an id keeps living on
            # the rank that first stored it unless this rank's slice of the batch carries it again
            stay = set(ids[j] for j in mine)
            gone = [i for i in have_local if i not in stay]
            if gone:
                self.shard.delete(ids=gone)
        if mine:
            if isinstance(emb, torch.Tensor):
                e = emb[torch.as_tensor(mine, device=emb.device)]
            else:
                e = emb[np.asarray(mine)]
            (self.shard.upsert if upsert else self.shard.add)(ids=[ids[j] for j in mine], embeddings=e,
                           metadatas=None if metadatas is None else [metadatas[j] for j in mine],
                           documents=None if documents is None else [documents[j] for j in mine])
            self._gseq = np.concatenate([self._gseq, self._next_seq + np.asarray(range(lo, hi), dtype=np.int64)])
            self._gseq_dev = None
        self._next_seq += len(new)

    def _held(self, ids) -> list:
        """the ids of the batch this rank's shard holds (live)"""
        if not len(ids):
            return []
        found = self.shard.rows_of(ids)
        return [i for i, r in zip(ids, found.tolist()) if r >= 0]

    def add(self, ids, embeddings=None, metadatas=None, documents=None):
        self._ingest(list(ids), embeddings, metadatas, documents, upsert=False)

    def upsert(self, ids, embeddings=None, metadatas=None, documents=None):
        self._ingest(list(ids), embeddings, metadatas, documents, upsert=True)

    def delete(self, ids=None, where=None):
        if (ids is None or len(ids) == 0) and not where:
            raise ValueError("You must provide either ids, where, or where_document to delete.")
        if ids is not None:
            ids = self._held(ids)
            if not ids and where is None:
                return
        self.shard.delete(ids=ids, where=where)

    def count(self) -> int:
        t = torch.tensor([self.shard.count()], dtype=torch.int64, device=self.comm_device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    # ---- query ---------------------------------------------------------------------
    def query_rows(self, query_embeddings, n_results=10, where=None):
        """Collective.  Returns (gseq [nq,k] int64 global sequence numbers, dist [nq,k] fp32,
        count [nq] int32) as numpy arrays, identical on every rank."""
        k = n_results
        rows, _, cnt, d64 = self.shard.query_rows(query_embeddings, k, where, want_dist64=True)
        g = np.where(rows >= 0, self._gseq[np.clip(rows, 0, max(len(self._gseq) - 1, 0))] if len(self._gseq) else -1, -1)
        dev = self.comm_device
        t_rows = torch.from_numpy(np.ascontiguousarray(g, dtype=np.int64)).to(dev)
        t_d64 = torch.from_numpy(d64).to(dev)
        t_cnt = torch.from_numpy(cnt).to(dev)
        R = self.world
        a_rows = torch.empty((R,) + tuple(t_rows.shape), dtype=torch.int64, device=dev)
        a_d64 = torch.empty((R,) + tuple(t_d64.shape), dtype=torch.float64, device=dev)
        a_cnt = torch.empty((R,) + tuple(t_cnt.shape), dtype=torch.int32, device=dev)
        # list form of all_gather: identical on nccl and gloo (gloo's *_into_tensor is shape-picky)
        dist.all_gather(list(a_rows.unbind(0)), t_rows, group=self.group)
        dist.all_gather(list(a_d64.unbind(0)), t_d64, group=self.group)
        dist.all_gather(list(a_cnt.unbind(0)), t_cnt, group=self.group)
        o_rows, o_dist, o_cnt = self.merge_fn(a_rows, a_d64, a_cnt, k)
        return o_rows.cpu().numpy(), o_dist.cpu().numpy(), o_cnt.cpu().numpy()

    def query(self, query_embeddings=None, n_results=10, where=None,
              include=("metadatas", "documents", "distances")):
        """Collective ``Collection.query`` with Chroma's nested-list result on every rank."""
        g, d, cnt = self.query_rows(query_embeddings, n_results, where)
        nq = g.shape[0]
        # winners owned by this rank -> payload; one object all_gather assembles the rest
        mine = {}
        if len(self._gseq):
            seqs = np.unique(np.concatenate([g[i, : cnt[i]] for i in range(nq)])) if nq else np.empty(0, dtype=np.int64)
            pos = np.minimum(np.searchsorted(self._gseq, seqs), len(self._gseq) - 1)
            own = self._gseq[pos] == seqs
            seqs, pos = seqs[own].tolist(), pos[own]
            meta, docs = self.shard._meta.meta, self.shard._docs
            for s, p, id_ in zip(seqs, pos.tolist(), self.shard.ids_of(pos)):      # one id-table call for all winners
                mine[s] = (id_, meta[p], docs[p])
        table = {}
        for part in self._all_gather_obj(mine):
            table.update(part)
        res = {"ids": [], "distances": None, "metadatas": None, "documents": None, "embeddings": None}
        for key in include:
            res[key] = []
        for i in range(nq):
            seqs = g[i, : cnt[i]].tolist()
            res["ids"].append([table[s][0] for s in seqs])
            if res["distances"] is not None:
                res["distances"].append(d[i, : cnt[i]].tolist())
            if res["metadatas"] is not None:
                res["metadatas"].append([table[s][1] for s in seqs])
            if res["documents"] is not None:
                res["documents"].append([table[s][2] for s in seqs])
        return res


class DeviceShard:
    """Minimal device-resident shard for throughput runs (bench.py): no ids, no metadata, rows
    numbered row_base + local.  query_device keeps every tensor on the GPU and enqueues
    local scan -> all_gather -> merge on the current stream without a host sync."""

    exchange_mode = "nccl all_gather_into_tensor + merge kernel, same stream"
    _P2P_MODE = ("b2r_xchg_push / b2r_xchg_merge: peer-to-peer stores into the peers' mailboxes over NVLink, stream "
                 "memory operations as the only waits, merge kernel; no NCCL on the data path")

    def __init__(self, dim, space="cosine", *, capacity=0, row_base=0, device=None, group=None,
                 keep_f32_master=True, world=None):
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else device
        self.group = group
        # world=1: a stand-alone shard inside a distributed job (no exchange)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        h = ctypes.c_void_p()
        _lib.check(self.lib.b2r_create(dim, _lib.SPACE_CODE[space], capacity, self.device,
                                       0 if keep_f32_master else _lib.FLAG_NO_F32_MASTER, ctypes.byref(h)))
        self.h = h
        self.dim = dim
        self._side = None           # stream of the pipelined exchange
        self._inflight = None
        self._xchg = None           # peer-to-peer exchange (enable_p2p_exchange)
        self._fused_prev = None     # (nq, k, outputs) of the fused batch whose merge has not been enqueued yet
        _lib.check(self.lib.b2r_set_row_base(h, row_base))

    def close(self):
        if self._xchg is not None:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)          # no peer may still be writing into this rank's mailbox
            self.lib.b2r_xchg_destroy(self._xchg)
            self._xchg = None
        if self.h is not None:
            self.lib.b2r_destroy(self.h)
            self.h = None

    def fallbacks(self) -> int:
        st = _lib.B2RStats()
        _lib.check(self.lib.b2r_get_stats(self.h, ctypes.byref(st)))
        return int(st.n_exact_fallbacks)

    def ingest(self, x: torch.Tensor):
        first = ctypes.c_int64()
        x = x.contiguous()
        _lib.check(self.lib.b2r_ingest_f32(self.h, x.data_ptr(), x.shape[0], None, ctypes.byref(first),
                                           torch.cuda.current_stream().cuda_stream), "b2r_ingest_f32")
        return first.value

    def alloc_out(self, nq, k):
        """Outputs of one batch.  rows / d64 / cnt are views of ONE flat allocation ('pack'), so the
        cross-shard exchange is a single all_gather of that block."""
        dev = torch.device("cuda", self.device)
        R = self.world
        off_rows, off_d64, off_cnt = 0, nq * k * 8, 2 * nq * k * 8
        stride = (off_cnt + nq * 4 + 15) // 16 * 16
        pack = torch.zeros((stride,), dtype=torch.uint8, device=dev)
        return {
            "pack": pack, "layout": (stride, off_rows, off_d64, off_cnt),
            "rows": pack[off_rows:off_d64].view(torch.int64).view(nq, k),
            "d64": pack[off_d64:off_cnt].view(torch.float64).view(nq, k),
            "cnt": pack[off_cnt:off_cnt + nq * 4].view(torch.int32),
            "dist": torch.empty((nq, k), dtype=torch.float32, device=dev),
            "a_pack": torch.empty((R, stride), dtype=torch.uint8, device=dev),
            "m_rows": torch.empty((nq, k), dtype=torch.int64, device=dev),
            "m_dist": torch.empty((nq, k), dtype=torch.float32, device=dev),
            "m_cnt": torch.empty((nq,), dtype=torch.int32, device=dev),
        }

    def alloc_host(self, nq, k):
        """Pinned host mirrors for query_host: the query batch staging and the merged results."""
        return {"rows": torch.empty((nq, k), dtype=torch.int64).pin_memory(),
                "dist": torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                "cnt": torch.empty((nq,), dtype=torch.int32).pin_memory(),
                "q_dev": torch.empty((nq, self.dim), dtype=torch.float32, device=torch.device("cuda", self.device))}

    def query_host(self, q_host: torch.Tensor, k: int, o: dict, h: dict):
        """End to end with HOST buffers: pinned query batch -> device, local exact top-k, exchange + merge, merged
        rows / distances / counts -> pinned host arrays, stream synchronised on return."""
        h["q_dev"].copy_(q_host, non_blocking=True)
        rows, dist_, cnt = self.query_device(h["q_dev"], k, o)
        h["rows"].copy_(rows, non_blocking=True)
        h["dist"].copy_(dist_, non_blocking=True)
        h["cnt"].copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h["rows"], h["dist"], h["cnt"]

    def query_local(self, q: torch.Tensor, k: int, o: dict):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.lib.b2r_query_ex(self.h, q.data_ptr(), q.shape[0], k, None, o["rows"].data_ptr(),
                                         o["dist"].data_ptr(), o["d64"].data_ptr(), o["cnt"].data_ptr(), st),
                   "b2r_query")

    def query_device(self, q: torch.Tensor, k: int, o: dict):
        """Replicated queries against the row-sharded corpus: local exact top-k, one all_gather of
        the candidate lists over NCCL, merge kernel.  Results in o['m_rows'|'m_dist'|'m_cnt']."""
        self.query_local(q, k, o)
        if self.world == 1:
            return o["rows"], o["dist"], o["cnt"]
        self._exchange(q.shape[0], k, o)
        return o["m_rows"], o["m_dist"], o["m_cnt"]

    def enable_p2p_exchange(self, nq_max=1024, k_max=128, default=True):
        """Replace the NCCL all_gather + merge launch by the library's own exchange kernel (b2r_xchg_*): every rank's
        lists go straight into the peers' mailboxes over NVLink and the same kernel merges.  Collective (every rank
        calls it); needs all ranks on one node with peer access.  torch.distributed is only used here, for the one-off
        hand-over of the IPC handles."""
        if self.world == 1:
            return
        if self._xchg is not None:
            if default and not self._xchg_default:
                self._xchg_default = True
                self.exchange_mode = self._P2P_MODE
            return
        # Every rank walks the same collectives whatever fails locally (a rank that raised half way would leave the others
        # waiting in a different collective): failures are agreed on first, then raised on every rank.
        rank = dist.get_rank(self.group)
        x, err = ctypes.c_void_p(), None
        buf = ctypes.create_string_buffer(64)
        try:
            _lib.check(self.lib.b2r_xchg_create(self.device, rank, self.world, nq_max, k_max, ctypes.byref(x)), "b2r_xchg_create")
            _lib.check(self.lib.b2r_xchg_ipc_handle(x, buf), "b2r_xchg_ipc_handle")
        except Exception as e:                       # noqa: BLE001 -- re-raised below, on every rank
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(buf.raw) if err is None else b"", group=self.group)
        if err is None and all(len(h_) == 64 for h_ in handles):
            try:
                _lib.check(self.lib.b2r_xchg_open(x, b"".join(handles)), "b2r_xchg_open")
            except Exception as e:                   # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer could not create its mailbox")
        dev = torch.device("cuda", self.device)
        ok = torch.tensor([0 if err is not None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) != 1:
            if x:
                self.lib.b2r_xchg_destroy(x)
            raise RuntimeError(f"peer-to-peer exchange unavailable on this node: {err or 'a peer failed to map the mailboxes'}")
        dist.barrier(group=self.group)
        self._xchg = x
        self._xchg_limits = (nq_max, k_max)
        self._xchg_default = default        # False: the mailboxes serve query_device_fused only, query_device keeps NCCL
        if not default:
            return
        self.exchange_mode = self._P2P_MODE

    def query_device_fused(self, q: torch.Tensor, k: int, o: dict):
        """Replicated queries, the exchange fused into the query's own kernels (b2r_query_push): the finalize stores each
        list into every rank's mailbox over NVLink as it emits it, and the merge of the PREVIOUS batch rides in this batch's
        last kernel -- by then its lists have arrived, so nothing waits and no collective, exchange kernel or extra launch stands
        between two scans.  Needs enable_p2p_exchange().  Results of this batch (o['m_rows'|'m_dist'|'m_cnt']) are in place
        after the next call or after drain(); `o` must not be reused before then (alternate two output sets)."""
        nq = q.shape[0]
        if self.world == 1:
            self.query_local(q, k, o)
            return
        if not self._uses_xchg(nq, k, fused=True):
            raise ValueError("query_device_fused: enable_p2p_exchange(nq_max, k_max) must cover this batch")
        st = torch.cuda.current_stream().cuda_stream
        prev, self._fused_prev = self._fused_prev, (nq, k, o)
        m = prev[2] if prev is not None else None       # the previous batch is merged inside this call's last kernel
        _lib.check(self.lib.b2r_query_push(self.h, self._xchg, q.data_ptr(), nq, k, None, o["rows"].data_ptr(),
                                           o["dist"].data_ptr(), o["cnt"].data_ptr(),
                                           m["m_rows"].data_ptr() if m else None, m["m_dist"].data_ptr() if m else None,
                                           m["m_cnt"].data_ptr() if m else None, st), "b2r_query_push")

    def _uses_xchg(self, nq, k, fused=False):
        return self._xchg is not None and (fused or self._xchg_default) and nq <= self._xchg_limits[0] and k <= self._xchg_limits[1]

    def _xchg_push(self, nq, k, o):
        _lib.check(self.lib.b2r_xchg_push(self._xchg, o["rows"].data_ptr(), o["d64"].data_ptr(), o["cnt"].data_ptr(), nq, k,
                                          torch.cuda.current_stream().cuda_stream), "b2r_xchg_push")

    def _xchg_merge(self, nq, k, o):
        _lib.check(self.lib.b2r_xchg_merge(self._xchg, nq, k, o["m_rows"].data_ptr(), o["m_dist"].data_ptr(), o["m_cnt"].data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "b2r_xchg_merge")

    def _exchange(self, nq, k, o):
        if self._uses_xchg(nq, k):
            self._xchg_push(nq, k, o)
            self._xchg_merge(nq, k, o)
            return
        dist.all_gather_into_tensor(o["a_pack"], o["pack"], group=self.group)      # the only data-path collective
        st = torch.cuda.current_stream().cuda_stream
        stride, off_rows, off_d64, off_cnt = o["layout"]
        _lib.check(self.lib.b2r_merge_shards_packed(o["a_pack"].data_ptr(), stride, off_rows, off_d64, off_cnt,
                                                    self.world, nq, k, o["m_rows"].data_ptr(), o["m_dist"].data_ptr(),
                                                    o["m_cnt"].data_ptr(), self.device, st), "b2r_merge_shards_packed")

    PIPELINE_MAX_BYTES = 64 << 10      # per-rank message up to which the exchange is moved to the side stream

    def query_device_pipelined(self, q: torch.Tensor, k: int, o: dict):
        """The same with the exchange taken off the scoring stream: the local top-k of this batch is enqueued on the
        current stream, its all_gather + merge on a side stream behind an event, so the next batch's scan starts while
        this batch's lists travel.  `o` must not be reused before the batch after next (two output sets, alternated);
        results are valid once o['ev_done'] has been reached -- `drain()` makes the current stream wait for every
        exchange still in flight."""
        cur = torch.cuda.current_stream()
        if "ev_done" in o:
            cur.wait_event(o["ev_done"])          # the previous user of these buffers has been merged
        if self.world > 1 and not self._uses_xchg(q.shape[0], k) and o["layout"][0] > self.PIPELINE_MAX_BYTES:
            # large lists (config 4: 1.6 MB per rank): NCCL's all_gather then wants many SMs at once and fights the next
            # batch's scan for them -- measured 10 ms per step instead of 0.9 -- so the exchange stays on the scan's stream,
            # where it costs 40 us of a multi-millisecond step
            self.query_device(q, k, o)
            if "ev_local" not in o:
                o["ev_local"], o["ev_done"] = torch.cuda.Event(), torch.cuda.Event()
            o["ev_done"].record(cur)
            self._inflight = o
            return
        self.query_local(q, k, o)
        if self.world == 1:
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        if "ev_local" not in o:
            o["ev_local"], o["ev_done"] = torch.cuda.Event(), torch.cuda.Event()
        nq = q.shape[0]
        if self._uses_xchg(nq, k):
            self._xchg_push(nq, k, o)              # on the scan's stream: a few microseconds, nothing waits inside it
            with torch.cuda.stream(self._side):    # the merge's stream waits on the mailbox words (this rank's push included)
                self._xchg_merge(nq, k, o)
                o["ev_done"].record(self._side)
        else:
            o["ev_local"].record(cur)
            self._side.wait_event(o["ev_local"])
            with torch.cuda.stream(self._side):
                self._exchange(nq, k, o)
                o["ev_done"].record(self._side)
        self._inflight = o

    def drain(self):
        if self._fused_prev is not None:           # the last fused batch has no successor to ride behind
            prev, self._fused_prev = self._fused_prev, None
            self._xchg_merge(*prev)
        if self._inflight is not None:
            torch.cuda.current_stream().wait_event(self._inflight["ev_done"])
            self._inflight = None
