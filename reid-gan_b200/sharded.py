"""Row-partitioned pseudo-label pass: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch) for the exchange steps.  The reference has no multi-GPU hot path (its only sharding is
faiss.IndexShards for search_option=2, utils/faiss_utils.py:92-106); the path is row-separable
(SURVEY.md section 8e), so rank r owns the contiguous rows [N*r/W, N*(r+1)/W) of every per-row stage:

    features   : all-gather of the fp32 rows (skipped when every rank already holds all N rows)
    a1  kNN    : own query rows against all N columns            -> all-gather neighbour lists (N x k1 int32)
    a2-a4      : own rows (R_half masks for all rows are recomputed locally: cheaper than a gather)
                                                                 -> all-gather V rows (CSR)
    a5  V_qe   : own rows                                        -> all-gather V_qe rows (CSR)
    a6  index  : replicated (tiny)
    a7  graph  : own rows of the eps-graph                       -> all-gather neighbour lists
    a8  labels : replicated union-find on the global graph (N x ~30 edges)

Outputs are byte-identical for every world size: each row's arithmetic is independent of the
partition and all accumulation orders are fixed.
The helpers only use torch.distributed collectives on whatever device the tensors live on, so the
host logic is testable on CPU with the gloo backend (tests/test_sharded_cpu.py).
"""
import torch
import torch.distributed as dist


def partition(N, world, rank):
    """Contiguous row block of `rank`: [N*rank//world, N*(rank+1)//world)."""
    return (N * rank) // world, (N * (rank + 1)) // world


class RowComm:
    """All-gathers of row-sharded arrays (fixed-width rows and CSR pieces) over a process group."""

    def __init__(self, N, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.N = N
        self.bounds = [partition(N, self.world, r) for r in range(self.world)]
        self.r0, self.r1 = self.bounds[self.rank]
        self.max_rows = max(b - a for a, b in self.bounds)
        self._into_tensor = dist.get_backend(group) == "nccl"

    def _all_gather_padded(self, t, length, max_len):
        """t: (>=length, ...) local; returns a (world, max_len, ...) tensor after a padded all-gather."""
        shape = (max_len,) + tuple(t.shape[1:])
        if t.shape[0] == max_len and t.is_contiguous():
            send = t
        else:
            send = torch.empty(shape, dtype=t.dtype, device=t.device)
            if length:
                send[:length] = t[:length]
            if length < max_len:
                send[length:].zero_()
        out = torch.empty((self.world,) + shape, dtype=t.dtype, device=t.device)
        if self._into_tensor:
            dist.all_gather_into_tensor(out, send, group=self.group)
        else:
            dist.all_gather(list(out.unbind(0)), send, group=self.group)
        return out

    def gather_rows(self, t):
        """(n_local, ...) -> (N, ...): fixed-width rows in rank order."""
        n = self.r1 - self.r0
        out = self._all_gather_padded(t, n, self.max_rows)
        if self.N == self.max_rows * self.world:            # even split: the gathered buffer IS the result
            return out.reshape((self.N,) + tuple(t.shape[1:]))
        return torch.cat([out[r, : b - a] for r, (a, b) in enumerate(self.bounds)], dim=0)

    def gather_lengths(self, length, dev):
        """One int64 per rank (tensor all-gather + a single host read-back)."""
        mine = torch.tensor([int(length)], dtype=torch.int64, device=dev)
        out = self._all_gather_padded(mine, 1, 1)
        return [int(v) for v in out.flatten().tolist()]

    def gather_concat(self, t, length, lens=None):
        """1-D local array with `length` valid entries -> concatenation over ranks, and the lengths."""
        if lens is None:
            lens = self.gather_lengths(length, t.device)
        out = self._all_gather_padded(t, length, max(max(lens), 1))
        return torch.cat([out[r, :l] for r, l in enumerate(lens)], dim=0), lens

    def gather_csr(self, cnt, idx, val):
        """Local CSR pieces (row counts, column indices, values) -> global (ptr, idx, val, nnz, max_row_nnz)."""
        from .faiss_rerank import _scan
        g_cnt = self.gather_rows(cnt)
        if g_cnt.is_cuda:
            g_ptr, total, mx = _scan(g_cnt, self.N, g_cnt.device)
            edges = g_ptr[[a for a, _ in self.bounds] + [self.N]].tolist()      # per-rank nnz without a 2nd collective
        else:                                   # CPU (gloo) test path: same arithmetic in torch
            g_ptr = torch.zeros(self.N + 1, dtype=torch.int64)
            g_ptr[1:] = torch.cumsum(g_cnt.to(torch.int64), 0)
            total, mx = int(g_ptr[-1]), int(g_cnt.max()) if self.N else 0
            edges = [int(g_ptr[a]) for a, _ in self.bounds] + [total]
        lens = [edges[r + 1] - edges[r] for r in range(self.world)]
        g_idx, _ = self.gather_concat(idx, lens[self.rank], lens)
        g_val, _ = self.gather_concat(val, lens[self.rank], lens)
        if g_idx.numel() == 0:
            g_idx = torch.zeros(1, dtype=idx.dtype, device=idx.device)
            g_val = torch.zeros(1, dtype=val.dtype, device=val.device)
        return g_ptr, g_idx, g_val, total, mx

    def gather_neighbors(self, slot_ptr, nbr_idx, nbr_cnt):
        """Per-row neighbour lists stored at slot_ptr (upper-bound slots) -> compact global lists:
        (ptr int64 (N+1), idx, cnt int32 (N))."""
        n = self.r1 - self.r0
        g_cnt = self.gather_rows(nbr_cnt[:n])
        dev = nbr_idx.device
        if g_cnt.is_cuda:
            from .faiss_rerank import _scan
            from ._lib import call, ptr as p_, stream_ptr
            g_ptr, total_all, _ = _scan(g_cnt, self.N, dev)
            edges = g_ptr[[a for a, _ in self.bounds] + [self.N]].tolist()
            lens = [edges[r + 1] - edges[r] for r in range(self.world)]
            compact = torch.empty(max(lens[self.rank], 1), dtype=torch.int32, device=dev)
            loc_ptr = (g_ptr[self.r0:self.r1 + 1] - edges[self.rank]).contiguous()
            call("reid_lists_compact", p_(slot_ptr), p_(nbr_idx), p_(nbr_cnt), p_(loc_ptr), n, p_(compact), stream_ptr())
        else:
            cnt64 = nbr_cnt[:n].to(torch.int64)
            total = int(cnt64.sum())
            if total:
                row_of = torch.repeat_interleave(torch.arange(n), cnt64)
                first = torch.cumsum(cnt64, 0) - cnt64
                pos = slot_ptr[:n][row_of] + (torch.arange(total) - first[row_of])
                compact = nbr_idx[pos]
            else:
                compact = nbr_idx[:0]
            g_ptr = torch.zeros(self.N + 1, dtype=torch.int64)
            g_ptr[1:] = torch.cumsum(g_cnt.to(torch.int64), 0)
            edges = [int(g_ptr[a]) for a, _ in self.bounds] + [int(g_ptr[-1])]
            lens = [edges[r + 1] - edges[r] for r in range(self.world)]
        g_idx, _ = self.gather_concat(compact, lens[self.rank], lens)
        if g_idx.numel() == 0:
            g_idx = torch.zeros(1, dtype=nbr_idx.dtype, device=dev)
        return g_ptr, g_idx, g_cnt.contiguous()


@torch.no_grad()
def pseudo_labels(x, k1=30, k2=6, eps=0.6, min_samples=4, knn="auto", centroids=False, group=None, N=None,
                  timers=False):
    """Row-sharded pass.  `x`: either all N rows (every rank holds a replica) or this rank's row block
    (then N must be given and the blocks are all-gathered first).  Returns the same dict as
    pipeline.pseudo_labels with GLOBAL labels on every rank."""
    from ._lib import call, ptr, stream_ptr
    from .dbscan import dbscan_from_neighbors
    from .faiss_rerank import jaccard_neighbors, rerank_state
    if not x.is_cuda:
        raise RuntimeError("sharded.pseudo_labels needs CUDA tensors; there is no CPU fallback")
    with torch.cuda.device(x.device):
        if N is None or x.shape[0] == N:
            N = x.shape[0]
            comm = RowComm(N, group)
        else:
            comm = RowComm(N, group)
            if x.shape[0] != comm.r1 - comm.r0:
                raise ValueError("rank %d holds %d rows, expected %d" % (comm.rank, x.shape[0], comm.r1 - comm.r0))
            x = comm.gather_rows(x.contiguous())                    # collective (1): features
        st = rerank_state(x.contiguous(), k1, k2, knn=knn, comm=comm, timers=timers)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timers else None
        if timers:
            ev[0].record()
        slot_ptr, nbr_idx, nbr_cnt, _ = jaccard_neighbors(st, eps)
        if timers:
            ev[1].record()
        g_ptr, g_idx, g_cnt = comm.gather_neighbors(slot_ptr, nbr_idx, nbr_cnt)
        if timers:
            ev[2].record()
        labels, core, ncl = dbscan_from_neighbors(N, g_ptr, g_idx, g_cnt, min_samples)
        if timers:
            ev[3].record()
            torch.cuda.synchronize()
            for nm, a, b in (("jaccard", 0, 1), ("gather_neighbors", 1, 2), ("dbscan", 2, 3)):
                st.timings[nm] = ev[a].elapsed_time(ev[b]) * 1e-3
        out = dict(labels=labels, core=core, num_clusters=ncl, state=st, nbr_cnt=nbr_cnt)
        if centroids:
            C = int(ncl.item())
            cen = torch.empty((C, x.shape[1]), dtype=torch.float32, device=x.device)
            if C:
                call("reid_centroids", ptr(x), N, x.shape[1], ptr(labels), C, 1, ptr(cen), None, stream_ptr())
            out["centroids"] = cen
        return out


@torch.no_grad()
def pseudo_labels_host(x_host, k1=30, k2=6, eps=0.6, min_samples=4, knn="auto", group=None):
    """End-to-end variant: `x_host` = all N rows in (pinned) host memory on every rank; each rank uploads
    only ITS row block, the blocks are all-gathered over NVLink, labels come back to the host."""
    N = x_host.shape[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    r0, r1 = partition(N, world, rank)
    dev = torch.device("cuda", torch.cuda.current_device())
    x_local = x_host[r0:r1].to(dev, non_blocking=True)
    out = pseudo_labels(x_local, k1, k2, eps, min_samples, knn=knn, group=group, N=N)
    return out["labels"].cpu().numpy()
