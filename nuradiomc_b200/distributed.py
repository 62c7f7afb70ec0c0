"""
Multi-GPU layer: one process per GPU (`torch.distributed`), pairs sharded statically, results gathered over NVLink.

Pairs are independent: every rank traces its contiguous block of vertices against all antennas, no
collective on the data path.  The only exchange is the gather of the compact (per-solution, CSR) results on one rank -- what
the reference does at file level when it merges the outputs of its per-job runs (NuRadioMC/utilities/runner.py:42-99,
NuRadioMC/utilities/merge_hdf5.py).  Two variants, same result on the gathering rank:

* `P2PGather` ("p2p", the product): the gathering rank allocates the result arrays and exports them through CUDA IPC; every
  rank maps them and passes the mapped base pointers, its own row segment (`row_base`) and its own slice of the per-pair
  arrays to the device-resident trace.  The solver and attenuation kernels then store every result row straight into the
  gathering GPU's HBM through NVLink while they compute: no staging buffer, no row counts on the host, no collective call;
  the transfer overlaps the arithmetic row by row.  Layout on the gathering rank: pairs in global order; the rows of rank r
  live in the segment [row_base[r], row_base[r] + n_rows[r]) (segments are sized for the worst case, so the CSR is addressed
  by sol_offset[i] and n_sol[i]; `compacted()` closes the gaps when a consumer wants sol_offset[i+1] - sol_offset[i]).
* `gather_rows` ("nccl", the baseline to compare with): every rank traces into its own HBM, the row counts are exchanged
  (one tiny all_gather + host read), then one grouped ncclSend/ncclRecv per array (`batch_isend_irecv`) moves exactly the
  filled rows; optionally pipelined over chunks of vertices so that the transfer of one chunk overlaps the kernels of the
  next (`ChunkedNcclGather`).  Works on any backend (the CPU tests run it on gloo).
"""
import ctypes as C

import numpy as np

PER_PAIR = ("n_sol", "status", "sol_offset")


def shard_bounds(n_items, world_size, rank):
    """contiguous block [lo, hi) of `n_items` for `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_vertices(n_vertices, world_size, rank, permutation_seed=None):
    """indices of the vertices of `rank`.  With `permutation_seed` the vertices are shuffled first (same permutation on
    every rank) so that geometric clustering of cheap shadow-zone pairs cannot unbalance the ranks."""
    lo, hi = shard_bounds(n_vertices, world_size, rank)
    if permutation_seed is None:
        return np.arange(lo, hi)
    perm = np.random.default_rng(permutation_seed).permutation(n_vertices)
    return np.sort(perm[lo:hi])


def bind_to_gpu_numa_node(device_index):
    """
    Bind this process (CPU affinity and, through first touch, its pinned host buffers) to the NUMA node the GPU hangs off:
    eight ranks that all run on node 0 push every D2H byte of the far GPUs through the socket interconnect.  Best effort
    (sysfs + sched_setaffinity, no libnuma): returns a dict describing what was done; never raises.
    """
    import os
    info = {"device": int(device_index), "numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        info["cpus"] = len(allowed)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as e:        # containers without sysfs access, cpusets that exclude the node, ...
        info["error"] = repr(e)
    return info


# ------------------------------------------------------------------------------------------------------------------
# baseline variant: NCCL (or gloo) send/recv of the filled rows
# ------------------------------------------------------------------------------------------------------------------
def exchange_counts(n_local, device, group=None):
    """list of one integer per rank (one small all_gather and a host read)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return [int(p.item()) for p in parts]


def gather_rows(local, n_local, dst=0, group=None, out=None, counts=None):
    """
    Gather the first `n_local` rows of every tensor of the dict `local` on rank `dst`, ranks in order (gatherv): exactly the
    filled rows travel, as one grouped send/recv per tensor (ncclGroupStart / ncclSend / ncclRecv under `batch_isend_irecv`).
    Returns (dict of concatenated tensors, counts) on `dst` and (None, counts) elsewhere.  `out`: optional dict of
    preallocated destination tensors (at least sum(counts) rows).  `counts`: row counts per rank if already known.
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    keys = sorted(local.keys())
    dev = local[keys[0]].device
    if counts is None:
        counts = exchange_counts(n_local, dev, group)
    starts = np.concatenate([[0], np.cumsum(counts)])
    total = int(starts[-1])
    ops, res = [], None
    if rank == dst:
        res = {}
        for k in keys:
            t = local[k]
            if out is not None and k in out:
                full = out[k]
                assert full.shape[0] >= total and full.shape[1:] == t.shape[1:] and full.dtype == t.dtype, k
            else:
                full = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            res[k] = full
            for r in range(world):
                seg = full[int(starts[r]):int(starts[r + 1])]
                if r == rank:
                    seg.copy_(t[:counts[r]])
                elif counts[r]:
                    ops.append(dist.P2POp(dist.irecv, seg, dist.get_global_rank(group, r) if group is not None else r, group))
    elif counts[rank]:
        peer = dist.get_global_rank(group, dst) if group is not None else dst
        for k in keys:
            ops.append(dist.P2POp(dist.isend, local[k][:counts[rank]].contiguous(), peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return res, counts


def gather_compact_result(res, n_pairs_local, dst=0, group=None, out=None):
    """
    Gather a compact (CSR) trace result (dict of tensors: per-pair "n_sol", "status", "sol_offset"[N+1]; everything else
    per-solution rows) on `dst`.  Returns on `dst` a dict with the pairs of all ranks in rank order, contiguous rows and
    a global sol_offset[N_total + 1]; None elsewhere.  The row counts come from sol_offset[N] (one host read per rank).
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    so = res["sol_offset"]
    n_rows = int(so[n_pairs_local].item())
    dev = so.device
    both = exchange_counts(n_rows * (1 << 32) + int(n_pairs_local), dev, group)     # one all_gather for both counts
    row_counts, pair_counts = [b >> 32 for b in both], [b & 0xffffffff for b in both]
    per_pair = {k: res[k] for k in PER_PAIR if k in res and k != "sol_offset"}
    per_pair["sol_offset"] = so[:n_pairs_local]
    per_row = {k: v for k, v in res.items() if k not in PER_PAIR}
    g_pair, _ = gather_rows(per_pair, n_pairs_local, dst, group, out, pair_counts)
    g_row, _ = gather_rows(per_row, n_rows, dst, group, out, row_counts) if per_row else ({}, None)
    if rank != dst:
        return None
    # shift every rank's offsets by the rows of the ranks before it; close the CSR
    pstart, rstart = np.concatenate([[0], np.cumsum(pair_counts)]), np.concatenate([[0], np.cumsum(row_counts)])
    so_all = g_pair["sol_offset"]
    for r in range(1, world):
        so_all[int(pstart[r]):int(pstart[r + 1])] += int(rstart[r])
    full_so = torch.empty(int(pstart[-1]) + 1, dtype=so_all.dtype, device=dev)
    full_so[:-1] = so_all[:int(pstart[-1])]
    full_so[-1] = int(rstart[-1])
    g_pair["sol_offset"] = full_so
    g_pair.update(g_row or {})
    g_pair["_row_counts"], g_pair["_pair_counts"] = row_counts, pair_counts
    return g_pair


def gather_compact(local, group=None, dst=0):
    """
    Round-1 interface, kept for callers of the padded layout: gather a dict of per-pair tensors (leading dimension = local
    pairs, unequal across ranks) on rank `dst` (gatherv over send/recv; returns None on the other ranks).
    """
    keys = sorted(local.keys())
    res, _ = gather_rows(local, local[keys[0]].shape[0], dst, group)
    return res


# ------------------------------------------------------------------------------------------------------------------
# product variant: the kernels store into the gathering GPU's HBM through NVLink peer mappings
# ------------------------------------------------------------------------------------------------------------------
class _DeviceArray:
    """__cuda_array_interface__ view of raw device memory (so that torch can wrap the gathered arrays without a copy)"""

    def __init__(self, ptr, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": np.dtype(dtype).str,
                                         "data": (int(ptr), False), "version": 2}


class P2PGather:
    """
    Destination of a fused trace + gather (see the module docstring).  Collective constructor: every rank of `group` calls
    it with its own pair count; `names` are the per-solution outputs to gather (the per-pair arrays n_sol, status,
    sol_offset always are).

        g = P2PGather(rt, n_pairs_local, names=("C0", "travel_time", ..., "attenuation_sparse"), Fs=37)
        res = g.trace(dv, da, outer=True, frequency=ff, ...)     # device-resident compact trace; rows land on rank `dst`
        g.finish()                                               # stream sync + barrier: rows visible on `dst`
        g.arrays                                                 # rank dst: dict of torch tensors over the gathered block

    Outputs not named in `names` (e.g. attenuation_sparse in the "records only" variant) stay in the local HBM of the
    rank that computed them (`res[...]`, local rows), sharded where the next stage consumes them.
    """

    def __init__(self, rt, n_pairs_local, names, Fs=0, F=0, dst=0, group=None, rows_per_pair=None):
        import torch
        import torch.distributed as dist
        from nuradiomc_b200 import _lib
        from nuradiomc_b200.SignalProp.analyticraytracing import _OUT_SPECS
        self.rt, self.dst, self.group = rt, dst, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = rt._device
        self.lib = _lib.load()
        S, K1 = rt.get_number_of_raytracing_solutions(), rt._n_reflections + 1
        self.names = [n for n in names if n not in PER_PAIR]
        cap_local = int(n_pairs_local * (S if rows_per_pair is None else rows_per_pair))
        dev = torch.device("cuda", self.device)
        if self.world > 1:
            both = exchange_counts(cap_local * (1 << 32) + int(n_pairs_local), dev, group)
        else:
            both = [cap_local * (1 << 32) + int(n_pairs_local)]
        self.row_caps, self.pair_counts = [b >> 32 for b in both], [b & 0xffffffff for b in both]
        self.row_base = [int(x) for x in np.concatenate([[0], np.cumsum(self.row_caps)])]
        self.pair_base = [int(x) for x in np.concatenate([[0], np.cumsum(self.pair_counts)])]
        rows_total, pairs_total = self.row_base[-1], self.pair_base[-1]
        # one block: every array starts on a 256-byte boundary
        self.layout, off = {}, 0
        for n in list(PER_PAIR) + self.names:
            if n == "sol_offset":
                dtype, shape = np.int64, (pairs_total,)
            else:
                dtype, trail = _OUT_SPECS[n]
                t = trail(S, K1, Fs, F)
                shape = (pairs_total,) if len(t) == 0 else (rows_total,) + tuple(t[1:])
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            self.layout[n] = (off, shape, np.dtype(dtype))
            off += (nbytes + 255) // 256 * 256
        self.block_bytes = max(off, 256)
        self._owned, self._mapped = None, None
        handle = None
        ptr = C.c_void_p()
        if self.rank == dst:
            hbuf = C.create_string_buffer(64)
            _lib.check(self.lib.nrmc_rt_peer_alloc(self.device, self.block_bytes, C.byref(ptr), hbuf), None, "peer_alloc")
            self._owned = ptr.value
            handle = hbuf.raw
        if self.world > 1:
            box = [handle]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
            handle = box[0]
            if self.rank != dst:
                _lib.check(self.lib.nrmc_rt_peer_open(self.device, handle, C.byref(ptr)), None,
                           "peer_open (CUDA IPC mapping of the gathering rank's block)")
                self._mapped = ptr.value
        self.base = self._owned if self.rank == dst else self._mapped
        self.arrays = None
        if self.rank == dst:
            self.arrays = {n: torch.as_tensor(_DeviceArray(self.base + o, shape, dt), device=dev)
                           for n, (o, shape, dt) in self.layout.items()}
        self._local = None

    def ptr(self, name):
        return self.base + self.layout[name][0]

    def nvlink_bytes(self, n_rows_local):
        """bytes this rank sends through NVLink per trace (0 on the gathering rank)"""
        if self.rank == self.dst:
            return 0
        per_row = sum(int(np.prod(self.layout[n][1][1:])) * self.layout[n][2].itemsize for n in self.names)
        return int(n_rows_local) * per_row + self.pair_counts[self.rank] * 16

    def trace(self, v, a, **kw):
        """device-resident compact trace of this rank's pairs whose result rows are stored in the gathering rank's block"""
        import torch
        r = self.rank
        out_ptrs = {n: self.ptr(n) for n in self.names}
        res = self.rt.trace_batch_device(v, a, compact=True, out=self._local, out_ptrs=out_ptrs, row_base=self.row_base[r],
                                         row_capacity=self.row_caps[r], **kw)
        self._local = res
        # the per-pair arrays are read back by the kernels (row scan, row lookup): they stay local and are shipped as three copies
        stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        n = self.pair_counts[r]
        from nuradiomc_b200 import _lib
        for name in PER_PAIR:
            if name not in res:
                continue
            off, _, dt = self.layout[name]
            _lib.check(self.lib.nrmc_rt_copy_async(C.c_void_p(self.base + off + self.pair_base[r] * dt.itemsize),
                                                   C.c_void_p(res[name].data_ptr()), n * dt.itemsize, C.c_void_p(stream)),
                       None, "copy_async")
        return res

    def trace_pushed(self, v, a, n_chunks=4, **kw):
        """
        The same gather with the rows moved by the copy engines instead of by the kernels' own stores ("p2p-dma"): the rank's
        vertices are traced chunk by chunk into local HBM (two alternating buffers); while chunk c+1 computes, the filled rows
        of chunk c are pushed into this rank's segment of the gathering rank's block by peer-to-peer cudaMemcpyAsync on a
        second stream (full-line NVLink writes at wire speed; the kernels' own 8-byte row stores reach about half of that).
        The only host involvement is reading one row count per chunk.  Same layout on the gathering rank as `trace`.
        """
        import torch
        from nuradiomc_b200 import _lib
        from nuradiomc_b200.SignalProp.analyticraytracing import BatchResult
        r, dev = self.rank, torch.device("cuda", self.device)
        outer = kw.get("outer", False)
        na = a.shape[1] if outer else 1
        nv = v.shape[1]
        n_chunks = max(1, min(int(n_chunks), nv))
        key = (v.data_ptr(), nv, n_chunks)
        if getattr(self, "_push_key", None) != key:
            bounds = [shard_bounds(nv, n_chunks, c) for c in range(n_chunks)]
            self._push_chunks = [(lo, hi, v[:, lo:hi].contiguous()) for lo, hi in bounds]
            self._push_bufs = [BatchResult(), BatchResult()]
            self._push_counts = torch.zeros(n_chunks, dtype=torch.int64).pin_memory()
            self._push_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]   # two copy engines share the rows
            self._push_key = key
        compute = torch.cuda.current_stream(dev)
        copy, copy2 = self._push_streams
        cap_chunk = max(self.row_caps[r] // n_chunks + 1024, 1024)
        ev_done, ev_copied = [None] * n_chunks, [None] * n_chunks
        state = {"rows": 0, "pairs": 0}

        def push(c):
            lo, hi, _ = self._push_chunks[c]
            res = self._push_bufs[c % 2]
            n_pairs_c = (hi - lo) * na
            ev_done[c].synchronize()                       # the GPU is already busy with chunk c + 1
            n = int(self._push_counts[c])
            row0 = self.row_base[r] + state["rows"]
            with torch.cuda.stream(copy):
                copy.wait_event(ev_done[c])
                copy2.wait_event(ev_done[c])
                res["sol_offset"][:n_pairs_c].add_(row0)   # global rows of this chunk's pairs
                cs, cs2 = copy.cuda_stream, copy2.cuda_stream
                half = n // 2
                for name in self.names:
                    off, shape, dt = self.layout[name]
                    rb = int(np.prod(shape[1:], dtype=np.int64)) * dt.itemsize
                    dst, src = self.base + off + row0 * rb, res[name].data_ptr()
                    _lib.check(self.lib.nrmc_rt_copy_async(C.c_void_p(dst), C.c_void_p(src), half * rb, C.c_void_p(cs)), None, "copy_async")
                    _lib.check(self.lib.nrmc_rt_copy_async(C.c_void_p(dst + half * rb), C.c_void_p(src + half * rb), (n - half) * rb,
                                                           C.c_void_p(cs2)), None, "copy_async")
                p0 = self.pair_base[r] + state["pairs"]
                for name in PER_PAIR:
                    if name in res:
                        off, _, dt = self.layout[name]
                        _lib.check(self.lib.nrmc_rt_copy_async(C.c_void_p(self.base + off + p0 * dt.itemsize), C.c_void_p(res[name].data_ptr()),
                                                               n_pairs_c * dt.itemsize, C.c_void_p(cs)), None, "copy_async")
                copy.wait_stream(copy2)
                ev_copied[c] = torch.cuda.Event()
                ev_copied[c].record(copy)
            state["rows"] += n
            state["pairs"] += n_pairs_c
        for c in range(n_chunks):
            lo, hi, vc = self._push_chunks[c]
            if c >= 2:
                compute.wait_event(ev_copied[c - 2])       # the buffer is free once its rows have left
            res = self.rt.trace_batch_device(vc, a, compact=True, out=self._push_bufs[c % 2], row_capacity=cap_chunk, **kw)
            self._push_bufs[c % 2] = res
            self._push_counts[c:c + 1].copy_(res["sol_offset"][(hi - lo) * na:(hi - lo) * na + 1], non_blocking=True)
            ev_done[c] = torch.cuda.Event()
            ev_done[c].record(compute)
            if c >= 1:
                push(c - 1)
        push(n_chunks - 1)
        compute.wait_stream(copy)
        self.n_rows_pushed = state["rows"]
        return state["rows"]

    def finish(self):
        """wait for this rank's stores, then for everybody's: afterwards `arrays` on the gathering rank is complete"""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(torch.device("cuda", self.device))
        if self.world > 1:
            dist.barrier(group=self.group)

    def compacted(self):
        """rank dst: the gathered result as ONE contiguous CSR (closes the gaps between the ranks' row segments);
        sol_offset gets its terminal entry.  A device-side copy of the filled rows, for consumers that need it."""
        import torch
        assert self.rank == self.dst
        A = self.arrays
        n_sol = A["n_sol"].to(torch.int64)
        counts = [int(n_sol[self.pair_base[r]:self.pair_base[r + 1]].sum().item()) for r in range(self.world)]
        starts = np.concatenate([[0], np.cumsum(counts)])
        out = {"n_sol": A["n_sol"], "status": A["status"]}
        so = torch.empty(self.pair_base[-1] + 1, dtype=torch.int64, device=A["n_sol"].device)
        for r in range(self.world):
            so[self.pair_base[r]:self.pair_base[r + 1]] = A["sol_offset"][self.pair_base[r]:self.pair_base[r + 1]] - (self.row_base[r] - int(starts[r]))
        so[-1] = int(starts[-1])
        out["sol_offset"] = so
        for n in self.names:
            out[n] = torch.cat([A[n][self.row_base[r]:self.row_base[r] + counts[r]] for r in range(self.world)], dim=0)
        return out

    def close(self):
        import torch.distributed as dist
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)          # nobody may still be storing into the block
        if self._mapped:
            self.lib.nrmc_rt_peer_close(C.c_void_p(self._mapped))
            self._mapped = None
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        if self._owned:
            self.arrays = None
            self.lib.nrmc_rt_peer_free(C.c_void_p(self._owned))
            self._owned = None
