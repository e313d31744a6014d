"""x-slab decomposed particle-mesh step: one slab of N/P mesh planes (and the particles inside it) per GPU.

The reference is single-process (README.md:49); this module is the multi-GPU form of its hot path
(`integration.integrate` integration.py:17 -> `leapfrog` :192 -> `solver.pm` solver.py:30 with
`linear_newton_solver = fft`), laid out as SURVEY 8(e) asks:

  rank r of P owns planes [r N/P, (r+1) N/P) of every grid and the particles whose cell floor(x N) is in them
  kick + drift + wrap          local                                  (psc_kick_drift_wrap)
  migration                    counts all-to-all + records all-to-all (psc_slab_count / pack / unpack / move)
  deposit                      local bins -> rho[nxl + 2 ghost planes] (psc_bin_particles_slab, psc_deposit_binned_slab)
                               ghost planes SENT to the neighbours and ADDED there
  FFT Poisson solve            2-D R2C per plane -> y-block all-to-all -> 1-D C2C along x -> Green ->
                               inverse 1-D -> all-to-all -> 2-D C2R      (psc_slab_fft_*, psc_green_slab)
  multigrid Poisson solve      (linear_newton_solver = multigrid) V-cycles on ghosted slabs: ghost planes refreshed
                               before every half-sweep / residual / prolongation, coarse levels thinner than two
                               planes per rank gathered and solved redundantly (slab_multigrid.py, psc_box_*)
  gradient + interpolation     potential with G = 1 + stencil-reach ghost planes COPIED from the neighbours
                               (psc_interp_kick_phi_binned_slab: gradient, TSC gather, half-kick, max|a|, max|v|)
  time step                    all-reduce(max) of two floats

Collectives go through a small `Comm` interface with three implementations: `SelfComm` (P = 1),
`TorchComm` (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests) and `ThreadComm` (P virtual
ranks as threads of one process sharing one GPU: lets the single-GPU test tier exercise the whole slab path).
The kernels are reached through an `ops` object (`CudaOps` = the C ABI); tests may pass another implementation
of the same methods (tests/slab_oracle_ops.py: the CPU oracle) to check the host logic without a GPU -- the
product default is CudaOps and nothing here falls back to it silently.
"""
import ctypes as C
import logging
import os
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

REC = 8  # floats per migration record (x y z vx vy vz id_lo id_hi)


# ------------------------------------------------------------------------------------------ comms
class SelfComm:
    """P = 1: every exchange is with oneself (periodic wrap)."""
    rank, size = 0, 1

    def exchange_counts(self, counts):
        return list(counts)

    def all_to_all_v(self, send, send_counts, recv_counts):
        return send

    def all_to_all_equal(self, send, out):
        out.copy_(send)
        return out

    def exchange_planes(self, to_left, to_right):
        return to_right, to_left

    def neighbor_counts(self, n_to_left, n_to_right):
        return n_to_right, n_to_left

    def neighbor_exchange(self, to_left, to_right, n_from_left, n_from_right):
        return to_right, to_left

    pair_exchange = neighbor_exchange

    def allreduce_max_(self, t):
        return t

    def allreduce_sum_(self, t):
        return t

    def barrier(self):
        pass


class TorchComm:
    """torch.distributed process group (NCCL on GPUs, gloo on CPU).  Everything is expressed as all-to-all /
    all-reduce so that the same code runs on both backends."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._mbox = None          # (buffer, handle, floats per slot) | False when peer memory is unavailable
        self._mbox_set = 0
        self._views = {}           # cached tensor views of the peers' mailbox slots
        self._use_mailbox = not os.environ.get("PSC_NO_PEER_MAILBOX")

    def exchange_counts(self, counts):
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        s = torch.tensor(list(counts), dtype=torch.int64, device=dev)
        r = torch.empty_like(s)
        dist.all_to_all_single(r, s, group=self.group)
        return [int(v) for v in r.cpu().tolist()]

    def all_to_all_v(self, send, send_counts, recv_counts):
        recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), list(recv_counts), list(send_counts), group=self.group)
        return recv

    def all_to_all_equal(self, send, out):
        dist.all_to_all_single(out, send, group=self.group)
        return out

    # -- neighbour messages through peer memory: every rank owns a "mailbox" in symmetric memory with two slots (from
    # the left / from the right neighbour) in two alternating sets; a sender copies its message straight into the
    # neighbour's slot over NVLink (a plain device-to-device copy into the mapped peer buffer), an inter-GPU barrier
    # on the stream publishes it.  The alternating sets make one barrier per exchange enough: a slot is rewritten two
    # exchanges later, after a barrier that its owner only reaches once it has consumed the previous content.
    def _mailbox(self, numel):
        """symmetric mailbox with room for `numel` float32 per slot, or None (NCCL p2p is used then).  Growing it is a
        collective (symmetric-memory rendezvous), so only messages whose size is the same on every rank may come through
        here: ghost planes and the fixed-capacity migration messages; rank-dependent sizes use pair_exchange."""
        if self._mbox is False:
            return None
        if self._mbox is None or self._mbox[2] < numel:
            cap = int(numel * 1.25) + 1024
            bufs = self.symmetric_buffers(4 * cap, 1)
            if not bufs:
                self._mbox = False
                return None
            self._mbox = (bufs[0][0], bufs[0][1], cap)
            self._mbox_set = 0
            self._views = {}
        return self._mbox

    def _peer_view(self, hdl, peer, numel, offset):
        """tensor view of a peer's mailbox slot (cached: creating one costs more host time than the copy it serves)"""
        key = (peer, numel, offset)
        v = self._views.get(key)
        if v is None:
            if len(self._views) > 64:
                self._views.clear()
            v = self._views[key] = hdl.get_buffer(peer, (numel,), torch.float32, offset)
        return v

    def _p2p(self, to_left, to_right, from_left, from_right):
        P, r = self.size, self.rank
        left, right = (r - 1) % P, (r + 1) % P
        n = max(to_left.numel(), to_right.numel(), from_left.numel(), from_right.numel())
        mb = self._mailbox(n) if (to_left.is_cuda and to_left.dtype == torch.float32 and self._use_mailbox) else None
        if mb is None:
            return self._p2p_nccl(to_left, to_right, from_left, from_right)
        buf, hdl, cap = mb
        base = self._mbox_set * 2 * cap          # slot 0: from the left neighbour, slot 1: from the right neighbour
        self._mbox_set ^= 1
        if to_left.numel():      # I am my left neighbour's RIGHT neighbour
            self._peer_view(hdl, left, to_left.numel(), base + cap).copy_(to_left.reshape(-1))
        if to_right.numel():
            self._peer_view(hdl, right, to_right.numel(), base).copy_(to_right.reshape(-1))
        hdl.barrier(channel=0)
        if from_left.numel():
            from_left.reshape(-1).copy_(buf[base:base + from_left.numel()])
        if from_right.numel():
            from_right.reshape(-1).copy_(buf[base + cap:base + cap + from_right.numel()])

    def _p2p_nccl(self, to_left, to_right, from_left, from_right):
        """One batch of point-to-point messages with the two neighbours (4 NCCL p2p ops instead of the P (P - 1)
        of an all-to-all).  Message order per peer is what pairs sends with receives when P == 2 (the peer is both
        neighbours): sends go out as [to_left, to_right], receives are posted as [from_right, from_left]."""
        P, r = self.size, self.rank
        left, right = (r - 1) % P, (r + 1) % P
        ops = []
        if to_left.numel():
            ops.append(dist.P2POp(dist.isend, to_left.contiguous(), left, self.group))
        if to_right.numel():
            ops.append(dist.P2POp(dist.isend, to_right.contiguous(), right, self.group))
        if from_right.numel():
            ops.append(dist.P2POp(dist.irecv, from_right, right, self.group))
        if from_left.numel():
            ops.append(dist.P2POp(dist.irecv, from_left, left, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_planes(self, to_left, to_right):
        assert to_right.shape == to_left.shape
        from_left, from_right = torch.empty_like(to_left), torch.empty_like(to_right)
        self._p2p(to_left, to_right, from_left, from_right)
        return from_left, from_right

    def neighbor_counts(self, n_to_left, n_to_right):
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        s = torch.tensor([int(n_to_left), int(n_to_right)], dtype=torch.int64, device=dev)
        r = torch.empty_like(s)
        self._p2p(s[0:1], s[1:2], r[0:1], r[1:2])
        r = r.cpu().tolist()
        return int(r[0]), int(r[1])

    def neighbor_exchange(self, to_left, to_right, n_from_left, n_from_right):
        shape = tuple(to_left.shape[1:])
        from_left = torch.empty((n_from_left,) + shape, dtype=to_left.dtype, device=to_left.device)
        from_right = torch.empty((n_from_right,) + shape, dtype=to_left.dtype, device=to_left.device)
        self._p2p(to_left, to_right, from_left, from_right)
        return from_left, from_right

    def pair_exchange(self, to_left, to_right, n_from_left, n_from_right):
        """neighbor_exchange restricted to point-to-point operations: a message whose size is 0 on both of its ends is
        not posted at all, and nothing synchronises the group -- what the overflow path of the migration needs, where
        only the pairs whose message did not fit take part."""
        shape = tuple(to_left.shape[1:])
        from_left = torch.empty((n_from_left,) + shape, dtype=to_left.dtype, device=to_left.device)
        from_right = torch.empty((n_from_right,) + shape, dtype=to_left.dtype, device=to_left.device)
        self._p2p_nccl(to_left, to_right, from_left, from_right)
        return from_left, from_right

    def symmetric_buffers(self, numel, count):
        """`count` float32 buffers of `numel` elements that every rank can address directly over NVLink
        (torch symmetric memory: CUDA VMM allocations exchanged at a rendezvous).  Returns [(tensor, handle)] --
        handle.buffer_ptrs_dev is the device array of the P peer pointers, handle.barrier() an inter-GPU barrier on
        the current stream -- or None when peer memory is unavailable (the caller then uses the NCCL all-to-all).
        Collective: every rank must call it."""
        if dist.get_backend(self.group) != "nccl" or os.environ.get("PSC_NO_PEER_MEMORY"):
            return None
        out = []
        ok = 1.0
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dev = torch.device("cuda", torch.cuda.current_device())
            for _ in range(count):
                t = symm_mem.empty(int(numel), dtype=torch.float32, device=dev)
                h = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
                out.append((t, h))
        except Exception as e:  # noqa: BLE001
            logging.warning(f"peer memory unavailable ({type(e).__name__}: {e}); using NCCL all-to-all transposes")
            ok = 0.0
        flag = torch.tensor([ok], dtype=torch.float32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return out if flag.item() > 0 else None

    def allreduce_max_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def allreduce_sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def barrier(self):
        dist.barrier(group=self.group)


class _ThreadWorld:
    def __init__(self, size):
        import queue
        self.size = size
        self.barrier = threading.Barrier(size)
        self.slots = [None] * size
        # point-to-point channels (src, dst, direction) for pair_exchange
        self.pipes = {(a, b, d): queue.Queue() for a in range(size) for b in range(size) for d in "LR"}


class ThreadComm:
    """P virtual ranks = P threads of one process (all on the current CUDA device, same stream).  Used by the
    single-GPU tests and tools to run the complete slab path without NCCL.  Create with ThreadComm.world(P)."""

    def __init__(self, world, rank):
        self.w, self.rank, self.size = world, rank, world.size

    @staticmethod
    def world(size):
        w = _ThreadWorld(size)
        return [ThreadComm(w, r) for r in range(size)]

    def _share(self, obj):
        self.w.slots[self.rank] = obj
        self.w.barrier.wait()
        allv = list(self.w.slots)
        self.w.barrier.wait()
        return allv

    def barrier(self):
        self.w.barrier.wait()

    def exchange_counts(self, counts):
        allc = self._share(list(counts))
        return [int(allc[s][self.rank]) for s in range(self.size)]

    def all_to_all_v(self, send, send_counts, recv_counts):
        offs = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
        alls = self._share((send, offs))
        parts = []
        for s in range(self.size):
            t, o = alls[s]
            parts.append(t[int(o[self.rank]):int(o[self.rank + 1])])
        out = torch.cat(parts)
        self.w.barrier.wait()  # senders may reuse their buffers only after everybody copied
        return out

    def all_to_all_equal(self, send, out):
        n = send.shape[0] // self.size
        alls = self._share(send)
        for s in range(self.size):
            out[s * n:(s + 1) * n] = alls[s][self.rank * n:(self.rank + 1) * n]
        self.w.barrier.wait()
        return out

    def exchange_planes(self, to_left, to_right):
        alls = self._share((to_left, to_right))
        left, right = (self.rank - 1) % self.size, (self.rank + 1) % self.size
        out = alls[left][1].clone(), alls[right][0].clone()
        self.w.barrier.wait()
        return out

    def neighbor_counts(self, n_to_left, n_to_right):
        alls = self._share((int(n_to_left), int(n_to_right)))
        left, right = (self.rank - 1) % self.size, (self.rank + 1) % self.size
        return alls[left][1], alls[right][0]

    def neighbor_exchange(self, to_left, to_right, n_from_left, n_from_right):
        out = self.exchange_planes(to_left, to_right)
        assert out[0].shape[0] == n_from_left and out[1].shape[0] == n_from_right
        return out

    def pair_exchange(self, to_left, to_right, n_from_left, n_from_right):
        """point-to-point only (no barrier): empty messages are not posted"""
        left, right = (self.rank - 1) % self.size, (self.rank + 1) % self.size
        if to_left.shape[0]:
            self.w.pipes[(self.rank, left, "L")].put(to_left.clone())
        if to_right.shape[0]:
            self.w.pipes[(self.rank, right, "R")].put(to_right.clone())
        from_left = self.w.pipes[(left, self.rank, "R")].get(timeout=120) if n_from_left else to_left[:0]
        from_right = self.w.pipes[(right, self.rank, "L")].get(timeout=120) if n_from_right else to_right[:0]
        assert from_left.shape[0] == n_from_left and from_right.shape[0] == n_from_right
        return from_left, from_right

    def _allreduce(self, t, fn):
        alls = self._share(t.clone())
        acc = alls[0].clone()
        for o in alls[1:]:
            acc = fn(acc, o)
        t.copy_(acc)
        return t

    def allreduce_max_(self, t):
        return self._allreduce(t, torch.maximum)

    def allreduce_sum_(self, t):
        return self._allreduce(t, torch.add)


def default_comm():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchComm()
    return SelfComm()


# ------------------------------------------------------------------------------------------ kernels
class CudaOps:
    """The slab kernels through the C ABI (include/pysco_b200.h).  One instance per slab."""

    def __init__(self, N, P, rank):
        self.N, self.P, self.rank = N, P, rank
        self.nxl = N // P
        self.x0 = rank * self.nxl
        self.nyl = N // P
        self.y0 = rank * self.nyl
        self.lib = _lib.load()
        self.dev = _lib.device()
        self._plan = None
        self._work = None
        self._scratch = None
        self._sorted = None
        self._table = -1      # which of the scratch's two bin tables describes the sorted arrays (-1: none)
        self._leavers = None
        self._mig = None
        self._sortws = None

    # -- particles
    def kick_drift_wrap(self, pos, vel, acc, half_dt, dt, dt_is_f64):
        _lib.check(self.lib.psc_kick_drift_wrap(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), pos.shape[0],
                                                float(half_dt), float(dt), int(dt_is_f64), _lib.stream()))

    def kick_drift_wrap_detect(self, pos, vel, acc, half_dt, dt, dt_is_f64):
        """kick + drift + wrap with the leavers found in the same pass: returns (counts[P + 1], rows) device
        tensors; counts[P] = number of leavers, rows[:min(counts[P], len(rows))] their rows."""
        n = pos.shape[0]
        cap = max(4096, n // 8)
        if self._leavers is None or self._leavers.numel() < cap:
            self._leavers = torch.empty((cap,), dtype=torch.int64, device=self.dev)
        counts = torch.empty((self.P + 1,), dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.psc_kick_drift_wrap_slab(
            _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(acc), n, float(half_dt), float(dt), int(dt_is_f64), self.N,
            self.nxl, self.P, self.rank, _lib.ptr(counts), _lib.ptr(self._leavers), self._leavers.numel(),
            _lib.stream()))
        return counts, self._leavers

    def pack_rows(self, pos, vel, ids, rows, offsets, nout):
        sendbuf = torch.empty((nout, REC), dtype=torch.float32, device=self.dev)
        holes = torch.empty((nout,), dtype=torch.int64, device=self.dev)
        cursor = torch.empty((self.P,), dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.psc_slab_pack_rows(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(ids), _lib.ptr(rows),
                                               rows.shape[0], self.N, self.nxl, self.P, self.rank, _lib.ptr(offsets),
                                               _lib.ptr(cursor), _lib.ptr(sendbuf), _lib.ptr(holes), _lib.stream()))
        return sendbuf, holes

    def pack_fixed(self, pos, vel, ids, detected, cap):
        """Leavers towards the left / right neighbour into fixed-capacity buffers (header record + cap records);
        returns (sendL, sendR, holesL, holesR, status[3]) -- all on the device, no host synchronisation."""
        n = pos.shape[0]
        if self._mig is None or self._mig[0].shape[0] != cap + 1:
            self._mig = (torch.zeros((cap + 1, REC), dtype=torch.float32, device=self.dev),
                         torch.zeros((cap + 1, REC), dtype=torch.float32, device=self.dev),
                         torch.empty((cap,), dtype=torch.int64, device=self.dev),
                         torch.empty((cap,), dtype=torch.int64, device=self.dev))
        sendL, sendR, holesL, holesR = self._mig
        status = torch.empty((3,), dtype=torch.int64, device=self.dev)
        if detected is not None:
            counts, rows = detected
            rows_ptr, cap_rows = _lib.ptr(rows), rows.numel()
        else:
            counts = torch.zeros((self.P + 1,), dtype=torch.int64, device=self.dev)
            rows_ptr, cap_rows = None, 0
        _lib.check(self.lib.psc_slab_pack_fixed(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(ids), n, rows_ptr,
                                                _lib.ptr(counts), cap_rows, self.N, self.nxl, self.P, self.rank, cap,
                                                _lib.ptr(sendL), _lib.ptr(sendR), _lib.ptr(holesL), _lib.ptr(holesR),
                                                _lib.ptr(status), _lib.stream()))
        return sendL, sendR, holesL, holesR, status

    def count_owners(self, pos):
        counts = torch.empty((self.P,), dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.psc_slab_count(_lib.ptr(pos), pos.shape[0], self.N, self.nxl, self.P, self.rank,
                                           _lib.ptr(counts), _lib.stream()))
        return counts

    def pack_leavers(self, pos, vel, ids, offsets, nout):
        sendbuf = torch.empty((nout, REC), dtype=torch.float32, device=self.dev)
        holes = torch.empty((nout,), dtype=torch.int64, device=self.dev)
        cursor = torch.empty((self.P,), dtype=torch.int64, device=self.dev)
        _lib.check(self.lib.psc_slab_pack_leavers(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(ids), pos.shape[0], self.N,
                                                  self.nxl, self.P, self.rank, _lib.ptr(offsets), _lib.ptr(cursor),
                                                  _lib.ptr(sendbuf), _lib.ptr(holes), _lib.stream()))
        return sendbuf, holes

    def unpack_rows(self, recvbuf, rows, pos, vel, ids):
        _lib.check(self.lib.psc_slab_unpack_rows(_lib.ptr(recvbuf), _lib.ptr(rows), rows.shape[0], _lib.ptr(pos),
                                                 _lib.ptr(vel), _lib.ptr(ids), _lib.stream()))

    def move_rows(self, src, dst, pos, vel, ids):
        _lib.check(self.lib.psc_slab_move_rows(_lib.ptr(src), _lib.ptr(dst), src.shape[0], _lib.ptr(pos),
                                               _lib.ptr(vel), _lib.ptr(ids), _lib.stream()))

    def max_abs(self, x):
        out = torch.zeros((1,), dtype=torch.float32, device=self.dev)
        if x.numel():
            _lib.check(self.lib.psc_max_abs(_lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream()))
        return out

    def morton_order(self, pos):
        """permutation that sorts the local particles by Morton key (morton.py:42-137, utils.py:1019-1075).  The keys,
        the permutation and the sort scratch live in one persistent (grow-only) workspace: a reorder allocates
        nothing, so it cannot push the caching allocator into cudaMalloc / cudaFree storms in the step after it."""
        n = pos.shape[0]
        nbytes = int(self.lib.psc_argsort_workspace_bytes(n))
        need = 16 * n + nbytes + 512
        if self._sortws is None or self._sortws.numel() < need:
            self._sortws = None
            self._sortws = torch.empty((int(need * 1.1) + 1024,), dtype=torch.uint8, device=self.dev)
        ws = self._sortws
        keys = ws[:8 * n].view(torch.int64)
        idx = ws[8 * n:16 * n].view(torch.int64)
        off = (16 * n + 255) // 256 * 256
        scratch = ws[off:]
        _lib.check(self.lib.psc_morton_keys(_lib.ptr(pos), n, _lib.ptr(keys), _lib.stream()))
        _lib.check(self.lib.psc_argsort_keys(_lib.ptr(keys), n, _lib.ptr(idx), _lib.ptr(scratch), scratch.numel(),
                                             _lib.stream()))
        return idx

    def gather_rows(self, idx, a, out=None):
        out = torch.empty_like(a) if out is None else out
        _lib.check(self.lib.psc_gather3(_lib.ptr(idx), _lib.ptr(a), _lib.ptr(out), a.shape[0], _lib.stream()))
        return out

    # -- particles <-> mesh
    def bin(self, pos):
        n = pos.shape[0]
        nbytes = int(self.lib.psc_bin_workspace_bytes_slab(n, self.N, self.nxl))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = None
            self._scratch = torch.empty((int(nbytes * 1.1) + 256,), dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.psc_bin_particles_slab(_lib.ptr(pos), n, self.N, self.x0, self.nxl,
                                                   _lib.ptr(self._scratch), self._scratch.numel(), _lib.stream()))
        return n

    # -- particles <-> mesh on particle arrays kept in bin order (no binned copy, no source-row indirection)
    sorted_layout = True

    def sort_by_bin(self, pos, vel, ids, pos_out, vel_out, ids_out, src_rows=None):
        """Sort of the rank's (position, velocity, id) rows into bin order; leaves the bin table in the scratch.
        src_rows: the input arrays are the output of the previous sort (its first src_rows rows), updated in place by the
        kick + drift and the migration since -- then every CTA sorts one source bin in shared memory
        (csrc/binned.cu step_sort_local_kernel); None: arrays in no particular order, one global atomic per particle."""
        n = pos.shape[0]
        nbytes = int(self.lib.psc_sorted_workspace_bytes_slab(n, self.N, self.nxl))
        if self._sorted is None or self._sorted.numel() < nbytes:
            # a new scratch has no table: the bin tables live at offsets that depend on nothing but (N, nxl), but a
            # fresh buffer is uninitialised
            self._sorted = None
            self._sorted = torch.empty((int(nbytes * 1.1) + 256,), dtype=torch.uint8, device=self.dev)
            self._table = -1
        src = self._table if (src_rows is not None and not os.environ.get("PSC_NO_LOCAL_SORT")) else -1
        _lib.check(self.lib.psc_sort_by_bin_slab(_lib.ptr(pos), _lib.ptr(vel), _lib.ptr(ids), n, self.N, self.x0, self.nxl,
                                                 src, int(src_rows or 0), _lib.ptr(self._sorted), self._sorted.numel(),
                                                 _lib.ptr(pos_out), _lib.ptr(vel_out), _lib.ptr(ids_out), _lib.stream()))
        self._table = 0 if src < 0 else 1 - src
        return n

    def deposit_sorted(self, pos, scheme):
        rho = torch.empty((self.nxl + 2, self.N, self.N), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_deposit_sorted_slab(_lib.ptr(pos), _lib.ptr(self._sorted), self._sorted.numel(),
                                                    self._table, pos.shape[0], self.N, self.x0, self.nxl, scheme,
                                                    _lib.ptr(rho), _lib.stream()))
        return rho

    def interp_kick_phi_sorted(self, phi_g, ghost, order, pos, vel, acc, scheme, half_dt, u_g=None, f=0.0, fr_n=0):
        mx = torch.zeros((2,), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_interp_kick_phi_sorted_slab(
            _lib.ptr(phi_g), _lib.ptr(u_g), float(f), int(fr_n), order, self.x0, self.nxl, ghost, _lib.ptr(pos),
            _lib.ptr(self._sorted), self._sorted.numel(), self._table, _lib.ptr(vel), _lib.ptr(acc), pos.shape[0],
            self.N, scheme, float(half_dt), _lib.ptr(mx), _lib.stream()))
        return mx

    def deposit(self, binned, scheme):
        rho = torch.empty((self.nxl + 2, self.N, self.N), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_deposit_binned_slab(_lib.ptr(self._scratch), self._scratch.numel(), binned, self.N,
                                                    self.x0, self.nxl, scheme, _lib.ptr(rho), _lib.stream()))
        return rho

    def affine(self, x, f1, f2):
        _lib.check(self.lib.psc_linear_operator(_lib.ptr(x), float(f1), float(f2), _lib.ptr(x), x.numel(),
                                                _lib.stream()))

    def interp_kick_phi(self, phi_g, ghost, order, binned, vel, acc, scheme, half_dt, u_g=None, f=0.0, fr_n=0):
        """u_g, f, fr_n: the f(R) fifth force, gradient of phi + f u^(fr_n + 1) (mesh.py:860-2069)"""
        mx = torch.zeros((2,), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_interp_kick_phi_binned_slab(
            _lib.ptr(phi_g), _lib.ptr(u_g), float(f), int(fr_n), order, self.x0, self.nxl, ghost, _lib.ptr(self._scratch),
            self._scratch.numel(), _lib.ptr(vel), _lib.ptr(acc), binned, self.N, scheme, float(half_dt), _lib.ptr(mx),
            _lib.stream()))
        return mx

    # -- transposed FFT
    def _fft_plan(self):
        if self._plan is None:
            h = C.c_void_p()
            _lib.check(self.lib.psc_slab_fft_plan_create(self.N, self.nxl, self.nyl, C.byref(h)))
            nbytes = int(self.lib.psc_slab_fft_workspace_bytes(h))
            self._work = torch.empty((max(nbytes, 256),), dtype=torch.uint8, device=self.dev)
            _lib.check(self.lib.psc_slab_fft_set_workspace(h, _lib.ptr(self._work)))
            self._plan = h
        return self._plan

    def spectrum_buffer(self):
        return torch.empty((self.nxl * self.N * (self.N // 2 + 1), 2), dtype=torch.float32, device=self.dev)

    def fft2d_r2c(self, planes, spec2d):
        _lib.check(self.lib.psc_slab_fft_r2c_planes(self._fft_plan(), _lib.ptr(planes), _lib.ptr(spec2d),
                                                    _lib.stream()))

    def fft2d_c2r(self, spec2d, planes):
        _lib.check(self.lib.psc_slab_fft_c2r_planes(self._fft_plan(), _lib.ptr(spec2d), _lib.ptr(planes),
                                                    _lib.stream()))

    def yblocks(self, src, dst, to_blocks):
        _lib.check(self.lib.psc_slab_yblocks(_lib.ptr(src), _lib.ptr(dst), self.N, self.nxl, self.nyl,
                                             int(to_blocks), _lib.stream()))

    def transpose_put(self, src, peer_ptrs_dev, forward):
        _lib.check(self.lib.psc_slab_transpose_put(_lib.ptr(src), int(peer_ptrs_dev), self.N, self.nxl, self.nyl,
                                                   self.P, self.rank, int(forward), _lib.stream()))

    def fft_x(self, spec_t, inverse):
        _lib.check(self.lib.psc_slab_fft_x(self._fft_plan(), _lib.ptr(spec_t), int(inverse), _lib.stream()))

    def pk_bins(self, spec_t, p):
        """this rank's share of the P(k) bins [3][N] (sum |k|, sum |delta_k W^-p|^2, mode count), float64"""
        bins = torch.empty((3, self.N), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.psc_pk_slab(_lib.ptr(spec_t), self.N, self.nyl, self.y0, int(p), _lib.ptr(bins),
                                        _lib.stream()))
        return bins

    def xfft_green(self, spec_t, kind, p, scale):
        """forward FFT along x, Green's function, backward FFT along x of the transposed spectrum in ONE kernel
        (csrc/fourier.cu xfft_green_kernel); False when N is not a power of two in [64, 2048] (the caller then runs the
        three separate steps)"""
        if not self.lib.psc_fft_poisson_supported(self.N) or os.environ.get("PSC_NO_FUSED_XFFT"):
            return False
        _lib.check(self.lib.psc_xfft_green_slab(_lib.ptr(spec_t), self.N, self.nyl, self.y0, kind, p, float(scale),
                                                _lib.stream()))
        return True

    def green(self, spec_t, kind, p, scale):
        _lib.check(self.lib.psc_green_slab(_lib.ptr(spec_t), self.N, self.nyl, self.y0, kind, p, float(scale),
                                           _lib.stream()))

    # -- multigrid on the slab (csrc/slab_mg.cu; host sequencing in slab_multigrid.py)
    def mg_gs_colour(self, xg, b, nxl, n, x0, colour, f_relax):
        _lib.check(self.lib.psc_box_gauss_seidel_colour(_lib.ptr(xg), _lib.ptr(b), nxl, n, x0, colour,
                                                        float(f_relax), _lib.stream()))

    def mg_operator(self, xg, nxl, n):
        out = torch.empty((nxl, n, n), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_box_operator(_lib.ptr(xg), nxl, n, _lib.ptr(out), _lib.stream()))
        return out

    def mg_restrict_residual(self, xg, b, nxl, n):
        out = torch.empty((nxl // 2, n // 2, n // 2), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_box_restrict_residual(_lib.ptr(xg), _lib.ptr(b), nxl, n, _lib.ptr(out),
                                                      _lib.stream()))
        return out

    def mg_restriction(self, fine, nxl, n, sign, out=None):
        if out is None:
            out = torch.empty((nxl // 2, n // 2, n // 2), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_box_restriction(_lib.ptr(fine), nxl, n, float(sign), _lib.ptr(out), _lib.stream()))
        return out

    def mg_add_prolongation(self, fine_g, coarse_g, nxlc, nc):
        _lib.check(self.lib.psc_box_add_prolongation(_lib.ptr(fine_g), _lib.ptr(coarse_g), nxlc, nc, _lib.stream()))

    def mg_diff_sumsq(self, a, fa, b):
        out = torch.zeros((1,), dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.psc_diff_sumsq(_lib.ptr(a), float(fa), _lib.ptr(b), a.numel(), _lib.ptr(out),
                                           _lib.stream()))
        return out

    def mg_gs_colour_fr(self, xg, b, rhs, q, nxl, n, x0, colour, f_relax, kind):
        _lib.check(self.lib.psc_box_gauss_seidel_colour_fr(_lib.ptr(xg), _lib.ptr(b), _lib.ptr(rhs), float(q), nxl, n,
                                                           x0, colour, float(f_relax), kind, _lib.stream()))

    def mg_operator_fr(self, xg, b, q, nxl, n, kind):
        out = torch.empty((nxl, n, n), dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.psc_box_operator_fr(_lib.ptr(xg), _lib.ptr(b), float(q), nxl, n, kind, _lib.ptr(out),
                                                _lib.stream()))
        return out

    def mg_init_fr(self, b, q, nxl, n, kind, out):
        _lib.check(self.lib.psc_box_initialise_potential_fr(_lib.ptr(b), float(q), nxl, n, kind, _lib.ptr(out),
                                                            _lib.stream()))

    def lincomb(self, x, f1, y, f2):
        """utils.linear_operator_vectors_inplace (utils.py:721-755): x = f1 x + f2 y"""
        _lib.check(self.lib.psc_lincomb(_lib.ptr(x), float(f1), _lib.ptr(y), float(f2), x.numel(), _lib.stream()))

    def axpy(self, y, x, a):
        """utils.add_vector_scalar_inplace (utils.py:264-297) with a float32 scalar: y += a x"""
        _lib.check(self.lib.psc_axpy(_lib.ptr(y), _lib.ptr(x), float(a), 0, y.numel(), _lib.stream()))

    def mg_cube_solve_fas(self, x_c, b_c, res_c, param, nlevel, coarsest):
        """a gathered coarse FAS problem, solved on every rank by the single-domain kernels (multigrid.py:540-570):
        returns the correction x_corr - x_c"""
        from . import multigrid, utils
        L_c = multigrid.operator(x_c, param, b_c)
        utils.linear_operator_vectors_inplace(res_c, np.float32(4), L_c, np.float32(1))
        corr = x_c.clone()
        if coarsest:
            multigrid.smoothing(corr, b_c, param["Npre"], param, res_c)
        else:
            multigrid._cycle_FAS("V", corr, b_c, param, nlevel + 1, res_c)
        utils.add_vector_scalar_inplace(corr, x_c, np.float32(-1))
        return corr

    def mond_rhs(self, phig, out, nxl, n, g0, fn, alpha):
        _lib.check(self.lib.psc_box_mond_rhs(_lib.ptr(phig), _lib.ptr(out), nxl, n, float(np.float32(g0)), int(fn),
                                             float(alpha), _lib.stream()))

    def mg_cube_solve(self, res_cube, param, nlevel, coarsest):
        """a gathered coarse level, solved on every rank by the single-domain kernels (multigrid.py:474-517)"""
        from . import laplacian, multigrid
        x = laplacian.initialise_potential(res_cube)
        if coarsest:
            laplacian.smoothing(x, res_cube, param["Npre"])
        else:
            multigrid._cycle("V", x, res_cube, param, nlevel + 1)
        return x

    def close(self):
        if self._plan is not None:
            self.lib.psc_slab_fft_plan_destroy(self._plan)
            self._plan = None
        self._work = self._scratch = None


# ------------------------------------------------------------------------------------------ the slab
_SCHEMES = {"cic": (_lib.CIC, 2), "tsc": (_lib.TSC, 3)}
_REACH = {2: 1, 3: 1, 5: 2, 7: 3}


class Slab:
    """State and step of one rank: particle buffers (with spare capacity for arrivals), ids, and the kernels."""

    def __init__(self, ncells_1d, comm=None, ops=None, capacity_factor=1.3):
        self.comm = comm if comm is not None else default_comm()
        self.N = int(ncells_1d)
        self.P, self.rank = self.comm.size, self.comm.rank
        if self.N % self.P or (self.N // self.P) % 8:
            raise ValueError(f"slab decomposition needs N / P to be a multiple of 8 (N={self.N}, P={self.P})")
        self.nxl = self.N // self.P
        self.x0 = self.rank * self.nxl
        self.ops = ops if ops is not None else CudaOps(self.N, self.P, self.rank)
        self.capacity_factor = capacity_factor
        self.np = 0
        self.pos = self.vel = self.acc = self.ids = None
        self.max_acc = self.max_vel = None
        self.potential = None  # owned planes [nxl, N, N] of the last solve (a view into the ghosted array)
        self.additional_field = None   # owned planes of the last step's Newtonian potential (MOND) / scalaron (f(R))
        self.migrated_last = (0, 0)
        self._warm_host_ops()
        self._spare3 = self._spare1 = None   # spare particle buffers the reorder gathers into (then swapped in)
        self._sorted_rows = None             # rows [0, _sorted_rows) are the output of the last sort into bins
        self._mg = None           # SlabMultigrid (slab_multigrid.py), built at the first multigrid solve
        self._peer = None         # symmetric (peer-addressable) spectrum buffers, resolved at the first solve
        self._mig_cap = None      # records per direction of the fixed-capacity migration buffers (same on all ranks)
        self._mig_want = 0        # largest message of the last migration; all-reduced in pm() to resize _mig_cap
        self.redo_count = 0       # migrations of this rank that had to repeat a message (capacity overflow)
        self.phase_marks = None   # bench.py: list of (phase name, CUDA event) recorded at phase boundaries

    def _warm_host_ops(self):
        """The migration bookkeeping uses a handful of torch index ops on small tensors; CUDA loads each kernel's
        module lazily on first use (tens of ms), so run every one of them once here rather than inside a step."""
        dev = self._device()
        if dev.type != "cuda":
            return
        holes = torch.arange(8, dtype=torch.int64, device=dev)
        low = holes[holes < 4]
        tail = torch.ones((4,), dtype=torch.bool, device=dev)
        tail[holes[holes >= 4] - 4] = False
        src = torch.nonzero(tail).flatten() + 4
        torch.cat([low, src, torch.arange(8, 12, dtype=torch.int64, device=dev)]).cpu()
        rec = torch.zeros((4, REC), dtype=torch.float32, device=dev)
        torch.cat([rec[1:3], rec[0:1]]).contiguous()
        torch.cat([holes[:3], rec[0, :2].contiguous().view(torch.int64)]).cpu()
        m = torch.zeros((2,), dtype=torch.float32, device=dev)
        torch.cat([m, torch.tensor([1.0], dtype=torch.float32, device=dev)]).cpu()

    def _mark(self, name):
        if self.phase_marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.phase_marks.append((name, ev))

    # -- particle storage
    @property
    def position(self):
        return self.pos[:self.np]

    @property
    def velocity(self):
        return self.vel[:self.np]

    @property
    def acceleration(self):
        return self.acc[:self.np]

    @property
    def particle_ids(self):
        return self.ids[:self.np]

    def _ensure_capacity(self, n):
        cap = 0 if self.pos is None else self.pos.shape[0]
        if n <= cap and self.pos is not None:
            return
        new_cap = int(n * self.capacity_factor) + 1024

        def grow(a, width, dtype):
            shape = (new_cap, width) if width else (new_cap,)
            b = torch.empty(shape, dtype=dtype, device=self._device())
            if a is not None and self.np:
                b[:self.np] = a[:self.np]
            return b

        self.pos = grow(self.pos, 3, torch.float32)
        self.vel = grow(self.vel, 3, torch.float32)
        self.acc = grow(self.acc, 3, torch.float32)
        self.ids = grow(self.ids, 0, torch.int64)

    def _device(self):
        return getattr(self.ops, "dev", torch.device("cpu"))

    def set_particles(self, position, velocity, ids):
        """Adopt particles (any of them may lie outside the slab: migrate() is called).  Arrays are [n,3] / [n]
        tensors on this rank's device; ids are the global rows of the particles (kept through migrations so that
        results can be put back in the reference's order)."""
        n = position.shape[0]
        self._sorted_rows = None
        self.np = 0
        self.pos = self.vel = self.acc = self.ids = None
        self._ensure_capacity(n)
        self.pos[:n] = position
        self.vel[:n] = velocity
        self.ids[:n] = ids
        self.acc[:n] = 0
        self.np = n
        self.migrate()

    # -- migration
    def migrate(self, detected=None, neighbours_only=False):
        """Send every particle that left the slab to its owner and take in the arrivals (O(migrants) row moves).
        detected = (counts[P + 1], rows) from ops.kick_drift_wrap_detect skips the two scans over all positions.
        neighbours_only: after a leapfrog drift the Courant condition keeps every particle within one cell of where
        it was (integration.py:324-326), so leavers can only go to the adjacent slabs and the exchange is two
        point-to-point messages instead of an all-to-all."""
        ops, comm, me = self.ops, self.comm, self.rank
        n = self.np
        rows = None
        if detected is not None:
            counts = detected[0].cpu().tolist()
            if counts[self.P] <= detected[1].numel():
                rows = detected[1][:counts[self.P]]
            counts = counts[:self.P]
        else:
            counts = ops.count_owners(self.pos[:n]).cpu().tolist()
        send_counts = [int(c) for c in counts]
        send_counts[me] = 0
        nout = sum(send_counts)
        left, right = (me - 1) % self.P, (me + 1) % self.P
        if neighbours_only and self.P > 1:
            if any(c and d not in (left, right) for d, c in enumerate(send_counts)):
                raise RuntimeError("a particle crossed more than one slab in a single step: the Courant condition "
                                   "is violated (Courant_factor > slab thickness in cells?)")
            if left == right:     # P == 2: one peer; its records are ordered [to_left | to_right] by destination
                n_l, n_r = send_counts[left], 0
            else:
                n_l, n_r = send_counts[left], send_counts[right]
            from_l, from_r = comm.neighbor_counts(n_l, n_r)
            recv_counts = None
            nin = from_l + from_r
        else:
            recv_counts = comm.exchange_counts(send_counts)
            nin = sum(recv_counts)
        self.migrated_last = (nout, nin)
        if comm.size == 1:
            return  # periodic wrap keeps every particle in the only slab
        dev = self._device()
        if nout:
            offsets = torch.tensor(np.concatenate([[0], np.cumsum(send_counts)[:-1]]).astype(np.int64), device=dev)
            if rows is not None:
                sendbuf, holes = ops.pack_rows(self.pos[:n], self.vel[:n], self.ids[:n], rows, offsets, nout)
            else:
                sendbuf, holes = ops.pack_leavers(self.pos[:n], self.vel[:n], self.ids[:n], offsets, nout)
        else:
            sendbuf = torch.empty((0, REC), dtype=torch.float32, device=dev)
            holes = torch.empty((0,), dtype=torch.int64, device=dev)
        if recv_counts is None:
            # records are ordered by destination rank: the part for the lower-numbered neighbour comes first
            first, second = (left, right) if left < right else (right, left)
            a, b = sendbuf[:send_counts[first]], sendbuf[send_counts[first]:]
            to_l, to_r = (a, b) if first == left else (b, a)
            if left == right:
                to_l, to_r = sendbuf, sendbuf[:0]
            # message sizes differ from rank to rank here: point-to-point only (the peer-memory mailbox grows
            # collectively, which ranks with different sizes would enter at different times -- ADVICE r1)
            got_l, got_r = comm.pair_exchange(to_l, to_r, from_l, from_r)
            recvbuf = torch.cat([got_l, got_r]) if (from_l and from_r) else (got_l if from_l else got_r)
        else:
            # every rank enters the all-to-all, even with nothing to send or receive
            recvbuf = comm.all_to_all_v(sendbuf, send_counts, recv_counts)
        if nout == 0 and nin == 0:
            return
        n_new = n - nout + nin
        self._ensure_capacity(n_new)
        if nin >= nout:
            rows = torch.cat([holes, torch.arange(n, n_new, dtype=torch.int64, device=dev)])
            ops.unpack_rows(recvbuf, rows, self.pos, self.vel, self.ids)
        else:
            low = holes[holes < n_new]           # holes that stay inside the shrunken array
            ops.unpack_rows(recvbuf, low[:nin].contiguous(), self.pos, self.vel, self.ids)
            dst = low[nin:].contiguous()
            if dst.numel():
                # stayers living in the tail [n_new, n) move into the remaining holes
                tail = torch.ones((n - n_new,), dtype=torch.bool, device=dev)
                tail[holes[holes >= n_new] - n_new] = False
                src = torch.nonzero(tail).flatten() + n_new
                assert src.numel() == dst.numel(), (src.numel(), dst.numel())
                ops.move_rows(src.contiguous(), dst, self.pos, self.vel, self.ids)
        self.np = n_new

    def migrate_neighbours(self, detected=None):
        """migrate() for the step after a leapfrog drift, in ONE exchange with each neighbouring slab and one host
        synchronisation: the Courant condition (integration.py:324-326) keeps every particle within one cell of
        where it was, so leavers only go to the adjacent slabs.  Both directions use fixed-capacity buffers whose
        first record carries the true count; a count above the capacity (known to both ends of that message)
        makes that pair repeat the message with a larger buffer."""
        ops, comm, me, P = self.ops, self.comm, self.rank, self.P
        if P == 1:
            self.migrated_last = (0, 0)
            return
        n = self.np
        dev = self._device()
        if self._mig_cap is None:
            self._mig_cap = 2 * self.N * self.N + 4096   # the same on every rank: two planes' worth of particles
        cap = self._mig_cap
        sendL, sendR, holesL, holesR, status = ops.pack_fixed(self.pos[:n], self.vel[:n], self.ids[:n], detected, cap)
        if P == 2:   # one peer: everything travels "left"; nothing is sent or expected on the right
            gotL, gotR = comm.neighbor_exchange(sendL, sendR[:0], 0, cap + 1)
        else:
            gotL, gotR = comm.neighbor_exchange(sendL, sendR, cap + 1, cap + 1)
        hdr = [status]
        hdr.append(gotL[0, :2].contiguous().view(torch.int64) if gotL.shape[0] else status[:1] * 0)
        hdr.append(gotR[0, :2].contiguous().view(torch.int64) if gotR.shape[0] else status[:1] * 0)
        outL, outR, bad, inL, inR = (int(v) for v in torch.cat(hdr).cpu().tolist())   # the step's one sync
        if bad:
            raise RuntimeError("a particle crossed more than one slab in a single step: the Courant condition "
                               "is violated (Courant_factor > slab thickness in cells?)")
        # the capacity has to stay identical on all ranks: the wish travels with the step's all-reduce(max) (pm())
        self._mig_want = max(outL, outR, inL, inR)
        if self._mig_want > cap:
            # rare: the messages that did not fit are repeated with their exact size.  BOTH ends of a message know its
            # true count (the header), so the decision is taken per message and only the pairs concerned communicate,
            # point to point -- a rank whose messages all fitted is not involved and goes on.
            return self._migrate_redo(detected, cap, outL, outR, inL, inR, gotL, gotR)
        holes = torch.cat([holesL[:outL], holesR[:outR]])
        recvbuf = torch.cat([gotL[1:1 + inL], gotR[1:1 + inR]])
        self._apply_migration(n, holes, recvbuf, outL + outR, inL + inR)

    def _migrate_redo(self, detected, cap, outL, outR, inL, inR, gotL, gotR):
        """Overflow path of migrate_neighbours.  All leavers are packed again with exact sizes (pack_leavers); a
        direction whose count exceeded `cap` is sent again in full, point to point; a direction that fitted keeps
        the records that already arrived and is NOT sent again."""
        ops, comm, me, P = self.ops, self.comm, self.rank, self.P
        self.redo_count += 1
        n = self.np
        dev = self._device()
        left, right = (me - 1) % P, (me + 1) % P
        send_counts = [0] * P
        send_counts[left] += outL
        send_counts[right] += outR
        nout = outL + outR
        if nout:
            offsets = torch.tensor(np.concatenate([[0], np.cumsum(send_counts)[:-1]]).astype(np.int64), device=dev)
            sendbuf, holes = ops.pack_leavers(self.pos[:n], self.vel[:n], self.ids[:n], offsets, nout)
        else:
            sendbuf = torch.empty((0, REC), dtype=torch.float32, device=dev)
            holes = torch.empty((0,), dtype=torch.int64, device=dev)
        first, second = (left, right) if left < right else (right, left)
        a, b = sendbuf[:send_counts[first]], sendbuf[send_counts[first]:]
        to_l, to_r = (a, b) if first == left else (b, a)
        if left == right:
            to_l, to_r = sendbuf, sendbuf[:0]
        again_l, again_r = to_l if to_l.shape[0] > cap else to_l[:0], to_r if to_r.shape[0] > cap else to_r[:0]
        need_l, need_r = (inL if inL > cap else 0), (inR if inR > cap else 0)
        red_l, red_r = comm.pair_exchange(again_l, again_r, need_l, need_r)
        from_l = red_l if need_l else gotL[1:1 + inL]
        from_r = red_r if need_r else gotR[1:1 + inR]
        self._apply_migration(n, holes, torch.cat([from_l, from_r]), nout, inL + inR)

    def _apply_migration(self, n, holes, recvbuf, nout, nin):
        """Fill the holes left by the leavers with the arrivals (and with tail particles if more left than came)."""
        ops = self.ops
        dev = self._device()
        self.migrated_last = (nout, nin)
        if nout == 0 and nin == 0:
            return
        n_new = n - nout + nin
        self._ensure_capacity(n_new)
        if nin >= nout:
            rows = torch.cat([holes, torch.arange(n, n_new, dtype=torch.int64, device=dev)])
            ops.unpack_rows(recvbuf.contiguous(), rows, self.pos, self.vel, self.ids)
        else:
            low = holes[holes < n_new]           # holes that stay inside the shrunken array
            ops.unpack_rows(recvbuf.contiguous(), low[:nin].contiguous(), self.pos, self.vel, self.ids)
            dst = low[nin:].contiguous()
            if dst.numel():
                # stayers living in the tail [n_new, n) move into the remaining holes
                tail = torch.ones((n - n_new,), dtype=torch.bool, device=dev)
                tail[holes[holes >= n_new] - n_new] = False
                src = torch.nonzero(tail).flatten() + n_new
                assert src.numel() == dst.numel(), (src.numel(), dst.numel())
                ops.move_rows(src.contiguous(), dst, self.pos, self.vel, self.ids)
        self.np = n_new

    def _peer_buffers(self):
        """Two symmetric spectrum buffers (receive sides of the two transposes), or None."""
        if self._peer is None:
            fn = getattr(self.comm, "symmetric_buffers", None)
            bufs = None
            if fn is not None and isinstance(self.ops, CudaOps) and self.P > 1:
                bufs = fn(self.nxl * self.N * (self.N // 2 + 1) * 2, 2)
            self._peer = bufs if bufs else False
        return self._peer or None

    # -- Poisson solve on the slab
    def fft_poisson(self, rhs_planes, out_planes, param):
        """solver.fft (solver.py:452-523) with the transposed slab FFT.  rhs_planes [nxl,N,N] is consumed;
        the potential is written to out_planes [nxl,N,N] (may be a view into a ghosted array)."""
        ops, comm = self.ops, self.comm
        name = param["linear_newton_solver"].casefold()
        p = int(param["MAS_index"])
        if name == "fft":
            kind, pp = (_lib.GREEN_PLAIN, 0) if p == 0 else (_lib.GREEN_COMPENSATED, p)
        elif name == "fft_7pt":
            kind, pp = _lib.GREEN_7PT, 0
        else:
            raise NotImplementedError(f"slab path: linear_newton_solver={param['linear_newton_solver']!r}, "
                                      "should be 'fft' or 'fft_7pt'")
        peer = self._peer_buffers()
        if peer is not None:
            # transposes as one kernel each over NVLink peer memory (csrc/slab.cu: slab_put_kernel)
            (A, hA), (B, hB) = peer
            a = ops.spectrum_buffer()
            ops.fft2d_r2c(rhs_planes, a)                        # [nxl][N][nz]
            ops.transpose_put(a, hA.buffer_ptrs_dev, True)      # -> every rank's A = [N][nyl][nz]
            del a
            hA.barrier(channel=0)
            self._x_solve(A, kind, pp, param)
            ops.transpose_put(A, hB.buffer_ptrs_dev, False)     # -> every rank's B = [nxl][N][nz]
            hB.barrier(channel=0)
            ops.fft2d_c2r(B, out_planes)
            return
        a = ops.spectrum_buffer()
        b = ops.spectrum_buffer()
        ops.fft2d_r2c(rhs_planes, a)            # [nxl][N][nz]
        ops.yblocks(a, b, True)                 # [P][nxl][nyl][nz]
        comm.all_to_all_equal(b.view(self.P, -1), a.view(self.P, -1))   # [N][nyl][nz]
        self._x_solve(a, kind, pp, param)
        comm.all_to_all_equal(a.view(self.P, -1), b.view(self.P, -1))   # [P(y block)][nxl][nyl][nz]
        ops.yblocks(b, a, False)                # [nxl][N][nz]
        ops.fft2d_c2r(a, out_planes)

    def _x_solve(self, spec_t, kind, pp, param):
        """the x part of the solve on the transposed spectrum [N][nyl][nz]: forward FFT along x, (P(k)), Green's
        function with the 1/N^3 of the inverse transform, backward FFT along x -- one kernel unless a P(k) is wanted"""
        ops = self.ops
        scale = 1.0 / float(self.N) ** 3
        want_pk = bool(param.get("save_pk", False))
        fused = getattr(ops, "xfft_green", None)
        if not want_pk and fused is not None and fused(spec_t, kind, pp, scale):
            return
        ops.fft_x(spec_t, False)
        if want_pk:
            self._write_pk(spec_t, param)
        ops.green(spec_t, kind, pp, scale)
        ops.fft_x(spec_t, True)

    def _write_pk(self, spec_t, param, from_density=False):
        """fourier.fourier_grid_to_Pk (fourier.py:22-100) + the scaling of solver.fft (solver.py:500-506) on the
        transposed spectrum: per-rank bins, summed over ranks, written by rank 0."""
        from . import iostream
        bins = self.comm.allreduce_sum_(self.ops.pk_bins(spec_t, param["MAS_index"]))
        if self.rank != 0:
            return
        b = bins.cpu().numpy()
        N = self.N
        kmax = int(2 * (N // 2) / 3)
        nm = b[2, 1:kmax]
        with np.errstate(invalid="ignore", divide="ignore"):
            k = (b[0, 1:kmax] / nm).astype(np.float32)
            pk = (b[1, 1:kmax] / nm).astype(np.float32)
        pk *= (param["boxlen"] / N ** 2) ** 3
        if not from_density:
            pk /= (1.5 * param["aexp"] * param["Om_m"]) ** 2 * param["parametrized_mu_z"] ** 2
        k *= 2 * np.pi / param["boxlen"]
        iostream.write_power_spectrum_to_ascii_file(k, pk, nm.astype(np.float32), param)

    def _pk_from_density(self, density_planes, param):
        """P(k) of the density when the solver is not spectral (solver.py:130-138): forward transposed FFT only, through
        the all-to-all form of the transpose (output steps only)."""
        ops = self.ops
        a = ops.spectrum_buffer()
        b = ops.spectrum_buffer()
        ops.fft2d_r2c(density_planes, a)
        ops.yblocks(a, b, True)
        self.comm.all_to_all_equal(b.view(self.P, -1), a.view(self.P, -1))
        ops.fft_x(a, False)
        self._write_pk(a, param, from_density=True)

    @staticmethod
    def _fr_force_factor(param):
        """(0.5 c^2 (-f_R(a)), fR_n): the fifth force is -grad(phi + f u^(n+1)) (solver.py:166-179)"""
        from .solver import _fr_background
        _, fR_a, c2 = _fr_background(param)
        return np.float32(0.5 * (-fR_a) * c2), int(param["fR_n"])

    def _scalaron(self, density_planes, G, param):
        """solver.get_additional_field for f(R) (solver.py:326-359): density term, q, first guess (the previous scalaron,
        or cubic / quartic.initialise_potential), multigrid.FAS on the slab.  Returns the scalaron with G ghost planes
        per side, filled (the force stencil needs them)."""
        from .slab_multigrid import SlabMultigrid
        from .solver import _fr_background
        if self._mg is None:
            self._mg = SlabMultigrid(self.comm, self.ops, self.N)
        ops, N, nxl = self.ops, self.N, self.nxl
        kind = SlabMultigrid.fr_kind(param)
        Rbar, fR_a, c2 = _fr_background(param)
        a = param["aexp"]
        f1 = np.float32(a * param["Om_m"] / (c2 * 6)) / (-fR_a)
        f2 = np.float32(Rbar / 3 * a ** 4 - param["Om_m"] * a) / (6 * c2) / (-fR_a)
        dens_term = density_planes.clone()
        ops.affine(dens_term, f1, f2)
        q = np.float32(-a ** 4 * Rbar / (18 * c2)) / (-fR_a)
        param["fR_q"] = q
        param["compute_additional_field"] = True
        u_g = torch.empty((nxl + 2 * G, N, N), dtype=torch.float32, device=density_planes.device)
        xg = u_g[G - 1:G + nxl + 1]
        own = u_g[G:G + nxl]
        if self.additional_field is None:
            ops.mg_init_fr(dens_term, np.float32(q), nxl, N, kind, own)
        else:
            own.copy_(self.additional_field)
        self._mg.fas(xg, dens_term, param)
        param["compute_additional_field"] = False
        self.additional_field = own
        from_left, from_right = self.comm.exchange_planes(u_g[G:2 * G], u_g[nxl:nxl + G])
        u_g[:G] = from_left
        u_g[nxl + G:] = from_right
        self._mark("scalaron FAS")
        return u_g

    def _poisson(self, rhs_planes, out_g, G, param, tables, previous):
        """one linear Poisson solve into the owned planes out_g[G : G + nxl] of a ghosted array"""
        if param["linear_newton_solver"].casefold() == "multigrid":
            self.multigrid_poisson(rhs_planes, out_g, G, param, tables, previous)
        else:
            self.fft_poisson(rhs_planes, out_g[G:G + self.nxl], param)

    def multigrid_poisson(self, rhs_planes, phi_g, G, param, tables, previous):
        """solver.initialise_potential + multigrid.linear (solver.py:218-282, multigrid.py:23-83) on the slab.  The
        solve runs in place inside phi_g (planes G - 1 .. G + nxl are the multigrid's ghosted array).  ``previous``
        (the same field at the last step, owned planes) is the first guess from the second call on -- rescaled by
        a D1(a) for the potential, as it is for an additional field."""
        from .slab_multigrid import SlabMultigrid
        if self._mg is None:
            self._mg = SlabMultigrid(self.comm, self.ops, self.N)
        nxl = self.nxl
        xg = phi_g[G - 1:G + nxl + 1]
        own = xg[1:nxl + 1]
        if previous is None:
            logging.info("Assign potential from density field")
            own.copy_(rhs_planes)
            self.ops.affine(own, SlabMultigrid.first_guess_factor(self.N), 0.0)
        else:
            logging.info("Rescale potential from previous step for Newtonian potential")
            own.copy_(previous)
            if not param["compute_additional_field"]:
                if tables is None:
                    raise ValueError("Slab.pm(param, tables=...) is needed from the second multigrid solve on: the "
                                     "previous potential is rescaled with the growth table (solver.py:274-281)")
                scaling = (param["aexp"] * tables[3](np.log(param["aexp"]))
                           / (param["aexp_old"] * tables[3](np.log(param["aexp_old"]))))
                self.ops.affine(own, np.float32(scaling), 0.0)
        self._mg.linear(xg, rhs_planes, param)

    # -- solver.pm on the slab
    def pm(self, param, kick=None, tables=None):
        """solver.pm (solver.py:30-215): theory newton / parametrized (linear_newton_solver = fft, fft_7pt or multigrid),
        mond (QUMOND: fft_7pt or multigrid, two solves around the nu-weighted source) and fr (scalaron by FAS multigrid,
        then the Newtonian solve; the fifth force enters through the gradient).  Fills self.acc[:np]; with
        kick = half_dt also applies the second half-kick to the velocities.  ``tables`` (the cosmology interpolators)
        are only needed by the multigrid warm start (solver.py:274-281).  Returns the device tensor [max|a|, max|v|]
        already reduced over ranks."""
        ops, comm, N, nxl = self.ops, self.comm, self.N, self.nxl
        theory = param["theory"].casefold()
        if theory not in ("newton", "parametrized", "mond", "fr"):
            raise NotImplementedError(f"{param['theory']=}, should be 'newton', 'fr', 'parametrized' or 'mond'")
        solver_name = param["linear_newton_solver"].casefold()
        if theory == "mond" and solver_name not in ("multigrid", "fft_7pt"):
            raise NotImplementedError(f"{param['linear_newton_solver']=}, should be 'multigrid' or 'fft_7pt'")
        ms = param["mass_scheme"].casefold()
        if ms not in _SCHEMES:
            raise NotImplementedError(f"{param['mass_scheme']=}, should be 'CIC' or 'TSC'")
        scheme, param["MAS_index"] = _SCHEMES[ms]
        order = param["gradient_stencil_order"]
        if order not in _REACH:
            raise NotImplementedError(f"Unsupported: gradient_order={order}")
        if theory == "parametrized":
            a = param["aexp"]
            evo = a ** (-3 * (1 + param["w0"] + param["wa"])) * np.exp(-3 * param["wa"] * (1 - a))
            olz = (param["Om_lambda"] * evo / (param["Om_m"] * a ** (-3) + param["Om_r"] * a ** (-4)
                                               + param["Om_lambda"] * evo))
            param["parametrized_mu_z"] = np.float32(1 + param["parametrized_mu0"] * olz / param["Om_lambda"])
        else:
            param["parametrized_mu_z"] = np.float32(1)
        sps = str(param.get("save_power_spectrum", "no")).casefold()
        if sps not in ("yes", "z_out", "no"):
            raise NotImplementedError(f"SAVE_POWER_SPECTRUM={sps!r}, should be 'yes', 'z_out' or 'no'")
        param["save_pk"] = sps == "yes" or (sps == "z_out" and bool(param["write_snapshot"]))
        n = self.np
        self._mark("migrate")
        sorted_layout = bool(getattr(ops, "sorted_layout", False)) and not os.environ.get("PSC_SLAB_SHADOW_BINNING")
        if sorted_layout:
            # the rank's particle arrays are sorted into bin order (the acceleration buffer, dead until the
            # interpolation rewrites it, and the reorder's spare take the sorted copies): deposit and interpolation
            # then read position / velocity in place
            self._sort_by_bin()
            binned = n
            rho = ops.deposit_sorted(self.pos[:n], scheme)
        else:
            binned = ops.bin(self.pos[:n])
            rho = ops.deposit(binned, scheme)                      # [nxl + 2, N, N] raw sums
        self._mark("bin+deposit")
        from_left, from_right = comm.exchange_planes(rho[0:1], rho[nxl + 1:nxl + 2])
        rho[1] += from_left[0]         # the left neighbour's plane nxl + 1 is my first owned plane
        rho[nxl] += from_right[0]      # the right neighbour's plane 0 is my last owned plane
        rhs = rho[1:nxl + 1]
        conversion = np.float32(N ** 3 / param["npart"]) if N ** 3 != param["npart"] else np.float32(1)
        if conversion != 1:
            ops.affine(rhs, conversion, 0.0)
        use_multigrid = solver_name == "multigrid"
        if use_multigrid and param["save_pk"]:
            self._pk_from_density(rhs, param)      # solver.py:130-138: P(k) of the density, not of the RHS
        G = 1 + _REACH[order]
        if nxl < G:
            raise ValueError(f"slab of {nxl} planes is thinner than the {G} ghost planes the stencils need")
        u_g, half_c2, fr_n = None, 0.0, 0
        if theory == "fr":
            u_g, half_c2, fr_n = self._scalaron(rhs, G, param), *self._fr_force_factor(param)
        f1 = np.float32(1.5 * param["aexp"] * param["Om_m"] * param["parametrized_mu_z"])
        ops.affine(rhs, f1, -f1)
        self._mark("density ghosts+rhs")
        phi_g = torch.empty((nxl + 2 * G, N, N), dtype=torch.float32, device=rho.device)
        if theory == "mond":
            # QUMOND (solver.py:104-127, 365-378, 404-431): Newtonian potential phi_N (the "additional field") ->
            # source div(nu(|grad phi_N| / g0) grad phi_N), written over the density buffer -> second solve
            param["compute_additional_field"] = True
            phiN_g = torch.empty((nxl + 2, N, N), dtype=torch.float32, device=rho.device)
            self._poisson(rhs, phiN_g, 1, param, tables, self.additional_field)
            param["compute_additional_field"] = False
            from_left, from_right = comm.exchange_planes(phiN_g[1:2], phiN_g[nxl:nxl + 1])
            phiN_g[0] = from_left[0]
            phiN_g[nxl + 1] = from_right[0]
            g0 = (param["mond_g0"] * 1e-3 * 1e-10 * param["unit_t"] ** 2 / param["unit_l"]
                  * param["aexp"] ** (1 + param["mond_scale_factor_exponent"]))
            fn = param["mond_function"].casefold()
            if fn not in _lib.MOND_FN:
                raise NotImplementedError(f"MOND_FUNCTION={fn!r}, should be 'simple', 'n', 'beta', 'gamma' or 'delta'")
            ops.mond_rhs(phiN_g, rhs, nxl, N, g0, _lib.MOND_FN[fn], 1.0 if fn == "simple" else param["mond_alpha"])
            self.additional_field = phiN_g[1:nxl + 1]
            save_pk, param["save_pk"] = param["save_pk"], False     # no P(k) of the MOND source (solver.py:476-478)
            self._poisson(rhs, phi_g, G, param, tables, self.potential)
            param["save_pk"] = save_pk
        else:
            param["compute_additional_field"] = False    # solver.py:121 (no additional field)
            self._poisson(rhs, phi_g, G, param, tables, self.potential)
        del rho, rhs
        self._mark("multigrid solve" if use_multigrid else "fft solve (2 transposes)")
        from_left, from_right = comm.exchange_planes(phi_g[G:2 * G], phi_g[nxl:nxl + G])
        phi_g[:G] = from_left          # the left neighbour's last G owned planes
        phi_g[nxl + G:] = from_right   # the right neighbour's first G owned planes
        half_dt = 0.0 if kick is None else kick
        self._mark("potential ghosts")
        vel_k = self.vel[:n] if kick is not None else None
        extra = (u_g, half_c2, fr_n) if u_g is not None else ()
        if sorted_layout:
            mx = ops.interp_kick_phi_sorted(phi_g, G, order, self.pos[:n], vel_k, self.acc[:n], scheme, half_dt, *extra)
        else:
            mx = ops.interp_kick_phi(phi_g, G, order, binned, vel_k, self.acc[:n], scheme, half_dt, *extra)
        self.potential = phi_g[G:G + nxl]
        if kick is None:
            mx[1] = ops.max_abs(self.vel[:n])[0]
        self._mark("gradient+interp+kick")
        # (torch.full, not torch.tensor: a host list would be copied from pageable memory, which blocks the host until
        # the stream has drained -- the all-reduce and the read below are then enqueued behind the interpolation
        # kernel instead of after it)
        mx = torch.cat([mx, torch.full((1,), float(self._mig_want), dtype=torch.float32, device=mx.device)])
        comm.allreduce_max_(mx)
        m = mx.cpu().numpy()
        self._mark("allreduce max")
        self.max_acc, self.max_vel = np.float32(m[0]), np.float32(m[1])
        if self._mig_cap is not None:
            # grow at once after an overflow; shrink only when the buffers (exchanged in full every step) have been far
            # larger than the traffic for a while: the yardstick is a slowly decaying maximum of the per-step traffic,
            # so one quiet step (a short step clamped to a snapshot time) does not halve the capacity just before
            # the next ordinary step needs it.  Identical on every rank: m[2] is the all-reduced maximum.
            self._mig_peak = max(float(m[2]), 0.97 * getattr(self, "_mig_peak", 0.0))
            if m[2] > self._mig_cap or 8 * self._mig_peak + 4096 < self._mig_cap:
                self._mig_cap = 2 * int(self._mig_peak) + 4096
        return mx[:2]

    # -- integration.integrate / leapfrog on the slab
    def leapfrog(self, dt, tables, param):
        """integration.leapfrog (integration.py:192-264)"""
        from . import utils
        n = self.np
        half_dt = np.float32(0.5 * dt)
        dt_is_f64 = 0 if isinstance(dt, np.float32) else 1
        self._mark("start")
        detected = None
        if self.P > 1:
            detected = self.ops.kick_drift_wrap_detect(self.pos[:n], self.vel[:n], self.acc[:n], half_dt, dt,
                                                       dt_is_f64)
        else:
            self.ops.kick_drift_wrap(self.pos[:n], self.vel[:n], self.acc[:n], half_dt, dt, dt_is_f64)
        param["t"] += dt
        param["aexp_old"] = param["aexp"]
        param["aexp"] = np.exp(tables[0](param["t"]))
        logging.info(f"{param['t']=} {param['aexp']=}")
        utils.set_units(param)
        self._mark("kick+drift+wrap")
        self.migrate_neighbours(detected)
        self.pm(param, kick=half_dt, tables=tables)

    def integrate(self, tables, param, t_snap_next=np.float32(0)):
        """integration.integrate (integration.py:17-118), leapfrog only"""
        if self.max_acc is None:
            raise RuntimeError("call pm() once before integrate() (the time step needs max|a|, max|v|)")
        if param["integrator"].casefold() != "leapfrog":
            raise NotImplementedError("slab path: integrator must be 'leapfrog'")
        if not (np.isfinite(self.max_acc) and np.isfinite(self.max_vel)):
            raise ValueError(f"math domain error: max|a| = {self.max_acc}, max|v| = {self.max_vel} (non-finite "
                             "particles after the force computation)")
        dx = np.float32(0.5 ** param["ncoarse"])
        cf = np.float32(param["Courant_factor"])
        dt1 = cf * np.sqrt(dx / self.max_acc)
        dt2 = cf * dx / self.max_vel
        aexp_factor = 1.0 + 0.01 * param["max_aexp_stepping"]
        dt3 = np.float32(tables[1](np.log(aexp_factor * param["aexp"])) - tables[1](np.log(param["aexp"])))
        dt = np.min([dt1, dt2, dt3])
        if (param["t"] + dt) > t_snap_next:
            dt = t_snap_next - param["t"]
            param["write_snapshot"] = True
        else:
            param["write_snapshot"] = False
        self.leapfrog(dt, tables, param)
        return dt

    def _sort_by_bin(self):
        """Sort the rank's particle rows (position, velocity, id) into bin order, out of place: positions go to the
        reorder's spare buffer, velocities to the acceleration buffer (its content is dead here: the kick has used it
        and the interpolation rewrites every row), ids to the spare id buffer; the buffers then swap roles."""
        n = self.np
        cap = self.pos.shape[0]
        dev = self._device()
        if self._spare3 is None or self._spare3.shape[0] != cap:
            self._spare3 = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            self._spare1 = torch.empty((cap,), dtype=torch.int64, device=dev)
        self.ops.sort_by_bin(self.pos[:n], self.vel[:n], self.ids[:n], self._spare3[:n], self.acc[:n], self._spare1[:n],
                             src_rows=self._sorted_rows)
        self.pos, self.vel, self.acc, self._spare3 = self._spare3, self.acc, self.vel, self.pos
        self.ids, self._spare1 = self._spare1, self.ids
        self._sorted_rows = n     # rows [0, n) are now in bin order (until something reorders them)

    def reorder(self):
        """utils.reorder_particles (utils.py:1019-1075) on the local particles (the Morton key's leading bits are
        x, so the local order is a contiguous piece of the global Morton order)."""
        n = self.np
        if n == 0:
            return
        self._sorted_rows = None      # the rows are about to be permuted
        idx = self.ops.morton_order(self.pos[:n])
        # gather into a persistent spare buffer and swap it in: no allocation, no copy back
        cap = self.pos.shape[0]
        if self._spare3 is None or self._spare3.shape[0] != cap:
            self._spare3 = torch.empty((cap, 3), dtype=torch.float32, device=self._device())
            self._spare1 = torch.empty((cap,), dtype=torch.int64, device=self._device())
        for name in ("pos", "vel", "acc"):
            a = getattr(self, name)
            self.ops.gather_rows(idx, a[:n], self._spare3[:n])
            setattr(self, name, self._spare3)
            self._spare3 = a
        torch.index_select(self.ids[:n], 0, idx, out=self._spare1[:n])
        self.ids, self._spare1 = self._spare1, self.ids

    # -- gathering results in the reference's order (tests / snapshots)
    def gather_to_root(self, npart_total):
        """All particles in global-id order on rank 0 (pos, vel, acc as CPU tensors); None elsewhere."""
        n = self.np
        payload = torch.cat([self.pos[:n].double(), self.vel[:n].double(), self.acc[:n].double(),
                             self.ids[:n].double().view(-1, 1)], dim=1)
        counts = [0] * self.P
        counts[0] = n
        recv_counts = self.comm.exchange_counts(counts)
        got = self.comm.all_to_all_v(payload, counts, recv_counts)
        if self.rank != 0:
            return None
        got = got.cpu()
        assert got.shape[0] == npart_total, (got.shape, npart_total)
        order = torch.argsort(got[:, 9].long())
        got = got[order]
        assert torch.equal(got[:, 9].long(), torch.arange(npart_total)), "particle ids lost or duplicated"
        return got[:, 0:3].float(), got[:, 3:6].float(), got[:, 6:9].float()


# ------------------------------------------------------------------------------------------ main.run on slabs
def _ics_per_slab(param, comm):
    """param['slab_ics']: 'slab' = every rank generates its own planes (initial_conditions.generate_slab), 'replicated'
    = every rank generates the whole set and adopts a share (meshes that fit one GPU), 'auto' (default) = per slab on
    more than one rank (with dealiased_ICS only when the 3/2 dealiasing grid splits over the ranks as well)"""
    how = str(param["slab_ics"]).casefold() if "slab_ics" in param.index else "auto"
    if how not in ("auto", "slab", "replicated"):
        raise NotImplementedError(f"{param['slab_ics']=}, should be 'slab', 'replicated' or 'auto'")
    if how != "auto":
        return how == "slab"
    if comm.size == 1:
        return False
    if bool(param["dealiased_ICS"]):
        n = int(round(float(param["npart"]) ** (1.0 / 3)))
        return (3 * n // 2) % comm.size == 0
    return True


def run(param, comm=None, initial_state=None, ops_factory=None):
    """`main.run` (main.py:30-156) on x-slabs: one call per rank (torchrun; or one thread per virtual rank with a
    ThreadComm).  Every rank derives the same background tables; the initial particles come from `initial_state`
    (global arrays, identical on every rank), from a snapshot number (`initial_conditions = i`: restart, every rank
    reads its own share of the files -- initial_conditions.py:79-107) or are generated from the parameter file: per slab
    (initial_conditions.generate_slab, every rank its own lattice planes: `slab_ics`, see _ics_per_slab) or identically
    on every rank (initial_conditions.generate: meshes that fit one GPU) -- each rank adopts its particles or a share
    and the first migration routes them to their slabs.  Snapshots: `slab_snapshots = gather` (default up to 256^3 particles) collects them
    on rank 0 in the reference's particle order and file layout; `slab_snapshots = parts` (default above) lets every
    rank write its own slab (iostream.write_snapshot_slab_part), so that no rank ever holds the global arrays.
    Theories and solvers: what Slab.pm
    supports (newton / parametrized / mond / fr; fft, fft_7pt, multigrid), leapfrog.  Returns (position, velocity) of the final state on rank 0
    (CPU tensors, reference order), None elsewhere.  ops_factory(N, P, rank) replaces the CUDA kernels (tests only)."""
    import pandas as pd
    from . import cosmotable, iostream, utils
    from . import main as _main
    comm = comm if comm is not None else default_comm()
    if isinstance(param, dict):
        param = pd.Series(param)
    elif not isinstance(param, pd.Series):
        raise ValueError(f"{type(param)=}, should be a dictionnary or a Pandas Series")
    if param["verbose"] not in (0, 1, 2):
        raise ValueError(f"{param['verbose']=}, should be 0, 1 or 2")
    root = comm.rank == 0
    param["write_snapshot"] = False
    extra = param["theory"].casefold()       # the reference's file-name tag (main.py:82-93)
    if extra == "fr":
        extra += f"{param['fR_logfR0']}_n{param['fR_n']}"
    elif extra == "mond":
        mond_function = param["mond_function"].casefold()
        extra += f"_g0_{param['mond_g0']}_exponent_{param['mond_scale_factor_exponent']}_{mond_function}"
        if "simple" != mond_function:
            extra += f"_{param['mond_alpha']}"
    elif extra == "parametrized":
        extra += f"_mu0_{param['parametrized_mu0']}"
    param["extra"] = f"{extra}_{param['linear_newton_solver']}_ncoarse{param['ncoarse']}"
    z_out = iostream.parse_z_out(param)
    if root:
        os.makedirs(f"{param['base']}/power", exist_ok=True)
        for i in range(len(z_out) + 1):
            os.makedirs(f"{param['base']}/output_{i:05d}", exist_ok=True)
    base_dir = param["base"]
    if not root:
        param = param.copy()
        param["base"] = ""          # only rank 0 writes tables, P(k) files and gathered snapshots
    tables = cosmotable.generate(param)
    param["aexp"] = 1.0 / (1 + param["z_start"])
    utils.set_units(param)
    if "nsteps" not in param.index:
        param["nsteps"] = 0
    N = 2 ** param["ncoarse"]
    S = Slab(N, comm=comm, ops=None if ops_factory is None else ops_factory(N, comm.size, comm.rank))
    dev = S._device()
    ic = param["initial_conditions"] if "initial_conditions" in param.index else None
    ics_snapshot = False
    if initial_state is None and isinstance(ic, (int, np.integer)):
        # restart: every saved parameter comes back (as main._initial_state), every rank reads its own share
        base = param["base"] if root else base_dir
        d = f"{base}/output_{int(ic):05d}"
        saved = iostream.read_param_file(f"{d}/param_{param['extra']}_{int(ic):05d}.txt")
        for key in saved.index:
            if key.casefold() not in ("nthreads", "base"):
                param[key] = saved[key]
        param["nsteps"], param["i_snap"] = int(param["nsteps"]), int(param["i_snap"])
        utils.set_units(param)
        p_, v_, i_ = iostream.read_snapshot_slab_parts(f"{d}/particles_{param['extra']}.parquet", comm.rank, comm.size)
        npart = int(param["npart"])
        S.set_particles(torch.from_numpy(p_).to(dev), torch.from_numpy(v_).to(dev), torch.from_numpy(i_).to(dev))
        del p_, v_, i_
    elif initial_state is None and _ics_per_slab(param, comm):
        # every rank generates the particles of its own lattice planes (initial_conditions.generate_slab): no rank
        # holds a global array, the white noise is the reference's, block by block
        from . import initial_conditions
        p_, v_, i_ = initial_conditions.generate_slab(param, tables, comm, device=dev)
        utils.set_units(param)
        param["t"] = tables[1](np.log(param["aexp"]))
        npart = int(param["npart"])
        S.set_particles(p_, v_, i_)
        del p_, v_, i_
        ics_snapshot = True
    else:
        if initial_state is None:
            from . import initial_conditions
            position, velocity = initial_conditions.generate(param, tables, write_snapshot=root, device=dev)
        else:
            position, velocity = initial_state
        utils.set_units(param)
        param["t"] = tables[1](np.log(param["aexp"]))
        position = torch.as_tensor(position, dtype=torch.float32).to(dev)
        velocity = torch.as_tensor(velocity, dtype=torch.float32).to(dev)
        npart = position.shape[0]
        ids = torch.arange(npart, dtype=torch.int64, device=dev)
        mine = slice(comm.rank, None, comm.size)
        S.set_particles(position[mine].contiguous(), velocity[mine].contiguous(), ids[mine].contiguous())
        del position, velocity, ids
    mode = str(param["slab_snapshots"]).casefold() if "slab_snapshots" in param.index else "auto"
    if mode not in ("auto", "gather", "parts"):
        raise NotImplementedError(f"{param['slab_snapshots']=}, should be 'gather', 'parts' or 'auto'")
    parts_mode = mode == "parts" or (mode == "auto" and npart > 256 ** 3)

    def snapshot():
        if parts_mode:
            p = param.copy()
            p["base"] = base_dir
            n = S.np
            iostream.write_snapshot_slab_part(S.pos[:n], S.vel[:n], S.ids[:n], p, comm.rank, root)
            comm.barrier()
        else:
            state = S.gather_to_root(npart)
            if root:
                iostream.write_snapshot_particles(state[0], state[1], param)
    if ics_snapshot:
        # the initial snapshot (initial_conditions.py:216-280), which the replicated generator writes itself
        if parts_mode:
            p = param.copy()
            p["base"], p["i_snap"] = base_dir, 0
            iostream.write_snapshot_slab_part(S.pos[:S.np], S.vel[:S.np], S.ids[:S.np], p, comm.rank, False)
            comm.barrier()
        else:
            state = S.gather_to_root(npart)
            if root:
                iostream.write_snapshot_particles_parquet(
                    f"{param['base']}/output_00000/particles_{param['extra']}.parquet", state[0], state[1])
        if root:
            param.to_csv(f"{param['base']}/output_00000/param_{param['extra']}.txt", sep="=", header=False)
    S.pm(param, tables=tables)
    aexp_out = np.sort(1.0 / (np.array(z_out) + 1))
    t_out = tables[1](np.log(aexp_out))
    param["i_snap"] = 1 if "i_snap" not in param.index else param["i_snap"] + 1
    # the time loop allocates a few thousand small Python objects per step: without this the cyclic collector makes a
    # full pass over the whole import-time heap (torch, pandas, scipy: ~3 ms) every couple of dozen steps
    import gc
    gc.collect()
    gc.freeze()
    while param["aexp"] < aexp_out[-1]:
        param["nsteps"] += 1
        S.integrate(tables, param, t_out[param["i_snap"] - 1])
        if (param["nsteps"] % param["n_reorder"]) == 0:
            S.reorder()
        if param["write_snapshot"]:
            snapshot()
            param["i_snap"] += 1
        logging.warning(f"{param['nsteps']=} {param['aexp']=} z = {1.0 / param['aexp'] - 1}")
    if parts_mode:       # the final state stays distributed: this rank's slab (position, velocity, ids)
        n = S.np
        out = (S.pos[:n].clone(), S.vel[:n].clone(), S.ids[:n].clone())
        S.ops.close()
        return out
    state = S.gather_to_root(npart)
    S.ops.close()
    return (state[0], state[1]) if root else None
