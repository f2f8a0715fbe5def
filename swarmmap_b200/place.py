"""Server-side cross-agent place recognition, sharded over GPUs (BASELINE.json config 5).

Role in the reference: AgentMediator::CheckOverlapCandidates walks every other agent's
KeyFrameDatabase on the CPU (code/src/AgentMediator.cc:140-202, KeyFrameDatabase::DetectLoopCandidates
code/src/KeyFrameDatabase.cc:74-185) and then matches candidates with SearchByBoW(KF,KF).  Here the
keyframe-descriptor database is partitioned by keyframe id range, one shard per rank/GPU; a query
keyframe's descriptors are brute-force matched against the local shard by the CUDA top-2 kernel
(swm_db_query_device), and the only exchange step is one all-gather of the fixed-size per-shard
candidate blocks (nq x k packed 64-bit keys + per-keyframe vote counts stay local), followed by a
deterministic merge (ascending key = distance, then global descriptor index) on every rank.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

__all__ = ["partition", "merge_topk", "PlaceShard", "unpack_keys"]

KEY_IDX_BITS = 48


def partition(n_keyframes, world_size):
    """Contiguous keyframe-id ranges per rank: [(first_kf, n_kf), ...] (remainder spread over the low ranks)."""
    base, rem = divmod(int(n_keyframes), int(world_size))
    out, first = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((first, n))
        first += n
    return out


def unpack_keys(keys):
    """packed keys (dist << 48 | global descriptor index) -> (dist, index) int64 arrays; ~0 -> (-1, -1)."""
    k = keys.to(torch.int64) if isinstance(keys, torch.Tensor) else torch.from_numpy(np.asarray(keys).astype(np.int64))
    none = k == -1
    d = (k >> KEY_IDX_BITS) & 0xFFFF
    i = k & ((1 << KEY_IDX_BITS) - 1)
    d = torch.where(none, torch.full_like(d, -1), d)
    i = torch.where(none, torch.full_like(i, -1), i)
    return d, i


def merge_topk(gathered, k):
    """gathered: (world, nq, k) packed keys -> (nq, k) smallest keys per query (unsigned order)."""
    w, nq, kk = gathered.shape
    flat = gathered.permute(1, 0, 2).reshape(nq, w * kk)
    # keys are < 2^63 except the all-ones "none" marker (-1 as int64): map it to int64 max for sorting
    big = torch.iinfo(torch.int64).max
    flat = torch.where(flat < 0, torch.full_like(flat, big), flat)
    out = torch.sort(flat, dim=1).values[:, :k]
    return torch.where(out == big, torch.full_like(out, -1), out)


class PlaceShard:
    """One rank's shard of the keyframe-descriptor database."""

    def __init__(self, desc, desc_per_kf, first_kf, device=0, group=None, local_topk=None):
        """desc: (n_desc, 32) uint8 (numpy or CUDA tensor) of this rank's keyframes, desc_per_kf each.
        local_topk: test hook replacing the CUDA kernel (CPU-only gloo tests of the exchange logic)."""
        self.group = group
        self.desc_per_kf = int(desc_per_kf)
        self.first_kf = int(first_kf)
        self._local_topk = local_topk
        self._h = None
        if local_topk is not None:
            self.desc = np.ascontiguousarray(desc, np.uint8)
            self.n_desc = len(self.desc)
            self.device = torch.device("cpu")
            return
        self._lib = _lib.load()
        self.device = torch.device("cuda", device)
        if isinstance(desc, torch.Tensor):
            self.d_desc = desc.to(self.device).contiguous()
        else:
            self.d_desc = torch.from_numpy(np.ascontiguousarray(desc, np.uint8)).to(self.device)
        self.n_desc = int(self.d_desc.shape[0])
        h = C.c_void_p()
        rc = self._lib.swm_db_create_device(device, self.d_desc.data_ptr(), self.n_desc, self.desc_per_kf,
                                            self.first_kf, C.byref(h))
        if rc != 0:
            raise _lib.SwmError(f"swm_db_create_device: {_lib.ERRORS.get(rc, rc)}")
        self._h = h

    def close(self):
        if self._h:
            self._lib.swm_db_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_kf(self):
        return (self.n_desc + self.desc_per_kf - 1) // self.desc_per_kf

    def query_local(self, q, k=2, th_votes=50):
        """q: (nq, 32) uint8.  Returns (keys (nq,k) int64 packed, votes (n_kf,) int32) for this shard."""
        if self._local_topk is not None:
            return self._local_topk(self, q, k, th_votes)
        dq = q.to(self.device).contiguous() if isinstance(q, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(q, np.uint8)).to(self.device)
        nq = int(dq.shape[0])
        keys = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        votes = torch.zeros(self.n_kf, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device)
        rc = self._lib.swm_db_query_device(self._h, dq.data_ptr(), nq, k, keys.data_ptr(), votes.data_ptr(),
                                           int(th_votes), C.c_void_p(st.cuda_stream))
        if rc != 0:
            raise _lib.SwmError(f"swm_db_query_device: {_lib.ERRORS.get(rc, rc)}")
        return keys, votes

    def query(self, q, k=2, th_votes=50):
        """Global top-k over all shards: local kernel -> all-gather of (nq,k) key blocks -> merge.
        Returns (keys (nq,k), local votes).  With no process group this is the 1-shard case."""
        keys, votes = self.query_local(q, k, th_votes)
        if not (dist.is_available() and dist.is_initialized()):
            return keys, votes
        world = dist.get_world_size(self.group)
        if world == 1:
            return keys, votes
        if self._h is None:  # CPU test hook (gloo): exchange + torch merge
            bucket = [torch.empty_like(keys) for _ in range(world)]
            dist.all_gather(bucket, keys, group=self.group)
            merged = merge_topk(torch.stack(bucket, 0), k)
            return merged, self.votes_from_global(merged, th_votes)
        # one all-gather of the (nq, k) key blocks into a (world, nq, k) buffer, then ONE merge kernel that also
        # casts this shard's votes from the global best matches
        nq = int(keys.shape[0])
        gathered = torch.empty((world, nq, k), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gathered, keys, group=self.group)
        merged = torch.empty_like(keys)
        votes = torch.zeros(self.n_kf, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device)
        rc = self._lib.swm_db_merge_gathered(self._h, gathered.data_ptr(), world, nq, k, merged.data_ptr(),
                                             votes.data_ptr(), int(th_votes), C.c_void_p(st.cuda_stream))
        if rc != 0:
            raise _lib.SwmError(f"swm_db_merge_gathered: {_lib.ERRORS.get(rc, rc)}")
        return merged, votes

    def query_sharded(self, q, comm, world, k=2, th_votes=50):
        """The same query through the C ABI alone (swm_db_query_sharded): local scan, ONE ncclAllGather issued by the
        library on the current stream, merge + votes -- what a C++ server calls.  comm: raw ncclComm_t (c_void_p),
        e.g. from nccl_comm_from_torch()."""
        dq = q.to(self.device).contiguous() if isinstance(q, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(q, np.uint8)).to(self.device)
        nq = int(dq.shape[0])
        merged = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        votes = torch.zeros(self.n_kf, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device)
        rc = self._lib.swm_db_query_sharded(self._h, comm, int(world), dq.data_ptr(), nq, k, merged.data_ptr(),
                                            votes.data_ptr(), int(th_votes), C.c_void_p(st.cuda_stream))
        if rc != 0:
            raise _lib.SwmError(f"swm_db_query_sharded: {_lib.ERRORS.get(rc, rc)}")
        return merged, votes

    def enable_peers(self, nq_max=2048, group=None, same_process=None):
        """Set up the peer-memory exchange (swm_db_peer_window / swm_db_peer_open).  Across processes (one rank per GPU,
        torch.distributed initialised): the 64-byte IPC handles of the windows are all-gathered through the group.
        same_process: a list of PlaceShard objects in rank order that share this process (tests on one GPU, a server
        driving several GPUs): call it on every shard after all of them exist."""
        if same_process is not None:
            world, rank = len(same_process), same_process.index(self)
            wins = (C.c_void_p * world)()
            for r, sh in enumerate(same_process):
                if getattr(sh, "_window", None) is None or sh._peer_world != world or sh._peer_nq_max < nq_max:
                    w = C.c_void_p()
                    rc = sh._lib.swm_db_peer_window(sh._h, world, int(nq_max), None, C.byref(w))
                    if rc != 0:
                        raise _lib.SwmError(f"swm_db_peer_window: {_lib.ERRORS.get(rc, rc)}")
                    sh._window, sh._peer_world, sh._peer_nq_max = w.value, world, int(nq_max)
                wins[r] = sh._window
            rc = self._lib.swm_db_peer_open(self._h, rank, None, wins)
            if rc != 0:
                raise _lib.SwmError(f"swm_db_peer_open: {_lib.ERRORS.get(rc, rc)}")
            return
        group = group if group is not None else self.group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        hdl = (C.c_uint8 * 64)()
        rc = self._lib.swm_db_peer_window(self._h, world, int(nq_max), hdl, None)
        if rc != 0:
            raise _lib.SwmError(f"swm_db_peer_window: {_lib.ERRORS.get(rc, rc)}")
        mine = torch.tensor(list(bytes(hdl)), dtype=torch.uint8, device=self.device)
        allh = torch.empty((world, 64), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        raw = bytes(allh.cpu().numpy().tobytes())
        rc = self._lib.swm_db_peer_open(self._h, rank, raw, None)
        if rc != 0:
            raise _lib.SwmError(f"swm_db_peer_open: {_lib.ERRORS.get(rc, rc)}")
        dist.barrier(group=group)  # every window is mapped everywhere before the first push

    def query_peers(self, q, k=2, th_votes=50):
        """Global top-k with the exchange over peer memory inside the merge kernel (swm_db_query_peers): a collective,
        every rank calls it with the same number of queries.  Returns (keys (nq,k), this shard's votes)."""
        dq = q.to(self.device).contiguous() if isinstance(q, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(q, np.uint8)).to(self.device)
        nq = int(dq.shape[0])
        merged = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        votes = torch.zeros(self.n_kf, dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device)
        rc = self._lib.swm_db_query_peers(self._h, dq.data_ptr(), nq, k, merged.data_ptr(), votes.data_ptr(),
                                          int(th_votes), C.c_void_p(st.cuda_stream))
        if rc != 0:
            raise _lib.SwmError(f"swm_db_query_peers: {_lib.ERRORS.get(rc, rc)}")
        return merged, votes

    def votes_from_global(self, merged, th_votes):
        """Per-keyframe votes of THIS shard from the merged result: a query votes for the keyframe owning
        its global best match when that distance is <= th_votes (TH_LOW); summed over ranks this equals
        the single-shard vote histogram."""
        d, i = unpack_keys(merged[:, 0])
        first = self.first_kf * self.desc_per_kf
        mask = (d >= 0) & (d <= th_votes) & (i >= first) & (i < first + self.n_desc)
        kf = torch.div(i[mask] - first, self.desc_per_kf, rounding_mode="floor")
        return torch.bincount(kf, minlength=self.n_kf).to(torch.int32)


class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def nccl_comm_from_torch(device, group=None):
    """A raw ncclComm_t over the ranks of the (initialised) torch.distributed group, created with the NCCL library
    the process already has loaded: rank 0 makes the unique id, torch broadcasts its 128 bytes, every rank calls
    ncclCommInitRank.  Returns (comm as c_void_p, nccl ctypes library); destroy with lib.ncclCommDestroy(comm)."""
    nccl = C.CDLL("libnccl.so.2")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = _NcclUniqueId()
    if rank == 0:
        nccl.ncclGetUniqueId.argtypes = [C.POINTER(_NcclUniqueId)]
        if nccl.ncclGetUniqueId(C.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    raw = C.string_at(C.byref(uid), 128) if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8, device=torch.device("cuda", device))
    dist.broadcast(t, src=0, group=group)
    C.memmove(C.byref(uid), bytes(t.cpu().tolist()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
    torch.cuda.set_device(device)
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank failed ({rc})")
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    return comm, nccl
