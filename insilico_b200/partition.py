"""Multi-GPU assembly: element blocks per GPU, owned row ranges, interface rows exchanged once per assembly.

SURVEY.md 8(e): the reference has no distributed path; the partition is B200-native.  One process per GPU
(torch.distributed).  Every rank assembles its own element block into a LOCAL CSR whose rows are the equations its
elements touch.  A row shared between two blocks is OWNED by one rank; the other rank's partial row ("ghost row") is
sent to the owner after the local assembly and added there (NCCL send/recv over NVLink on GPUs, gloo on CPU in the
tests).  The owner's pattern for its rows is complete because the neighbour's elements touching those rows are passed
to the engine as pattern-only halo elements (isl_mesh_set_owned).

The exchange plan is static: it is computed once after registerFields from the global (row, column) keys of the ghost
entries, so an assembly step sends packed values only.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def _active_before_plane(z, nz_planes, m):
    """number of ACTIVE (interior) nodes in planes [0, z) of a box whose whole boundary is constrained"""
    return int(np.clip(z - 1, 0, nz_planes - 2)) * m


def structured_laplace_slab(e1, e2, e3_total, rank, world, dirichlet_fun):
    """z-slab `rank` of `world` of the structured Q1 hex mesh e1 x e2 x e3_total on the unit cube, scalar field with
    Dirichlet data on the whole boundary (the flow of reference/04-heat/dirichlet.cpp:79-151).

    Local node set = planes [z0 - halo, z1]; owned elements first, then the halo layer below (pattern only).
    The shared plane z0 is owned by this rank, the plane z1 by rank+1 (ghost rows here) unless this is the last rank.
    Local equation numbers are the global ones minus `eqn_offset`."""
    from . import meshgen
    assert e3_total % world == 0
    e3 = e3_total // world
    z0, z1 = rank * e3, (rank + 1) * e3
    halo = 1 if rank > 0 else 0
    n1, n2 = e1 + 1, e2 + 1
    plane = n1 * n2
    # nodes of planes [z0-halo, z1]; elements of layers [z0-halo, z1)
    coords, conn_all, node_off = meshgen.unit_cube_hex(e1, e2, e3_total, k0=z0 - halo, k1=z1)
    layer = e1 * e2
    if halo:
        conn = np.concatenate([conn_all[layer:], conn_all[:layer]])  # owned first, halo last
    else:
        conn = conn_all
    n_owned = e3 * layer
    onb = meshgen.boundary_node_mask(coords)
    status = onb.astype(np.uint8)[:, None]
    presc = np.where(onb, dirichlet_fun(coords), 0.0)[:, None]
    eqn = np.full((len(coords), 1), -1, dtype=np.int64)
    act = ~onb
    eqn[act, 0] = np.arange(int(act.sum()), dtype=np.int64)
    m = (n1 - 2) * (n2 - 2)
    nzp = e3_total + 1
    eqn_offset = _active_before_plane(z0 - halo, nzp, m)
    n_local = int(act.sum())
    # row classes in local numbering (planes are contiguous in the numbering)
    own_lo = _active_before_plane(z0, nzp, m) - eqn_offset
    own_hi = _active_before_plane(z1 if rank < world - 1 else z1 + 1, nzp, m) - eqn_offset
    ghost_lo, ghost_hi = own_hi, _active_before_plane(z1 + 1, nzp, m) - eqn_offset
    return dict(coords=coords, conn=np.ascontiguousarray(conn), n_owned_elems=n_owned, n_obj=len(coords),
                elem_dof=np.ascontiguousarray(conn), eqn=eqn, status=status, presc=presc,
                values=np.zeros_like(presc), n_eqn_local=n_local, eqn_offset=eqn_offset,
                n_eqn_global=_active_before_plane(nzp, nzp, m), owned_rows=(own_lo, own_hi),
                ghost_rows=(ghost_lo, ghost_hi), ghost_owner=rank + 1 if rank < world - 1 else -1,
                ghost_source=rank - 1 if rank > 0 else -1)


class GhostExchange:
    """Static exchange plan for ghost rows that form one contiguous local row range sent to a single owner
    (z-slabs).  Works on CPU tensors (gloo) and CUDA tensors (nccl)."""

    def __init__(self, rank, world, wl):
        self.rank, self.world = rank, world
        self.off = wl["eqn_offset"]
        self.g_lo, self.g_hi = wl["ghost_rows"]
        self.dst, self.src = wl["ghost_owner"], wl["ghost_source"]
        self.pos = None
        self.recv_rows = None

    def setup(self, rowptr, col):
        """rowptr, col: local CSR pattern (torch tensors).  Exchanges the global keys of the ghost entries and
        locates them in the owner's CSR."""
        dev = rowptr.device
        self.seg = (0, 0)
        send_keys = torch.zeros(0, dtype=torch.int64, device=dev)
        if self.dst >= 0:
            a, b = int(rowptr[self.g_lo]), int(rowptr[self.g_hi])
            self.seg = (a, b)
            counts = rowptr[self.g_lo + 1:self.g_hi + 1] - rowptr[self.g_lo:self.g_hi]
            rows = torch.repeat_interleave(torch.arange(self.g_lo, self.g_hi, device=dev, dtype=torch.int64), counts)
            send_keys = ((rows + self.off) << 32) | (col[a:b].to(torch.int64) + self.off)
        # sizes first, then keys
        n_send = torch.tensor([send_keys.numel(), self.g_hi - self.g_lo if self.dst >= 0 else 0], dtype=torch.int64, device=dev)
        n_recv = torch.zeros(2, dtype=torch.int64, device=dev)
        self._sendrecv(n_send, n_recv)
        self.n_recv, self.n_recv_rows = int(n_recv[0]), int(n_recv[1])
        recv_keys = torch.zeros(self.n_recv, dtype=torch.int64, device=dev)
        self._sendrecv(send_keys, recv_keys)
        self.send_rows0 = torch.tensor([self.g_lo + self.off], dtype=torch.int64, device=dev)
        recv_row0 = torch.zeros(1, dtype=torch.int64, device=dev)
        self._sendrecv(self.send_rows0, recv_row0)
        if self.src >= 0:
            r = (recv_keys >> 32) - self.off_of_self()
            c = (recv_keys & 0xffffffff) - self.off_of_self()
            start, end = rowptr[r], rowptr[r + 1]
            pos = torch.full_like(r, -1)
            width = int((end - start).max()) if r.numel() else 0
            for k in range(width):
                idx = start + k
                ok = (idx < end) & (pos < 0)
                hit = ok & (col[torch.where(ok, idx, start)].to(torch.int64) == c)
                pos = torch.where(hit, idx, pos)
            if bool((pos < 0).any()):
                raise RuntimeError("ghost entry missing in the owner's pattern (halo elements not registered?)")
            self.pos = pos.contiguous()
            self.recv_row_lo = int(recv_row0[0]) - self.off
        return self

    def off_of_self(self):
        return self.off

    def _sendrecv(self, send, recv):
        ops = []
        if self.dst >= 0:
            ops.append(dist.P2POp(dist.isend, send, self.dst))
        if self.src >= 0:
            ops.append(dist.P2POp(dist.irecv, recv, self.src))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange(self, val, rhs, add_fn=None):
        """one assembly step: ghost values and ghost rhs rows go to the owner and are added there"""
        dev = val.device
        a, b = self.seg
        send_v = val[a:b] if self.dst >= 0 else val[:0]
        send_r = rhs[self.g_lo:self.g_hi] if self.dst >= 0 else rhs[:0]
        if not hasattr(self, "_rv") or self._rv.device != dev:
            self._rv = torch.empty(self.n_recv if self.src >= 0 else 0, dtype=val.dtype, device=dev)
            self._rr = torch.empty(self.n_recv_rows if self.src >= 0 else 0, dtype=val.dtype, device=dev)
        ops = []
        if self.dst >= 0:
            ops += [dist.P2POp(dist.isend, send_v, self.dst), dist.P2POp(dist.isend, send_r, self.dst)]
        if self.src >= 0:
            ops += [dist.P2POp(dist.irecv, self._rv, self.src), dist.P2POp(dist.irecv, self._rr, self.src)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.src >= 0:
            if add_fn is not None:
                add_fn(self.pos, self._rv, self.recv_row_lo, self._rr)
            else:
                val.index_add_(0, self.pos, self._rv)
                rhs[self.recv_row_lo:self.recv_row_lo + self.n_recv_rows] += self._rr


class _DevArray:
    """zero-copy view of engine device memory for torch (CUDA array interface)"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def wrap_device(ptr, n, dtype, device):
    typestr = {torch.float64: "<f8", torch.int64: "<i8", torch.int32: "<i4"}[dtype]
    if n == 0:
        return torch.zeros(0, dtype=dtype, device=device)
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class DistributedAssembly:
    """binds a GhostExchange to an Engine: device CSR wrapped as torch tensors, all work on the engine's stream"""

    def __init__(self, eng, wl, rank, world):
        self.eng, self.wl, self.rank, self.world = eng, wl, rank, world
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(eng.stream, device=self.device)
        self.plan = GhostExchange(rank, world, wl)

    def setup_fields(self):
        wl = self.wl
        self.eng.set_owned_elements(wl["n_owned_elems"])
        self.eng.set_field(0, 1, 1, wl["n_obj"], wl["elem_dof"], wl["eqn"], wl["status"], wl["presc"], wl["values"])

    def _tensors(self):
        n, nnz = self.eng.finish_assembly()
        rp, col, val, rhs = self.eng.device_csr()
        return (wrap_device(rp, n + 1, torch.int64, self.device), wrap_device(col, nnz, torch.int32, self.device),
                wrap_device(val, nnz, torch.float64, self.device), wrap_device(rhs, n, torch.float64, self.device))

    def setup_exchange(self):
        with torch.cuda.stream(self.stream):
            rp, col, self.val, self.rhs = self._tensors()
            self.plan.setup(rp, col)
            self.stream.synchronize()

    def _add(self, pos, rv, row_lo, rr):
        # the engine's own scatter-add kernels (RED.ADD.F64) on the engine stream
        self.eng.unpack_add_entries(0, pos.data_ptr(), pos.numel(), rv.data_ptr())
        if not hasattr(self, "_rows") or self._rows.numel() != rr.numel():
            self._rows = torch.arange(row_lo, row_lo + rr.numel(), dtype=torch.int64, device=self.device)
        self.eng.unpack_add_entries(1, self._rows.data_ptr(), rr.numel(), rr.data_ptr())

    def exchange(self):
        with torch.cuda.stream(self.stream):
            self.plan.exchange(self.val, self.rhs, add_fn=self._add)
