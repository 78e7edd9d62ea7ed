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


def _run_p2p(ops):
    """ops: [("send" | "recv", tensor, peer), ...].  nccl moves CUDA tensors directly; gloo has no point-to-point
    operations on CUDA tensors, so they are staged through the host (used by the two-process test on ONE GPU, where
    NCCL refuses two ranks on the same device)."""
    if not ops:
        return
    stage = dist.get_backend() == "gloo"
    real, back = [], []
    for kind, t, peer in ops:
        if stage and t.is_cuda:
            if kind == "send":
                c = t.detach().cpu().contiguous()
            else:
                c = torch.empty(t.shape, dtype=t.dtype)
                back.append((t, c))
            real.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, c, peer))
        else:
            real.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer))
    for w in dist.batch_isend_irecv(real):
        w.wait()
    for t, c in back:
        t.copy_(c)


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
            ops.append(("send", send, self.dst))
        if self.src >= 0:
            ops.append(("recv", recv, self.src))
        _run_p2p(ops)

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
            ops += [("send", send_v, self.dst), ("send", send_r, self.dst)]
        if self.src >= 0:
            ops += [("recv", self._rv, self.src), ("recv", self._rr, self.src)]
        _run_p2p(ops)
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


def engine_comm_init(eng, rank, world):
    """NCCL communicator inside the engine (isl_comm_init): rank 0 makes the id, torch.distributed only carries the
    128 bytes to the other ranks.  ISL_TORCH_EXCHANGE=1 keeps the round-1 exchange through torch point-to-point calls."""
    import os
    if os.environ.get("ISL_TORCH_EXCHANGE") or not dist.is_initialized() or dist.get_backend() != "nccl":
        return False
    box = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    # NCCL may print its version banner on stdout when the communicator is created; stdout carries the bench's JSON line
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        eng.comm_init(box[0], rank, world)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return True


class DistributedAssembly:
    """binds a GhostExchange to an Engine: device CSR wrapped as torch tensors, all work on the engine's stream"""

    def __init__(self, eng, wl, rank, world):
        self.eng, self.wl, self.rank, self.world = eng, wl, rank, world
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(eng.stream, device=self.device)
        self.plan = GhostExchange(rank, world, wl)
        self.native = engine_comm_init(eng, rank, world)

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
        if self.native:
            wl = self.wl
            self.eng.finish_assembly()
            l2g = np.arange(wl["n_eqn_local"], dtype=np.int64) + wl["eqn_offset"]
            segs = [(wl["ghost_owner"],) + tuple(wl["ghost_rows"])] if wl["ghost_owner"] >= 0 else []
            self.eng.exchange_setup(l2g, wl["owned_rows"][0], wl["owned_rows"][1], segs)
            return
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
        if self.native:
            self.eng.exchange()    # isl_exchange: NCCL on the engine's communication stream, overlapped with interior patches
            return
        # a Q1 stiffness launch may still be deferred inside the engine (it waits one call for a body force to fuse):
        # the ghost rows must be in memory before they are sent
        self.eng.flush()
        with torch.cuda.stream(self.stream):
            self.plan.exchange(self.val, self.rhs, add_fn=self._add)


# =====================================================================================================================
# General element-block partition (unstructured case): any mesh, any number of fields.
# Every rank computes the same partition from the global description; one process per GPU assembles its element block
# into a local CSR over (owned rows | ghost rows | halo-only ids) and ships the ghost rows to their owners.
# =====================================================================================================================
def morton_order(coords, conn, bits=10):
    """element order along a Z-curve of the centroids: contiguous blocks of it are compact element blocks"""
    cen = coords[conn].mean(axis=1)
    lo, hi = cen.min(axis=0), cen.max(axis=0)
    q = np.minimum(((cen - lo) / np.maximum(hi - lo, 1e-300) * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    key = np.zeros(len(conn), dtype=np.int64)
    dim = coords.shape[1]
    for b in range(bits):
        for d in range(dim):
            key |= ((q[:, d] >> b) & 1) << (b * dim + d)
    return np.argsort(key, kind="stable")


def general_partition(coords, conn, fields, n_eqn, rank, world, order=None):
    """Local workload of `rank`.

    coords [n_nodes, dim], conn [n_elems, npe]; fields: list of dicts with elem_dof [n_elems, ndpe], eqn [n_obj, ds]
    (global equation numbers, < 0 where not ACTIVE), status, presc, values [n_obj, ds]; n_eqn: global system size.
    order: element permutation whose contiguous blocks become the element blocks (default: Morton order).

    Returns a dict: local mesh (owned elements first, then halo elements that only shape the pattern of owned rows),
    local fields with LOCAL equation numbers, l2g (local -> global equation), n_owned_rows, ghost segments per owner."""
    ne = len(conn)
    order = morton_order(coords, conn) if order is None else np.asarray(order)
    elem_owner = np.empty(ne, dtype=np.int64)
    elem_owner[order] = np.arange(ne, dtype=np.int64) * world // ne
    # row owner = lowest rank among the elements touching the row
    row_owner = np.full(n_eqn, world, dtype=np.int64)
    elem_rows = []
    for f in fields:
        er = f["eqn"][f["elem_dof"]].reshape(ne, -1)            # [ne, ndpe*ds] global equation numbers (or < 0)
        elem_rows.append(er)
        ok = er >= 0
        np.minimum.at(row_owner, er[ok], np.broadcast_to(elem_owner[:, None], er.shape)[ok])
    assert n_eqn == 0 or row_owner.max() < world, "an equation is touched by no element"
    rows_all = np.concatenate(elem_rows, axis=1)
    owned_e = np.nonzero(elem_owner == rank)[0]
    touches_owned = ((rows_all >= 0) & (row_owner[np.maximum(rows_all, 0)] == rank)).any(axis=1)
    halo_e = np.nonzero(touches_owned & (elem_owner != rank))[0]
    local_e = np.concatenate([owned_e, halo_e])
    # local nodes
    nodes = np.unique(conn[local_e])
    g2l_node = np.full(len(coords), -1, dtype=np.int64); g2l_node[nodes] = np.arange(len(nodes))
    out = dict(coords=np.ascontiguousarray(coords[nodes]), conn=np.ascontiguousarray(g2l_node[conn[local_e]].astype(np.int32)),
               n_owned_elems=len(owned_e), elements=local_e, fields=[])
    # local equation numbering: owned rows (ascending), ghost rows grouped by owner (ascending), halo-only ids
    touched_by_owned = np.unique(rows_all[owned_e][rows_all[owned_e] >= 0])
    touched_local = np.unique(rows_all[local_e][rows_all[local_e] >= 0])
    owned_rows = np.nonzero(row_owner == rank)[0]
    assert np.isin(owned_rows, touched_local).all()
    ghost = touched_by_owned[row_owner[touched_by_owned] != rank]
    ghost = ghost[np.lexsort((ghost, row_owner[ghost]))]
    rest = np.setdiff1d(touched_local, np.concatenate([owned_rows, ghost]))
    l2g = np.concatenate([owned_rows, ghost, rest]).astype(np.int64)
    g2l = np.full(n_eqn, -1, dtype=np.int64); g2l[l2g] = np.arange(len(l2g))
    for f, er in zip(fields, elem_rows):
        objs = np.unique(f["elem_dof"][local_e])
        g2l_obj = np.full(len(f["eqn"]), -1, dtype=np.int64); g2l_obj[objs] = np.arange(len(objs))
        eq = f["eqn"][objs]
        leq = np.where(eq >= 0, g2l[np.maximum(eq, 0)], -1)
        out["fields"].append(dict(fe_deg=f["fe_deg"], ds=f["ds"], n_obj=len(objs),
                                  elem_dof=np.ascontiguousarray(g2l_obj[f["elem_dof"][local_e]].astype(np.int32)),
                                  eqn=np.ascontiguousarray(leq.astype(np.int64)), status=np.ascontiguousarray(f["status"][objs]),
                                  presc=np.ascontiguousarray(f["presc"][objs]), values=np.ascontiguousarray(f["values"][objs])))
    owners = row_owner[ghost]
    out.update(l2g=l2g, n_eqn_local=len(l2g), n_owned_rows=len(owned_rows), n_ghost_rows=len(ghost),
               ghost_segments=[(int(r), int(np.searchsorted(owners, r, "left")) + len(owned_rows),
                                int(np.searchsorted(owners, r, "right")) + len(owned_rows)) for r in np.unique(owners)],
               n_eqn_global=int(n_eqn))
    return out


class GeneralExchange:
    """Ghost rows -> owners for general_partition(): every rank sends, per owner, one contiguous slice of its CSR values
    and of its rhs; the owner adds them at positions found once from the (global row, global column) keys.
    Point-to-point (batch_isend_irecv), so it runs on gloo (CPU tensors) and nccl (CUDA tensors)."""

    def __init__(self, rank, world, wl):
        self.rank, self.world, self.wl = rank, world, wl
        self.send = []   # (dst, val slice, row slice)
        self.recv = []   # (src, positions in val, local rows)

    def setup(self, rowptr, col):
        dev = rowptr.device
        l2g = torch.from_numpy(self.wl["l2g"]).to(dev)
        g2l = torch.full((self.wl["n_eqn_global"],), -1, dtype=torch.int64, device=dev)
        g2l[l2g] = torch.arange(len(l2g), device=dev)
        # who sends to whom, and how much: exchange (n_entries, n_rows) for every ordered pair via all_gather
        mine = torch.zeros(self.world, 2, dtype=torch.int64, device=dev)
        keys = {}
        for dst, lo, hi in self.wl["ghost_segments"]:
            a, b = int(rowptr[lo]), int(rowptr[hi])
            counts = rowptr[lo + 1:hi + 1] - rowptr[lo:hi]
            rows = torch.repeat_interleave(torch.arange(lo, hi, device=dev, dtype=torch.int64), counts)
            keys[dst] = torch.stack([l2g[rows], l2g[col[a:b].to(torch.int64)]])
            mine[dst, 0], mine[dst, 1] = b - a, hi - lo
            self.send.append((dst, (a, b), (lo, hi)))
        if dist.get_backend() == "gloo" and mine.is_cuda:   # no CUDA collectives in gloo: through the host
            mc = mine.cpu()
            tc = [torch.zeros_like(mc) for _ in range(self.world)]
            dist.all_gather(tc, mc)
            table = [t.to(dev) for t in tc]
        else:
            table = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(table, mine)
        ops, rbuf = [], {}
        for dst, _, (lo, hi) in self.send:
            ops.append(("send", keys[dst].contiguous(), dst))
            ops.append(("send", l2g[lo:hi].contiguous(), dst))
        for src in range(self.world):
            n_ent, n_rows = int(table[src][self.rank, 0]), int(table[src][self.rank, 1])
            if src == self.rank or n_rows == 0:
                continue
            rbuf[src] = (torch.zeros(2, n_ent, dtype=torch.int64, device=dev), torch.zeros(n_rows, dtype=torch.int64, device=dev))
            ops.append(("recv", rbuf[src][0], src))
            ops.append(("recv", rbuf[src][1], src))
        _run_p2p(ops)
        for src, (k, grows) in rbuf.items():
            r, c = g2l[k[0]], g2l[k[1]]
            if bool((r < 0).any()) or bool((r >= self.wl["n_owned_rows"]).any()):
                raise RuntimeError("received a row this rank does not own")
            start, end = rowptr[r], rowptr[r + 1]
            pos = torch.full_like(r, -1)
            width = int((end - start).max()) if r.numel() else 0
            for j in range(width):
                idx = start + j
                ok = (idx < end) & (pos < 0)
                hit = ok & (col[torch.where(ok, idx, start)].to(torch.int64) == c)
                pos = torch.where(hit, idx, pos)
            if bool((pos < 0).any()):
                raise RuntimeError("ghost entry missing in the owner's pattern (halo elements not registered?)")
            self.recv.append((src, pos.contiguous(), g2l[grows].contiguous()))
        return self

    def exchange(self, val, rhs, add_fn=None):
        ops, bufs = [], []
        for dst, (a, b), (lo, hi) in self.send:
            ops += [("send", val[a:b], dst), ("send", rhs[lo:hi], dst)]
        for src, pos, rows in self.recv:
            bv = torch.empty(pos.numel(), dtype=val.dtype, device=val.device)
            br = torch.empty(rows.numel(), dtype=val.dtype, device=val.device)
            bufs.append((pos, rows, bv, br))
            ops += [("recv", bv, src), ("recv", br, src)]
        _run_p2p(ops)
        for pos, rows, bv, br in bufs:
            if add_fn is not None:
                add_fn(pos, bv, rows, br)
            else:
                val.index_add_(0, pos, bv)
                rhs.index_add_(0, rows, br)


class GeneralDistributedAssembly:
    """general_partition() bound to an Engine (one process per GPU): local mesh / fields on the device, ghost-row
    exchange with the engine's scatter-add kernels on the engine stream"""

    def __init__(self, eng, wl, rank, world, shape, geom_deg):
        self.eng, self.wl = eng, wl
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(eng.stream, device=self.device)
        self.plan = GeneralExchange(rank, world, wl)
        self.native = engine_comm_init(eng, rank, world)
        eng.set_mesh(shape, geom_deg, wl["coords"], wl["conn"])
        eng.set_owned_elements(wl["n_owned_elems"])
        for i, f in enumerate(wl["fields"]):
            eng.set_field(i, f["fe_deg"], f["ds"], f["n_obj"], f["elem_dof"], f["eqn"], f["status"], f["presc"], f["values"])

    def setup_exchange(self):
        if self.native:
            wl = self.wl
            self.eng.finish_assembly()
            self.eng.exchange_setup(wl["l2g"], 0, wl["n_owned_rows"], wl["ghost_segments"])
            return
        with torch.cuda.stream(self.stream):
            n, nnz = self.eng.finish_assembly()
            rp, col, val, rhs = self.eng.device_csr()
            self.val = wrap_device(val, nnz, torch.float64, self.device)
            self.rhs = wrap_device(rhs, n, torch.float64, self.device)
            self.plan.setup(wrap_device(rp, n + 1, torch.int64, self.device), wrap_device(col, nnz, torch.int32, self.device))
            self.stream.synchronize()

    def _add(self, pos, bv, rows, br):
        self.eng.unpack_add_entries(0, pos.data_ptr(), pos.numel(), bv.data_ptr())
        self.eng.unpack_add_entries(1, rows.data_ptr(), rows.numel(), br.data_ptr())

    def exchange(self):
        if self.native:
            self.eng.exchange()
            return
        self.eng.flush()   # see DistributedAssembly.exchange
        with torch.cuda.stream(self.stream):
            self.plan.exchange(self.val, self.rhs, add_fn=self._add)


# ---- Taylor-Hood Stokes on a cube of tetrahedra, cut into z-slabs, generated per rank ---------------------------------
def _unit_hash(ids, salt):
    """deterministic value in [-1, 1) per integer id: every rank perturbs a shared vertex identically"""
    x = (ids.astype(np.uint64) + np.uint64(salt)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def structured_stokes_slab(n, rank, world, perturb=0.1, permute=True, seed=54321):
    """Local workload of `rank` for the driven-cavity blocks (reference/07-drivenCavity/drivenCavity.cpp:176-275) on
    6 n^3 tetrahedra (every cube of the n^3 lattice split as unitCube.hpp:107-134), P2 velocity x 3 + P1 pressure, WITHOUT
    ever building the global mesh: cube layers [z0, z1) are owned, the layer above is the halo, and the numbering is in
    closed form (velocity DoF objects = points of the (2n+1)^3 half-step lattice, pressure DoF objects = vertices; global
    equation = lexicographic rank among the ACTIVE DoFs, velocity block first, as numberDoFsConsecutively with the
    drivenCavity offsets gives for that object order).  Same dict as general_partition(); rows are owned by the lowest
    rank whose elements touch them, i.e. the slab below owns a shared plane."""
    from . import engine as E
    from . import meshgen
    V, F, m = n + 1, 2 * n + 1, 2 * n - 1
    bounds = [n * r // world for r in range(world + 1)]
    z0, z1 = bounds[rank], bounds[rank + 1]
    assert z1 > z0, "every rank needs at least one cube layer"
    zc1 = z1 + (1 if rank < world - 1 else 0)        # local cube layers [z0, zc1): owned + one halo layer above
    layer_rank = np.searchsorted(np.asarray(bounds[1:]), np.arange(n), side="right")

    # vertices of the local layers, lexicographic, and their (perturbed) positions
    kv = np.arange(z0, zc1 + 1)
    gi, gj, gk = np.meshgrid(np.arange(V), np.arange(V), kv, indexing="ij")
    order3 = lambda a: np.ascontiguousarray(a.transpose(2, 1, 0)).reshape(-1)     # x fastest
    vi, vj, vk = order3(gi), order3(gj), order3(gk)
    coords = np.stack([vi, vj, vk], axis=1).astype(np.float64) / n
    if perturb:
        gid = vi + V * (vj + V * vk.astype(np.int64))
        interior = (vi > 0) & (vi < n) & (vj > 0) & (vj < n) & (vk > 0) & (vk < n)
        d = np.stack([_unit_hash(gid, 11 + c) for c in range(3)], axis=1)
        nrm = np.linalg.norm(d, axis=1, keepdims=True)
        d = d / np.maximum(nrm, 1e-300) * (_unit_hash(gid, 17)[:, None] * (perturb / n))
        coords[interior] += d[interior]
    # tetrahedra: integer vertex coordinates [ne, 4, 3]
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(z0, zc1), indexing="ij")
    cube = np.stack([order3(ci), order3(cj), order3(ck)], axis=1)                  # [nc, 3], layer slowest
    corner = np.array([[c & 1, (c >> 1) & 1, c >> 2] for c in range(8)])
    tets = corner[meshgen._H_HEX[meshgen._TETS]]                                   # [6, 4, 3] offsets inside the cube
    vint = (cube[:, None, None, :] + tets[None]).reshape(-1, 4, 3)                 # [6 nc, 4, 3]
    n_owned = 6 * n * n * (z1 - z0)
    if permute:
        rng = np.random.default_rng(seed + rank)
        p = np.concatenate([rng.permutation(n_owned), np.arange(n_owned, len(vint))])
        vint = vint[p]
    conn = (vint[..., 0] + V * (vint[..., 1] + V * (vint[..., 2] - z0))).astype(np.int32)

    def owner_of_plane(k, half):        # lowest cube layer touching lattice plane k (half-step lattice if half)
        lay = np.maximum((k - 1) // 2 if half else k - 1, 0)
        return layer_rank[np.minimum(lay, n - 1)]

    fields, classes, geqs = [], [], []
    # velocity: support points of the P2 element as combinations of its vertices -> half-step lattice coordinates
    sp = E.support_points(E.TET, 2)
    w2 = np.rint(2 * np.array([E.shape_eval(E.TET, 1, s)[0] for s in sp])).astype(np.int64)     # [10, 4], rows sum to 2
    fint = np.einsum("la,ead->eld", w2, vint.astype(np.int64))                     # [ne, 10, 3]
    ed_u = (fint[..., 0] + F * (fint[..., 1] + F * (fint[..., 2] - 2 * z0))).astype(np.int32)
    fk_l = np.arange(2 * z0, 2 * zc1 + 1)
    fi, fj, fk = [order3(a) for a in np.meshgrid(np.arange(F), np.arange(F), fk_l, indexing="ij")]
    bnd = (fi == 0) | (fi == 2 * n) | (fj == 0) | (fj == 2 * n) | (fk == 0) | (fk == 2 * n)
    st = np.zeros((len(fi), 3), dtype=np.uint8); st[bnd] = E.CONSTRAINED
    presc = np.zeros((len(fi), 3)); presc[fk == 2 * n, 0] = 1.0
    g = 3 * ((fi - 1) + m * ((fj - 1) + m * (fk.astype(np.int64) - 1)))
    geq = np.where(bnd[:, None], -1, g[:, None] + np.arange(3)[None])
    n_u = 3 * m ** 3
    fields.append(dict(fe_deg=2, ds=3, n_obj=len(fi), elem_dof=np.ascontiguousarray(ed_u), status=st, presc=presc,
                       values=np.zeros((len(fi), 3))))
    geqs.append(geq)
    classes.append(np.broadcast_to(owner_of_plane(fk, True)[:, None], geq.shape))
    # pressure: one DoF per vertex, vertex 0 pinned
    gidp = vi + V * (vj + V * vk.astype(np.int64))
    stp = np.zeros((len(vi), 1), dtype=np.uint8); stp[gidp == 0] = E.CONSTRAINED
    geqp = np.where(gidp == 0, -1, n_u + gidp - 1)[:, None]
    fields.append(dict(fe_deg=1, ds=1, n_obj=len(vi), elem_dof=np.ascontiguousarray(conn.copy()), status=stp,
                       presc=np.zeros((len(vi), 1)), values=np.zeros((len(vi), 1))))
    geqs.append(geqp)
    classes.append(owner_of_plane(vk, False)[:, None])
    # local numbering: owned rows, ghost rows (owned by the slab below), rows only the halo layer touches
    allg = np.concatenate([q.reshape(-1) for q in geqs])
    own = np.concatenate([np.asarray(c).reshape(-1) for c in classes])
    cls = np.where(own == rank, 0, np.where(own < rank, 1, 2))
    act = np.nonzero(allg >= 0)[0]
    act = act[np.lexsort((allg[act], cls[act]))]
    leq = np.full(len(allg), -1, dtype=np.int64); leq[act] = np.arange(len(act))
    l2g = allg[act]
    n_own_rows, n_ghost = int((cls[act] == 0).sum()), int((cls[act] == 1).sum())
    a = 0
    for f, q in zip(fields, geqs):
        f["eqn"] = np.ascontiguousarray(leq[a:a + q.size].reshape(q.shape)); a += q.size
    return dict(coords=coords, conn=np.ascontiguousarray(conn), n_owned_elems=n_owned, fields=fields, l2g=l2g,
                n_eqn_local=len(l2g), n_owned_rows=n_own_rows, n_ghost_rows=n_ghost,
                ghost_segments=[(rank - 1, n_own_rows, n_own_rows + n_ghost)] if n_ghost else [],
                n_eqn_global=int(n_u + V ** 3 - 1), n_elems_global=6 * n ** 3)
