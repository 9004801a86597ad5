"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed).

Two ways the path shards:

* by configuration (training sets) — handled inside `kliff_b200.legacy.loss` (each rank takes a
  share of every batch; 641-number gradient all-reduce);
* **one large supercell by spatial domain with ghost-atom halos** — this module.  The reference
  has no counterpart (its `transform` is serial per configuration and dense, sym_fn.py:98-212); the
  closest piece of its semantics is `assemble_forces` (kliff/neighbor/neighbor.py:260-296), which
  folds partial forces on image atoms back onto their owners inside one process.  Here the same
  fold crosses ranks: every rank owns the atoms of one slab of the cell (slabs along lattice
  vector a1), imports the atoms within the cutoff of its two faces from its neighbours (periodic
  wrap included) as ghosts, computes descriptors and dG/dr for the atoms it owns with the usual
  kernels (periodic images along a2/a3 come from the ordinary padding kernel with PBC off along
  a1), and sends the partial forces that land on ghosts back to their owners.

The exchange itself is plain torch.distributed point-to-point (NCCL on the GPU box, gloo in the
CPU tests): two messages per neighbour and direction, latency-bound at these sizes.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops


def _face_distance(cell, axis):
    """Distance between the two cell faces perpendicular to reciprocal direction `axis`
    (volume / |a_j x a_k|, as neighbor_list.cpp:337-353)."""
    a = np.asarray(cell, dtype=np.float64)
    j, k = (axis + 1) % 3, (axis + 2) % 3
    x = np.cross(a[j], a[k])
    return abs(float(np.dot(a[axis], x))) / float(np.linalg.norm(x))


class DistTransport:
    """torch.distributed point-to-point (NCCL / gloo).  `post` only remembers the two outgoing
    messages; `collect` exchanges sizes (unless known) and payloads in one batch each."""

    def __init__(self, group=None):
        self.group = group
        self._out = {}

    def post(self, dd, kind, to_left, to_right):
        self._out[kind] = (to_left, to_right)

    def collect(self, dd, kind, width, expect=None):
        to_left, to_right = self._out.pop(kind)
        _, _, nb_left, nb_right = dd._sent
        dev, dt = to_left.device, to_left.dtype
        empty = torch.zeros((0, width), dtype=dt, device=dev)
        if dd.world == 1:  # both neighbours are this rank itself (periodic) or absent
            return (to_right if nb_left is not None else empty,
                    to_left if nb_right is not None else empty)
        g = self.group
        if expect is None:
            n_out = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=dev)
            n_in = torch.zeros(2, dtype=torch.int64, device=dev)
            reqs = []
            if nb_left is not None:
                reqs.append(dist.P2POp(dist.isend, n_out[0:1], nb_left[0], g, tag=1))
            if nb_right is not None:
                reqs.append(dist.P2POp(dist.isend, n_out[1:2], nb_right[0], g, tag=2))
            # receive from the right neighbour first: with two ranks both messages come from the
            # same peer and arrive in its send order (left-face message first)
            if nb_right is not None:
                reqs.append(dist.P2POp(dist.irecv, n_in[1:2], nb_right[0], g, tag=1))
            if nb_left is not None:
                reqs.append(dist.P2POp(dist.irecv, n_in[0:1], nb_left[0], g, tag=2))
            for w in dist.batch_isend_irecv(reqs):
                w.wait()
            n_left, n_right = int(n_in[0]), int(n_in[1])
        else:
            n_left, n_right = expect
        in_left = torch.zeros((n_left, width), dtype=dt, device=dev)
        in_right = torch.zeros((n_right, width), dtype=dt, device=dev)
        reqs = []
        if nb_left is not None and to_left.shape[0]:
            reqs.append(dist.P2POp(dist.isend, to_left, nb_left[0], g, tag=3))
        if nb_right is not None and to_right.shape[0]:
            reqs.append(dist.P2POp(dist.isend, to_right, nb_right[0], g, tag=4))
        if nb_right is not None and n_right:
            reqs.append(dist.P2POp(dist.irecv, in_right, nb_right[0], g, tag=3))
        if nb_left is not None and n_left:
            reqs.append(dist.P2POp(dist.irecv, in_left, nb_left[0], g, tag=4))
        if reqs:
            for w in dist.batch_isend_irecv(reqs):
                w.wait()
        return in_left, in_right


class InProcessTransport:
    """Mailbox shared by several SlabDecomposition objects living in ONE process: every emulated
    rank posts, then every emulated rank collects.  Used to exercise the multi-rank path on a
    single GPU (and in CPU tests) without torch.distributed."""

    def __init__(self):
        self.box = {}

    def post(self, dd, kind, to_left, to_right):
        _, _, nb_left, nb_right = dd._sent
        if nb_left is not None:
            self.box[(kind, dd.rank, nb_left[0], "L")] = to_left
        if nb_right is not None:
            self.box[(kind, dd.rank, nb_right[0], "R")] = to_right

    def collect(self, dd, kind, width, expect=None):
        _, _, nb_left, nb_right = dd._sent
        ref = None
        for v in self.box.values():
            ref = v
            break
        def take(nb, side):
            if nb is None:
                return torch.zeros((0, width), dtype=ref.dtype if ref is not None else torch.float64,
                                   device=ref.device if ref is not None else "cpu")
            return self.box.pop((kind, nb[0], dd.rank, side))
        # the left neighbour's message to ITS right neighbour (me), and vice versa
        return take(nb_left, "R"), take(nb_right, "L")


class SlabDecomposition:
    """Slabs of equal fractional thickness along lattice vector a1.

    rank/world may be given explicitly (tests emulate several ranks in one process); otherwise they
    come from torch.distributed.
    """

    def __init__(self, cell, pbc, cutoff, rank=None, world=None, group=None, transport=None):
        self.cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
        self.pbc = [bool(p) for p in pbc]
        self.cutoff = float(cutoff)
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world = int(rank), int(world)
        self.inv = np.linalg.inv(self.cell)
        self.halo = self.cutoff / _face_distance(self.cell, 0)  # halo width, fractional along a1
        if self.halo * (1 + 1e-12) > 1.0 / self.world:
            raise ValueError(
                f"slab thickness {1.0 / self.world:.4f} (fractional) is thinner than the halo "
                f"{self.halo:.4f}: use fewer ranks or a larger cell")
        self._sent = None
        self.transport = transport if transport is not None else DistTransport(group)

    # ------------------------------------------------------------------ ownership
    def fractional0(self, coords):
        f = coords @ torch.as_tensor(self.inv[:, 0], dtype=coords.dtype, device=coords.device)
        return f

    def wrap(self, coords):
        """Fold atoms into [0, 1) along a1 (only when that direction is periodic)."""
        if not self.pbc[0]:
            return coords
        f = self.fractional0(coords)
        shift = torch.floor(f)
        a1 = torch.as_tensor(self.cell[0], dtype=coords.dtype, device=coords.device)
        return coords - shift[:, None] * a1[None, :]

    def owner(self, coords):
        """Owning rank of every atom (coords already wrapped)."""
        f = self.fractional0(coords)
        o = torch.floor(f * self.world).to(torch.int64)
        return o.clamp_(0, self.world - 1)

    def _faces(self, coords_owned):
        f = self.fractional0(coords_owned)
        lo, hi = self.rank / self.world, (self.rank + 1) / self.world
        h = self.halo * (1.0 + 1e-12) + 1e-12
        left = torch.nonzero(f - lo <= h, as_tuple=False).reshape(-1)
        right = torch.nonzero(hi - f <= h, as_tuple=False).reshape(-1)
        return left, right

    def _neighbours(self):
        """(rank, image shift applied by the SENDER) for the left and right neighbour, or None."""
        W, r = self.world, self.rank
        left = right = None
        if r > 0:
            left = (r - 1, 0.0)
        elif self.pbc[0]:
            left = (W - 1, +1.0)   # my left-face atoms become right ghosts of the last slab: f + 1
        if r < W - 1:
            right = (r + 1, 0.0)
        elif self.pbc[0]:
            right = (0, -1.0)      # my right-face atoms become left ghosts of slab 0: f - 1
        return left, right

    # ------------------------------------------------------------------ halo exchange
    def _payload(self, coords_owned, code_owned, idx, nb):
        dev = coords_owned.device
        if nb is None:
            return torch.zeros((0, 4), dtype=torch.float64, device=dev)
        a1 = torch.as_tensor(self.cell[0], dtype=coords_owned.dtype, device=dev)
        x = coords_owned.index_select(0, idx) + nb[1] * a1[None, :]
        return torch.cat([x, code_owned.index_select(0, idx).to(torch.float64)[:, None]], dim=1).contiguous()

    def post_halo(self, coords_owned, code_owned):
        """First half of the exchange: select face atoms and hand them to the transport."""
        left_idx, right_idx = self._faces(coords_owned)
        nb_left, nb_right = self._neighbours()
        self._sent = (left_idx, right_idx, nb_left, nb_right)
        self.transport.post(self, "halo", self._payload(coords_owned, code_owned, left_idx, nb_left),
                            self._payload(coords_owned, code_owned, right_idx, nb_right))

    def collect_halo(self):
        """Second half: ghosts from the left neighbour first, then from the right.
        Returns (ghost_coords [G,3], ghost_code [G])."""
        in_left, in_right = self.transport.collect(self, "halo", width=4)
        self._ghost_counts = (in_left.shape[0], in_right.shape[0])
        ghosts = torch.cat([in_left, in_right], dim=0)
        return ghosts[:, :3].contiguous(), ghosts[:, 3].to(torch.int32).contiguous()

    def exchange_halo(self, coords_owned, code_owned):
        self.post_halo(coords_owned, code_owned)
        return self.collect_halo()

    def post_ghost_forces(self, forces_local, n_owned):
        gl, gr = self._ghost_counts
        self._f_own = forces_local[:n_owned].clone()
        self.transport.post(self, "force", forces_local[n_owned:n_owned + gl].contiguous(),
                            forces_local[n_owned + gl:n_owned + gl + gr].contiguous())

    def collect_ghost_forces(self):
        left_idx, right_idx, _, _ = self._sent
        got_left, got_right = self.transport.collect(self, "force", width=3,
                                                     expect=(left_idx.shape[0], right_idx.shape[0]))
        f_own = self._f_own
        if got_left.shape[0]:
            f_own.index_add_(0, left_idx, got_left)
        if got_right.shape[0]:
            f_own.index_add_(0, right_idx, got_right)
        return f_own

    def reduce_ghost_forces(self, forces_local, n_owned):
        """forces_local [n_owned + G, 3]: partial forces on owned atoms and ghosts.  Ghost rows go
        back to the ranks that own those atoms and are added there (the cross-rank
        `assemble_forces`).  Returns total forces on the owned atoms [n_owned, 3]."""
        self.post_ghost_forces(forces_local, n_owned)
        return self.collect_ghost_forces()

    # ------------------------------------------------------------------ local compute
    def local_fingerprints(self, descriptor, coords_owned, code_owned, ghost_coords, ghost_code, grad=True):
        """Descriptors (+ dG/dr) of the owned atoms of this slab.  Returns a dict with zeta
        [n_owned, D] and the force-scatter pack over the local atoms (owned, then ghosts)."""
        dev = coords_owned.device
        n_own = coords_owned.shape[0]
        local = torch.cat([coords_owned, ghost_coords], dim=0)
        code_local = torch.cat([code_owned.to(torch.int32), ghost_code.to(torch.int32)])
        nl = local.shape[0]
        pbc = [False, self.pbc[1], self.pbc[2]]
        pc, _, pi = ops.create_paddings(self.cutoff, self.cell, pbc, local)
        allc = torch.cat([local, pc]) if pc.shape[0] else local
        image = torch.cat([torch.arange(nl, dtype=torch.int32, device=dev), pi])
        code = code_local[image.long()]
        counts, offsets, indices, maxn = ops.build_neighbors(allc, n_own, self.cutoff)
        zeta, dz, dz_ptr = ops.symfun_forward(descriptor._params, allc, code, offsets, indices, maxn,
                                              n_centers=n_own, grad=grad, mode=descriptor.kernel_mode)
        pack = None
        if grad:
            pack = dict(offsets=offsets, indices=indices, image=image, dz=dz, dz_ptr=dz_ptr,
                        D=descriptor._params.n_desc, n_out=nl)
        return dict(zeta=zeta, pack=pack, n_owned=n_own, n_local=nl, n_pad=int(pc.shape[0]), coords=allc)

    def forces(self, dEdG_owned, fp):
        """Total forces on the owned atoms [n_owned, 3] from dE/dG of the owned atoms."""
        f_local = ops.force_scatter(dEdG_owned, fp["pack"]).reshape(fp["n_local"], 3)
        return self.reduce_ghost_forces(f_local, fp["n_owned"])


# ---------------------------------------------------------------------------------------------
# Sharding by configuration: the multi-GPU form of kliff/parallel.py
# ---------------------------------------------------------------------------------------------
def _dist_world():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def parmap1(f, X, *args, tuple_X=False, nprocs=None):
    """`kliff.parallel.parmap1` (kliff/parallel.py:8-101) with one process per GPU instead of forked
    host workers: same call, same return value (the list of `f(x, *args)` in the order of `X`, or
    `f(*x, *args)` with `tuple_X=True`).

    The reference forks `nprocs` workers and feeds them through queues; a CUDA context does not
    survive a fork, and the unit of parallelism here is the rank.  Under torch.distributed with more
    than one rank, rank r evaluates items r, r + world, ... and the results are exchanged with
    `all_gather_object`, so every rank returns the complete list (a collective call: every rank
    must make it with the same `X`).  In a single process the items are evaluated in order on the
    current device.  `nprocs` is accepted for signature compatibility and not used."""
    X = list(X)

    def call(x):
        return f(*x, *args) if tuple_X else f(x, *args)

    r, w = _dist_world()
    if w == 1:
        return [call(x) for x in X]
    mine = [(i, call(X[i])) for i in range(r, len(X), w)]
    parts = [None] * w
    dist.all_gather_object(parts, mine)
    out = [None] * len(X)
    for part in parts:
        for i, y in part:
            out[i] = y
    return out


# kliff/parallel.py:104-190 differs from parmap1 only in its transport (Pipe instead of Queue)
parmap2 = parmap1


def config_costs(configs, rc):
    """Estimated descriptor work of every configuration, atoms · n² (the triple loop of
    `generate_one_atom`, sym_fn.cpp:416-602, visits n(n-1)/2 pairs per atom).  n is estimated from the
    number density without building a neighbor list: n ≈ ρ · 4/3 π rc³, capped by what a non-periodic
    cluster can hold.  One vectorised pass (≈ 0.5 µs per configuration)."""
    m = len(configs)
    if m == 0:
        return np.zeros(0)
    natoms = np.fromiter((c.get_num_atoms() for c in configs), dtype=np.float64, count=m)
    cells = np.concatenate([np.asarray(c.cell, dtype=np.float64).reshape(9) for c in configs]).reshape(m, 3, 3)
    periodic = np.fromiter((all(c.PBC) for c in configs), dtype=bool, count=m)
    vol = np.abs(np.linalg.det(cells))
    with np.errstate(divide="ignore", invalid="ignore"):
        dense = np.where(vol > 1e-12, natoms / vol * (4.0 / 3.0) * np.pi * float(rc) ** 3, np.inf)
    n = np.where(periodic & np.isfinite(dense), dense, np.minimum(natoms - 1.0, dense))
    return natoms * np.maximum(n, 1.0) ** 2


def config_cost(conf, rc):
    """`config_costs` of one configuration."""
    return float(config_costs([conf], rc)[0])


def shard_members(costs, world, mode="auto", spread=2.0):
    """Global indices of the configurations every rank keeps: a list of `world` ascending int64 arrays.

    `round_robin`: rank r keeps r, r + world, ...  `balanced`: every run of `world` consecutive
    configurations still gives each rank exactly ONE member (so any global batch is split into shares
    that differ by at most one configuration, whatever the batch size), but inside a run the costliest
    configuration goes to the rank with the smallest load so far (longest-processing-time-first over
    the runs).  `auto` balances only when the costs differ by more than `spread` (max / min); a uniform
    set keeps the strided form.  Deterministic in (costs, world): every rank computes the same answer
    without communication."""
    costs = np.asarray(costs, dtype=np.float64)
    n = costs.shape[0]
    if mode not in ("auto", "balanced", "round_robin"):
        raise ValueError(f"shard mode {mode!r}: expected 'auto', 'balanced' or 'round_robin'")
    if mode == "auto":
        lo = float(costs.min()) if n else 0.0
        mode = "balanced" if n and (lo <= 0.0 or float(costs.max()) / lo > spread) else "round_robin"
    if mode == "round_robin" or world == 1:
        return [np.arange(r, n, world, dtype=np.int64) for r in range(world)]
    load = np.zeros(world)
    members = [[] for _ in range(world)]
    for g in range(0, n, world):
        idx = np.arange(g, min(g + world, n))
        heavy_first = idx[np.lexsort((idx, -costs[idx]))]
        light_first = np.lexsort((np.arange(world), load))
        for i, r in zip(heavy_first, light_first):
            members[r].append(int(i))
            load[r] += costs[i]
    return [np.asarray(sorted(m), dtype=np.int64) for m in members]
