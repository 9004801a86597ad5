"""GPU-resident packed fingerprints: what the descriptor kernels produce for a list of
configurations and what the force-scatter kernel consumes.

The reference keeps, per configuration, a dense `dzetadr_forces` of shape (N, D, 3N)
(kliff/legacy/descriptors/symmetry_function/sym_fn.py:163-174) and re-uploads it every batch
(calculator_torch.py:211).  Here all configurations are concatenated into one atom array
(per configuration: contributing atoms, then its padding atoms) with one CSR neighbor list, and
the gradient stays in the kernels' native per-atom block layout [D][n+1][3]:

    coords   f64 [A, 3]      all atoms of all configurations
    code     i32 [A]         species code (cutoff-matrix index)
    image    i32 [A]         ordinal of the contributing atom an atom is an image of
    offsets  i64 [A + 1]     CSR over atoms (padding atoms have empty lists)
    indices  i32 [sum n]     neighbor atom ids, ascending per atom
    centers  i32 [Nc]        atom id of every contributing atom (configuration-major)
    zeta     f64 [Nc, D]     raw descriptors
    dz       f64 [...]       blocks, atom t at dz[dz_ptr[t]:], 128-byte aligned
    cfg_ptr  host int64 [C+1]  centre range of configuration c = [cfg_ptr[c], cfg_ptr[c+1])

Memory: 36 KB per atom in fp64 for set51 at n = 28.
"""
import numpy as np
import torch

from . import ops


class PackedGradient:
    """Host copy of one configuration's gradient blocks, picklable; stands in for the dense
    `dzetadr_forces` array inside a fingerprints pickle (`to_dense()` gives the reference's
    (N, D, 3N) array)."""

    def __init__(self, values, counts, neighbors, image, n_desc):
        self.values = values        # 1-D, blocks [D][n_t+1][3] back to back (already / stdev)
        self.counts = counts        # i32 [N]
        self.neighbors = neighbors  # i32 [sum counts], LOCAL atom ids (contributing + padding)
        self.image = image          # i32 [N + P], local owner of every atom
        self.n_desc = n_desc

    def to_dense(self):
        N, D = len(self.counts), self.n_desc
        out = np.zeros((N, D, N, 3), dtype=self.values.dtype)
        pos = nb = 0
        for t in range(N):
            n = int(self.counts[t])
            blk = self.values[pos:pos + 3 * D * (n + 1)].reshape(D, n + 1, 3)
            ids = np.concatenate((self.neighbors[nb:nb + n], [t]))
            for slot, a in enumerate(ids):
                out[t, :, self.image[a], :] += blk[:, slot, :]
            pos += 3 * D * (n + 1)
            nb += n
        return out.reshape(N, D, 3 * N)

    @property
    def shape(self):
        N = len(self.counts)
        return (N, self.n_desc, 3 * N)


def host_arrays(configs, species_code):
    """Flat host arrays of a list of configurations for the batched kernels: atom offsets, stacked
    coordinates, species codes, cells and PBC flags.  One pass over the Python objects with as few
    per-configuration numpy calls as possible (20 000 configurations: 0.12 s -> a third of that);
    species lists are translated once per distinct list.  Raises KeyError for an unknown species."""
    nc = len(configs)
    sizes = np.fromiter((c.get_num_atoms() for c in configs), np.int64, count=nc)
    cfg_ptr = np.zeros(nc + 1, np.int64)
    np.cumsum(sizes, out=cfg_ptr[1:])
    n = int(cfg_ptr[-1])
    if n == 0:
        return cfg_ptr, np.zeros((0, 3)), np.zeros(0, np.int32), np.zeros((nc, 3, 3)), np.zeros((nc, 3), bool)
    try:  # one concatenate when every configuration already holds an (n, 3) array
        coords = np.concatenate([c.coords for c in configs], dtype=np.float64)
        if coords.ndim != 2 or coords.shape != (n, 3):
            raise ValueError
    except (ValueError, TypeError):
        coords = np.concatenate([np.asarray(c.coords, np.float64).reshape(-1, 3) for c in configs])
    seen, codes = {}, []
    for c in configs:
        sp = c.species
        key = tuple(sp)
        arr = seen.get(key)
        if arr is None:
            arr = seen[key] = np.asarray([species_code[s] for s in key], np.int32)
        codes.append(arr)
    code_cb = np.concatenate(codes)
    try:
        cells = np.array([c.cell for c in configs], np.float64).reshape(nc, 3, 3)
        pbcs = np.array([c.PBC for c in configs]).astype(bool).reshape(nc, 3)
    except ValueError:  # ragged input (a cell given as a flat list, ...): per-configuration conversion
        cells = np.stack([np.asarray(c.cell, np.float64).reshape(3, 3) for c in configs])
        pbcs = np.stack([np.asarray(c.PBC).astype(bool).reshape(3) for c in configs])
    return cfg_ptr, coords, code_cb, cells, pbcs


class PackedFingerprints:
    def __init__(self, D, device):
        self.D = D
        self.device = device
        self.coords = self.code = self.image = self.offsets = self.indices = None
        self.centers = self.zeta = self.dz = self.dz_ptr = None
        self.cfg_ptr = np.zeros(1, np.int64)   # centre ranges
        self.atom_ptr = np.zeros(1, np.int64)  # atom ranges (contributing + padding)
        self.maxn = 0
        self.scale = None  # 1/stdev when normalised, folded into the contraction

    @property
    def n_configs(self):
        return len(self.cfg_ptr) - 1

    @property
    def n_centers(self):
        return int(self.cfg_ptr[-1])

    # ------------------------------------------------------------------ assembly
    @classmethod
    def from_parts(cls, D, device, parts):
        """parts: per configuration dict(coords, code, image_local, counts, indices_local,
        n_contrib) of device tensors; concatenates with index shifting."""
        self = cls(D, device)
        atom_base, cen_base = 0, 0
        coords, code, image, counts, indices, centers = [], [], [], [], [], []
        atom_ptr, cfg_ptr = [0], [0]
        maxn = 0
        for p in parts:
            na = p["coords"].shape[0]
            coords.append(p["coords"])
            code.append(p["code"])
            image.append(p["image_local"] + cen_base)
            counts.append(p["counts"])
            indices.append(p["indices_local"] + atom_base)
            centers.append(torch.arange(atom_base, atom_base + p["n_contrib"], dtype=torch.int32,
                                        device=device))
            atom_base += na
            cen_base += p["n_contrib"]
            atom_ptr.append(atom_base)
            cfg_ptr.append(cen_base)
            maxn = max(maxn, p["maxn"])
        self.coords = torch.cat(coords)
        self.code = torch.cat(code).to(torch.int32)
        self.image = torch.cat(image).to(torch.int32)
        cnt = torch.cat(counts)
        self.offsets = torch.zeros(cnt.shape[0] + 1, dtype=torch.int64, device=device)
        torch.cumsum(cnt, 0, out=self.offsets[1:])
        self.indices = torch.cat(indices).to(torch.int32) if indices else torch.zeros(0, dtype=torch.int32, device=device)
        self.centers = torch.cat(centers)
        self.atom_ptr = np.asarray(atom_ptr, np.int64)
        self.cfg_ptr = np.asarray(cfg_ptr, np.int64)
        self.maxn = maxn
        return self

    @classmethod
    def from_configs(cls, D, device, configs, species_code, infl_dist):
        """Batched assembly: paddings + neighbor lists of ALL configurations in one pass of the
        kb_batch_* kernels (one CTA per configuration) instead of one NeighborList per
        configuration (descriptor.py:186-238 loops `for conf in configs`)."""
        self = cls(D, device)
        cfg_ptr, coords, code_cb, cells, pbcs = host_arrays(configs, species_code)
        d_coords = torch.as_tensor(coords, device=device)
        out = ops.batch_neighbors(cfg_ptr, cells, pbcs, d_coords, infl_dist)
        self.coords, self.image, self.centers = out["coords"], out["image"], out["centers"]
        self.offsets, self.indices, self.maxn = out["offsets"], out["indices"], out["maxn"]
        self.code = torch.as_tensor(code_cb, device=device)[self.image.long()]
        self.atom_ptr = out["atom_ptr"].astype(np.int64)
        self.cfg_ptr = cfg_ptr
        return self

    def compute(self, params, grad=True, mode=0, dz_dtype=torch.float64, zeta_host=None,
                piece=1 << 18):
        """Run the descriptor kernels over every centre atom.

        `zeta_host` (pinned host f64 [n_centers, D]): stream the descriptors to the host while the
        kernels are still working — the centre list is processed in pieces and each piece's rows
        are copied device->host on a side stream as soon as that piece is done, so the PCIe time
        (8 D bytes per atom) hides behind the next piece's compute.  `self.host_ready` is the event
        to wait for before reading `zeta_host`."""
        self.host_ready = None
        n = self.n_centers
        if zeta_host is None or n <= piece:
            self.zeta, self.dz, self.dz_ptr = ops.symfun_forward(
                params, self.coords, self.code, self.offsets, self.indices, self.maxn,
                centers=self.centers, grad=grad, mode=mode, dz_dtype=dz_dtype)
            if zeta_host is not None:
                zeta_host.copy_(self.zeta, non_blocking=True)
                self.host_ready = torch.cuda.Event()
                self.host_ready.record()
            return self
        if tuple(zeta_host.shape) != (n, self.D) or zeta_host.dtype != torch.float64 or not zeta_host.is_pinned():
            raise ValueError("zeta_host must be a pinned float64 [n_centers, D] host tensor")
        dev = self.coords.device
        self.zeta = torch.empty((n, self.D), dtype=torch.float64, device=dev)
        if grad:
            self.dz_ptr, total = ops.block_offsets(n, self.centers, self.offsets, self.D)
            self.dz = torch.empty((total,), dtype=dz_dtype, device=dev)
        else:
            self.dz = self.dz_ptr = None
        main = torch.cuda.current_stream()
        side = getattr(PackedFingerprints, "_side_stream", None)
        if side is None or side.device != dev:
            side = PackedFingerprints._side_stream = torch.cuda.Stream(device=dev)
        for lo in range(0, n, piece):
            hi = min(n, lo + piece)
            ops.symfun_forward(params, self.coords, self.code, self.offsets, self.indices, self.maxn,
                               centers=self.centers[lo:hi], grad=grad, mode=mode,
                               dz_ptr=self.dz_ptr[lo:hi + 1] if grad else None, dz=self.dz,
                               zeta_out=self.zeta[lo:hi])
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)
            with torch.cuda.stream(side):
                zeta_host[lo:hi].copy_(self.zeta[lo:hi], non_blocking=True)
        self.zeta.record_stream(side)
        self.host_ready = torch.cuda.Event()
        self.host_ready.record(side)
        return self

    # ------------------------------------------------------------------ per-configuration views
    def config_slice(self, c):
        return int(self.cfg_ptr[c]), int(self.cfg_ptr[c + 1])

    def counts_of(self, c):
        lo, hi = self.config_slice(c)
        a0 = int(self.atom_ptr[c])
        return (self.offsets[a0 + 1:a0 + (hi - lo) + 1] - self.offsets[a0:a0 + (hi - lo)]).to(torch.int32)

    def pack_for(self, lo, hi):
        """Argument pack of the force-scatter kernels for the centre range [lo, hi) (a run of
        whole configurations).  Forces come out relative to contributing atom `lo`."""
        plan = self.scatter_plan()
        sp = plan["slot_ptr"]
        if lo == 0 and hi == self.n_centers:
            base, cnt = 0, plan["n_slots"]
        else:
            host = self.slot_ptr_host()
            base, cnt = int(host[lo]), int(host[hi] - host[lo])
        sub = dict(slot_ptr=sp[lo:hi + 1], seg_ptr=plan["seg_ptr"][lo:hi + 1], seg_src=plan["seg_src"],
                   slot_base=base, n_slots=cnt)
        return dict(offsets=self.offsets, indices=self.indices, image=self.image, dz=self.dz,
                    dz_ptr=self.dz_ptr[lo:hi + 1], centers=self.centers[lo:hi], D=self.D,
                    n_out=hi - lo, out_base=lo, scale=self.scale, plan=sub)

    def scatter_plan(self):
        """Owner-sorted plan of the deterministic force assembly over ALL centres (built once; integer
        work): slot ids per owner atom in ascending order (ops.scatter_plan)."""
        if getattr(self, "_plan", None) is None:
            self._plan = ops.scatter_plan(self.offsets, self.indices, self.image, self.n_centers, self.centers)
            self._slot_ptr_host = None
        return self._plan

    def slot_ptr_host(self):
        """Host copy of the plan's slot offsets (needed to slice the plan for a sub-range of configurations)."""
        self.scatter_plan()
        if self._slot_ptr_host is None:
            self._slot_ptr_host = self._plan["slot_ptr"].cpu().numpy()
        return self._slot_ptr_host

    def dense_config(self, c, want_forces=True, want_stress=False):
        """The reference's dense per-configuration views (sym_fn.py:163-185)."""
        lo, hi = self.config_slice(c)
        a0 = int(self.atom_ptr[c])
        pack = dict(offsets=self.offsets, indices=self.indices, dz=self.dz, dz_ptr=self.dz_ptr[lo:hi + 1],
                    centers=self.centers[lo:hi], D=self.D,
                    image=self.image - lo)  # owners local to the configuration
        del a0
        return ops.dense_fold(pack, self.coords, hi - lo, want_forces, want_stress)

    def packed_gradient_host(self, c, dtype, inv_scale=None):
        """PackedGradient (host, picklable) of configuration c, optionally divided by stdev."""
        lo, hi = self.config_slice(c)
        a0, a1 = int(self.atom_ptr[c]), int(self.atom_ptr[c + 1])
        n = hi - lo
        counts = self.counts_of(c).cpu().numpy()
        ptr = self.dz_ptr[lo:hi + 1].cpu().numpy()
        raw = self.dz[int(ptr[0]):int(ptr[-1])].to(torch.float64).cpu().numpy()
        if n:
            # blocks are [D][n_t + 1][3], possibly followed by alignment padding: gather the used parts and
            # scale row d by 1/stdev[d] — index arithmetic on the host, no per-atom launches
            rowlen = 3 * (counts.astype(np.int64) + 1)
            lens = self.D * rowlen
            first = np.cumsum(lens) - lens
            within = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(first, lens)
            vals = raw[np.repeat((ptr[:-1] - ptr[0]).astype(np.int64), lens) + within]
            if inv_scale is not None:
                inv = inv_scale.detach().to(torch.float64).cpu().numpy()
                vals = vals * inv[within // np.repeat(rowlen, lens)]
        else:
            vals = np.zeros(0)
        o0, o1 = int(self.offsets[a0]), int(self.offsets[a0 + n])
        nbrs = (self.indices[o0:o1] - a0).cpu().numpy().astype(np.int32)
        image = (self.image[a0:a1] - lo).cpu().numpy().astype(np.int32)
        return PackedGradient(vals.astype(dtype), counts.astype(np.int32), nbrs, image, self.D)

    @classmethod
    def from_packed_gradients(cls, D, device, grads, dz_dtype=torch.float64):
        """Rebuild the device store from PackedGradient objects (fingerprints loaded from disk).
        Coordinates are not needed for the force contraction and are left empty."""
        self = cls(D, device)
        atom_base = cen_base = 0
        counts, indices, image, centers, vals, sizes = [], [], [], [], [], []
        atom_ptr, cfg_ptr = [0], [0]
        for g in grads:
            N, A = len(g.counts), len(g.image)
            full = np.zeros(A, np.int64)
            full[:N] = g.counts
            counts.append(full)
            indices.append(g.neighbors.astype(np.int64) + atom_base)
            image.append(g.image.astype(np.int64) + cen_base)
            centers.append(np.arange(atom_base, atom_base + N))
            vals.append(np.asarray(g.values, np.float64))
            sizes.append(3 * D * (np.asarray(g.counts, np.int64) + 1))
            atom_base += A
            cen_base += N
            atom_ptr.append(atom_base)
            cfg_ptr.append(cen_base)
        cnt = np.concatenate(counts)
        off = np.zeros(len(cnt) + 1, np.int64)
        np.cumsum(cnt, out=off[1:])
        sz = np.concatenate(sizes)
        padded = (sz + 15) // 16 * 16
        ptr = np.zeros(len(sz) + 1, np.int64)
        np.cumsum(padded, out=ptr[1:])
        flat = np.zeros(int(ptr[-1]))
        src = np.concatenate(vals) if vals else np.zeros(0)
        spos = np.zeros(len(sz) + 1, np.int64)
        np.cumsum(sz, out=spos[1:])
        for t in range(len(sz)):
            flat[ptr[t]:ptr[t] + sz[t]] = src[spos[t]:spos[t + 1]]
        self.offsets = torch.as_tensor(off, device=device)
        self.indices = torch.as_tensor(np.concatenate(indices).astype(np.int32), device=device)
        self.image = torch.as_tensor(np.concatenate(image).astype(np.int32), device=device)
        self.centers = torch.as_tensor(np.concatenate(centers).astype(np.int32), device=device)
        self.dz = torch.as_tensor(flat, device=device).to(dz_dtype)
        self.dz_ptr = torch.as_tensor(ptr, device=device)
        self.atom_ptr = np.asarray(atom_ptr, np.int64)
        self.cfg_ptr = np.asarray(cfg_ptr, np.int64)
        self.maxn = int(cnt.max()) if len(cnt) else 0
        return self

    # ------------------------------------------------------------------ force contraction
    def forces(self, dEdG, lo, hi):
        """Flat forces f64 [3*(hi-lo)] of the contributing atoms [lo, hi), differentiable w.r.t.
        dEdG (shape [hi-lo, D]): F = -sum dEdG * scale * dG/dr."""
        return ops.force_scatter(dEdG, self.pack_for(lo, hi))
