"""Tensor-level wrappers over the C-ABI (include/kliff_b200.h).

Everything here takes and returns CUDA torch tensors (torch is the owner of device memory and
streams — plumbing, not compute); every numerical step is a hand-written sm_100a kernel inside
libkliff_b200.so.  There is no CPU path: CPU tensors raise.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import KB_GROUP_WIDTH, KB_MAX_GROUPS, KB_MAX_SPECIES, KB_MAX_TWO, check, kb_symfun_params

KIND = {"g1": 1, "g2": 2, "g3": 3, "g4": 4, "g5": 5}
_CANON_ZETA = {1.0: 0, 2.0: 1, 4.0: 2, 16.0: 3}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, byte_offset=0):
    return C.c_void_p(t.data_ptr() + byte_offset) if t is not None else C.c_void_p(0)


def _dz_code(t):
    """KB_F64 / KB_F32 for a packed-gradient tensor (or a torch / numpy dtype)."""
    dt = t.dtype if torch.is_tensor(t) else t
    if dt in (torch.float64, np.float64):
        return _lib.KB_F64
    if dt in (torch.float32, np.float32):
        return _lib.KB_F32
    raise TypeError(f"packed gradients are float64 or float32, not {dt}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "kliff_b200 ops need CUDA tensors: there is no CPU implementation of the hot path"
            )


def device_info():
    sm, ma, mi = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    check(_lib.lib().kb_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "kb_device_info")
    return sm.value, ma.value, mi.value


# --------------------------------------------------------------------------- descriptor params
def build_params(hyperparams, rcut):
    """Hyper-parameter dict (insertion ordered; keys g1..g5 as in sym_fn.py:237-283) and the
    species-pair cutoff matrix -> kb_symfun_params.

    Descriptor index = running sum of rows in family order (sym_fn.cpp:96-100).  Angular rows are
    grouped by (family, eta); a group holds at most 8 (zeta, lambda) entries.
    """
    rcut = np.asarray(rcut, dtype=np.float64)
    S = rcut.shape[0]
    if rcut.shape != (S, S) or S > KB_MAX_SPECIES:
        raise ValueError(f"cutoff matrix must be square with at most {KB_MAX_SPECIES} species")
    P = kb_symfun_params()
    P.n_species = S
    for a in range(S):
        for b in range(S):
            P.rcut[a * S + b] = float(rcut[a, b])
    d = 0
    two = []      # (kind, p0, p1, out)
    groups = []   # dict(kind, eta, entries=[(zeta, lam, out)])
    for key, rows in hyperparams.items():
        kind = KIND[key.lower()]
        if kind == 1:
            two.append((1, 0.0, 0.0, d))
            d += 1
            continue
        for r in rows:
            if kind == 2:
                two.append((2, float(r["eta"]), float(r["Rs"]), d))
            elif kind == 3:
                two.append((3, float(r["kappa"]), 0.0, d))
            else:
                z, lam, eta = float(r["zeta"]), float(r["lambda"]), float(r["eta"])
                tgt = None
                for g in groups:
                    if (g["kind"] == kind and g["eta"] == eta and len(g["entries"]) < KB_GROUP_WIDTH
                            and all((z, lam) != (e[0], e[1]) for e in g["entries"])):
                        tgt = g
                        break
                if tgt is None:
                    tgt = {"kind": kind, "eta": eta, "entries": []}
                    groups.append(tgt)
                tgt["entries"].append((z, lam, d))
            d += 1
    if len(two) > KB_MAX_TWO or len(groups) > KB_MAX_GROUPS:
        raise ValueError("too many symmetry-function parameter sets for the compiled-in limits")
    P.n_desc = d
    P.n_two = len(two)
    for q, (k, p0, p1, out) in enumerate(two):
        P.two_kind[q], P.two_p0[q], P.two_p1[q], P.two_out[q] = k, p0, p1, out
    canonical = all(z in _CANON_ZETA and lam in (-1.0, 1.0)
                    for g in groups for (z, lam, _) in g["entries"])
    P.n_groups = len(groups)
    P.canonical_zl = 1 if (canonical and groups) else 0
    for gi, g in enumerate(groups):
        P.grp_kind[gi] = g["kind"]
        P.grp_eta[gi] = g["eta"]
        for e in range(KB_GROUP_WIDTH):
            P.grp_out[gi][e] = -1
            P.grp_zeta[gi][e] = 1.0
            P.grp_lambda[gi][e] = 1.0
        if canonical:
            P.grp_n[gi] = KB_GROUP_WIDTH
            for e in range(KB_GROUP_WIDTH):
                P.grp_zeta[gi][e] = (1.0, 2.0, 4.0, 16.0)[e // 2]
                P.grp_lambda[gi][e] = (-1.0, 1.0)[e % 2]
            for z, lam, out in g["entries"]:
                P.grp_out[gi][2 * _CANON_ZETA[z] + (1 if lam > 0 else 0)] = out
        else:
            P.grp_n[gi] = len(g["entries"])
            for e, (z, lam, out) in enumerate(g["entries"]):
                P.grp_zeta[gi][e], P.grp_lambda[gi][e], P.grp_out[gi][e] = z, lam, out
    return P


# --------------------------------------------------------------------------- paddings
def create_paddings(cutoff, cell, pbc, coords, species=None):
    """neighlist.create_paddings on the GPU.  coords: cuda f64 [N,3]; species: cuda i32 [N] or
    None.  Returns (pad_coords f64[P,3], pad_species i32[P] | None, pad_image i32[P])."""
    _need_cuda(coords, species)
    coords = coords.contiguous()
    if coords.dtype != torch.float64:
        raise TypeError("coords must be float64")
    n = coords.shape[0]
    cell_h = (C.c_double * 9)(*np.asarray(cell, dtype=np.float64).reshape(9))
    pbc_h = (C.c_int32 * 3)(*[int(bool(p)) for p in pbc])
    npad = C.c_int64(0)
    plan = C.c_void_p(0)
    L = _lib.lib()
    check(L.kb_pad_begin(n, float(cutoff), cell_h, pbc_h, _ptr(coords), C.byref(npad),
                         C.byref(plan), _stream()), "kb_pad_begin")
    P = npad.value
    dev = coords.device
    pc = torch.empty((P, 3), dtype=torch.float64, device=dev)
    pi = torch.empty((P,), dtype=torch.int32, device=dev)
    ps = None
    if species is not None:
        species = species.contiguous()
        ps = torch.empty((P,), dtype=torch.int32, device=dev)
    check(L.kb_pad_finish(plan, _ptr(species), _ptr(pc), _ptr(ps), _ptr(pi), _stream()),
          "kb_pad_finish")
    return pc, ps, pi


# --------------------------------------------------------------------------- neighbor list
def build_neighbors(coords, n_need, infl_dist, cutoff=None, need=None, half=False):
    """NeighList.build on the GPU.  coords: cuda f64 [Ntot,3] (contributing atoms first).
    Returns (counts i32[Ntot], offsets i64[Ntot+1], indices i32[total] sorted ascending per atom,
    maxn)."""
    _need_cuda(coords, need)
    coords = coords.contiguous()
    if coords.dtype != torch.float64:
        raise TypeError("coords must be float64")
    ntot = coords.shape[0]
    dev = coords.device
    counts = torch.empty((ntot,), dtype=torch.int32, device=dev)
    offsets = torch.empty((ntot + 1,), dtype=torch.int64, device=dev)
    total, maxn, plan = C.c_int64(0), C.c_int32(0), C.c_void_p(0)
    L = _lib.lib()
    check(L.kb_nl_begin(ntot, int(n_need), _ptr(coords), _ptr(need), float(infl_dist),
                        float(infl_dist if cutoff is None else cutoff), int(bool(half)),
                        _ptr(counts), _ptr(offsets), C.byref(total), C.byref(maxn), C.byref(plan),
                        _stream()), "kb_nl_begin")
    indices = torch.empty((total.value,), dtype=torch.int32, device=dev)
    check(L.kb_nl_finish(plan, _ptr(indices), _stream()), "kb_nl_finish")
    return counts, offsets, indices, maxn.value


def build_neighbors_multi(coords, n_need, infl_dist, cutoffs, need=None, half=False):
    """`NeighList.build(coords, influence_distance, cutoffs[k], need_neigh)` with several cutoffs
    (neighbor_list.cpp:148-160, 210-217: one binning by the influence distance, one list per cutoff,
    each keeping the pairs with rsq < cutoff_k²).  Returns a list of k (counts, offsets, indices, maxn)
    tuples in the order of `cutoffs` — list k is what `get_neigh(cutoffs, k, i)` serves
    (neighbor_list_bind.cpp:78-146).  One kb_nl_begin / kb_nl_finish pair per cutoff over the same
    coordinates; the reference's Python layer only ever asks for one (neighbor.py:107)."""
    cutoffs = [float(c) for c in np.atleast_1d(np.asarray(cutoffs, dtype=np.float64))]
    if not cutoffs:
        raise ValueError("build_neighbors_multi: at least one cutoff is needed")
    if max(cutoffs) > float(infl_dist):
        # the 27-cell scan only covers the influence distance (the reference would silently miss pairs)
        raise ValueError("build_neighbors_multi: every cutoff must be <= influence_distance")
    return [build_neighbors(coords, n_need, infl_dist, cutoff=c, need=need, half=half) for c in cutoffs]


def batch_neighbors(cfg_ptr, cells, pbcs, coords, cutoff):
    """Paddings + neighbor lists of many configurations in one pass (kb_batch_*).

    cfg_ptr: host int [C+1] contributing ranges into coords (cuda f64 [N,3]); cells: host
    [C,3,3]; pbcs: host [C,3].  Returns dict(coords, image, centers, atom_ptr (host int64),
    offsets, indices, maxn) in the packed batch layout (packed.py)."""
    _need_cuda(coords)
    coords = coords.contiguous()
    if coords.dtype != torch.float64:
        raise TypeError("coords must be float64")
    cfg_ptr = np.ascontiguousarray(cfg_ptr, dtype=np.int32)
    cells = np.ascontiguousarray(cells, dtype=np.float64).reshape(-1)
    pbcs = np.ascontiguousarray(np.asarray(pbcs).astype(bool), dtype=np.int32).reshape(-1)
    n_cfg = len(cfg_ptr) - 1
    if len(cells) != 9 * n_cfg or len(pbcs) != 3 * n_cfg or int(cfg_ptr[-1]) != coords.shape[0]:
        raise ValueError("batch_neighbors: inconsistent configuration arrays")
    dev = coords.device
    L = _lib.lib()
    total_atoms, plan = C.c_int64(0), C.c_void_p(0)
    check(L.kb_batch_begin(n_cfg, cfg_ptr.ctypes.data_as(C.c_void_p), cells.ctypes.data_as(C.c_void_p),
                           pbcs.ctypes.data_as(C.c_void_p), float(cutoff), _ptr(coords),
                           C.byref(total_atoms), C.byref(plan), _stream()), "kb_batch_begin")
    A, N = total_atoms.value, coords.shape[0]
    coords_all = torch.empty((A, 3), dtype=torch.float64, device=dev)
    image = torch.empty((A,), dtype=torch.int32, device=dev)
    centers = torch.empty((N,), dtype=torch.int32, device=dev)
    atom_ptr = torch.empty((n_cfg + 1,), dtype=torch.int64, device=dev)
    counts = torch.empty((A,), dtype=torch.int32, device=dev)
    offsets = torch.zeros((A + 1,), dtype=torch.int64, device=dev)
    pairs, maxn = C.c_int64(0), C.c_int32(0)
    rc = L.kb_batch_neighbors(plan, _ptr(coords_all), _ptr(image), _ptr(centers), _ptr(atom_ptr),
                              _ptr(counts), _ptr(offsets), C.byref(pairs), C.byref(maxn), _stream())
    if rc != _lib.KB_OK:
        L.kb_batch_destroy(plan, _stream())
        check(rc, "kb_batch_neighbors")
    indices = torch.empty((pairs.value,), dtype=torch.int32, device=dev)
    check(L.kb_batch_finish(plan, _ptr(coords_all), _ptr(offsets), _ptr(indices), _stream()),
          "kb_batch_finish")
    return dict(coords=coords_all, image=image, centers=centers,
                atom_ptr=atom_ptr.cpu().numpy() if n_cfg else np.zeros(1, np.int64),
                offsets=offsets, indices=indices, maxn=maxn.value)


# --------------------------------------------------------------------------- descriptor
def block_offsets(n_centers, centers, offsets, n_desc):
    dev = offsets.device
    dz_ptr = torch.empty((n_centers + 1,), dtype=torch.int64, device=dev)
    total = C.c_int64(0)
    check(_lib.lib().kb_symfun_block_offsets(n_centers, _ptr(centers), _ptr(offsets), n_desc,
                                             _ptr(dz_ptr), C.byref(total), _stream()),
          "kb_symfun_block_offsets")
    return dz_ptr, total.value


def symfun_forward(params, coords, species, offsets, indices, maxn, n_centers=None, centers=None,
                   grad=True, mode=_lib.KB_SYMFUN_AUTO, dz_ptr=None, dz=None, dz_dtype=torch.float64,
                   zeta_out=None):
    """Descriptors (and packed dG/dr) of centre atoms.  Returns (zeta f64[n_centers, D],
    dz [total] | None, dz_ptr i64[n_centers+1] | None).  Arithmetic is fp64; `dz_dtype` float32
    rounds the finished blocks on the way out (half the memory and contraction traffic).
    `zeta_out` (contiguous cuda f64 [n_centers, D], e.g. a row range of a larger array) receives the
    descriptors instead of a fresh allocation; with `dz_ptr` a slice of a batch-wide offset array
    this lets a caller run a long centre list in pieces (PackedFingerprints.compute)."""
    _need_cuda(coords, species, offsets, indices, centers)
    if centers is not None:
        n_centers = centers.shape[0]
    if n_centers is None:
        raise ValueError("n_centers or centers required")
    D = params.n_desc
    dev = coords.device
    if zeta_out is not None:
        if (not zeta_out.is_cuda or zeta_out.dtype != torch.float64 or not zeta_out.is_contiguous()
                or tuple(zeta_out.shape) != (n_centers, D)):
            raise ValueError("zeta_out must be a contiguous cuda float64 [n_centers, D] tensor")
        zeta = zeta_out
    else:
        zeta = torch.empty((n_centers, D), dtype=torch.float64, device=dev)
    if grad:
        if dz_ptr is None:
            dz_ptr, total = block_offsets(n_centers, centers, offsets, D)
            dz = torch.empty((total,), dtype=dz_dtype, device=dev)
    else:
        dz, dz_ptr = None, None
    if indices.numel() == 0:
        indices = torch.zeros((1,), dtype=torch.int32, device=dev)
    check(_lib.lib().kb_symfun_forward(C.byref(params), n_centers, _ptr(centers), _ptr(coords),
                                       _ptr(species), _ptr(offsets), _ptr(indices), int(maxn),
                                       _ptr(zeta), _ptr(dz), _dz_code(dz) if dz is not None else 0,
                                       _ptr(dz_ptr), int(mode), _stream()),
          "kb_symfun_forward")
    return zeta, dz, dz_ptr


# --------------------------------------------------------------------------- force assembly
class _ForceScatter(torch.autograd.Function):
    """F = -sum_{t,d} dEdG[t,d] * scale[d] * dz_t[d, slot, c] folded through `image`.  Linear in
    dEdG, so the backward (w.r.t. dEdG) is the adjoint contraction and is itself differentiable
    (double backward = forward), which `create_graph=True` training needs."""

    @staticmethod
    def forward(ctx, dEdG, pack):
        ctx.pack = pack
        return _scatter_fwd(dEdG, pack)

    @staticmethod
    def backward(ctx, gF):
        return _ForceGather.apply(gF, ctx.pack), None


class _ForceGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gF, pack):
        ctx.pack = pack
        return _scatter_bwd(gF, pack)

    @staticmethod
    def backward(ctx, gg):
        return _ForceScatter.apply(gg, ctx.pack), None


def scatter_plan(offsets, indices, image, n_out, centers=None, n_centers=None):
    """Owner-sorted plan of the deterministic force assembly (kb_scatter_plan_*): returns
    dict(slot_ptr i64[n_centers+1], seg_ptr i64[n_out+1], seg_src i32[slots], n_slots)."""
    _need_cuda(offsets, indices, image)
    n_centers = int(centers.shape[0]) if centers is not None else int(n_centers)
    dev = offsets.device
    slot_ptr = torch.empty(n_centers + 1, dtype=torch.int64, device=dev)
    total = C.c_int64(0)
    check(_lib.lib().kb_scatter_plan_begin(n_centers, _ptr(centers), _ptr(offsets), _ptr(slot_ptr),
                                           C.byref(total), _stream()), "kb_scatter_plan_begin")
    seg_ptr = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
    seg_src = torch.empty(max(total.value, 1), dtype=torch.int32, device=dev)
    check(_lib.lib().kb_scatter_plan_finish(n_centers, _ptr(centers), _ptr(offsets), _ptr(indices), _ptr(image),
                                            n_out, _ptr(slot_ptr), total.value, _ptr(seg_ptr), _ptr(seg_src),
                                            _stream()), "kb_scatter_plan_finish")
    return dict(slot_ptr=slot_ptr, seg_ptr=seg_ptr, seg_src=seg_src, n_slots=total.value)


def _scatter_fwd(dEdG, pack):
    _need_cuda(dEdG)
    dEdG = dEdG.contiguous().to(torch.float64)
    n_centers = dEdG.shape[0]
    F = torch.zeros((pack["n_out"] * 3,), dtype=torch.float64, device=dEdG.device)
    idx = pack["indices"]
    if pack.get("deterministic", True):
        # owner-sorted reduction (bit-reproducible); the plan is integer work, cached with the pack
        plan = pack.get("plan")
        if plan is None:
            if int(pack.get("out_base", 0)) != 0:
                raise ValueError("a pack for a sub-range of centres needs the plan of its parent (plan=...)")
            plan = scatter_plan(pack["offsets"], idx, pack["image"], pack["n_out"], pack.get("centers"), n_centers)
            plan.update(slot_base=0)
            pack["plan"] = plan
        check(_lib.lib().kb_force_scatter_fwd_seg(
            n_centers, _ptr(pack.get("centers")), _ptr(pack["offsets"]), pack["D"], _ptr(dEdG),
            _ptr(pack.get("scale")), _ptr(pack["dz"]), _dz_code(pack["dz"]), _ptr(pack["dz_ptr"]),
            _ptr(plan["slot_ptr"]), plan["n_slots"], plan["slot_base"], pack["n_out"], _ptr(plan["seg_ptr"]),
            _ptr(plan["seg_src"]), None, _ptr(F), _stream()), "kb_force_scatter_fwd_seg")
        return F
    check(_lib.lib().kb_force_scatter_fwd(n_centers, _ptr(pack.get("centers")), _ptr(pack["offsets"]),
                                          _ptr(idx), _ptr(pack["image"]), pack["D"], _ptr(dEdG),
                                          _ptr(pack.get("scale")), _ptr(pack["dz"]), _dz_code(pack["dz"]),
                                          _ptr(pack["dz_ptr"]),
                                          _ptr(F, -24 * int(pack.get("out_base", 0))), _stream()),
          "kb_force_scatter_fwd")
    return F


def _scatter_bwd(gF, pack):
    _need_cuda(gF)
    gF = gF.contiguous().to(torch.float64)
    n_centers = pack["dz_ptr"].shape[0] - 1
    out = torch.empty((n_centers, pack["D"]), dtype=torch.float64, device=gF.device)
    check(_lib.lib().kb_force_scatter_bwd(n_centers, _ptr(pack.get("centers")), _ptr(pack["offsets"]),
                                          _ptr(pack["indices"]), _ptr(pack["image"]), pack["D"],
                                          _ptr(gF, -24 * int(pack.get("out_base", 0))),
                                          _ptr(pack.get("scale")), _ptr(pack["dz"]), _dz_code(pack["dz"]),
                                          _ptr(pack["dz_ptr"]), _ptr(out), _stream()),
          "kb_force_scatter_bwd")
    return out


def force_scatter(dEdG, pack):
    """pack: dict(offsets, indices, image, dz, dz_ptr, D, n_out[, centers, scale]).  Returns the
    flat force vector f64[3*n_out] (same layout as calculator_torch.py:266-269)."""
    return _ForceScatter.apply(dEdG, pack)


def dense_fold(pack, coords, n_contrib, want_forces=True, want_stress=False):
    """Dense (n_centers, D, 3*n_contrib) / (n_centers, D, 6) views of sym_fn.py:163-185."""
    n_centers = pack["dz_ptr"].shape[0] - 1
    D = pack["D"]
    dev = pack["dz"].device
    dense = torch.zeros((n_centers, D, 3 * n_contrib), dtype=torch.float64, device=dev) if want_forces else None
    stress = torch.empty((n_centers, D, 6), dtype=torch.float64, device=dev) if want_stress else None
    check(_lib.lib().kb_dense_fold(n_centers, _ptr(pack.get("centers")), _ptr(pack["offsets"]),
                                   _ptr(pack["indices"]), _ptr(pack["image"]), _ptr(coords), D,
                                   int(n_contrib), _ptr(pack["dz"]), _dz_code(pack["dz"]),
                                   _ptr(pack["dz_ptr"]), _ptr(dense), _ptr(stress), _stream()),
          "kb_dense_fold")
    return dense, stress


def moments(zeta, mean=None):
    """Column sums of zeta [n, D] (mean=None) or of (zeta - mean)^2: the two passes of np.mean / np.std
    (descriptor.py:116-118) as deterministic device reductions.  Returns f64 [D]."""
    _need_cuda(zeta)
    zeta = zeta.contiguous()
    out = torch.empty(zeta.shape[1], dtype=torch.float64, device=zeta.device)
    check(_lib.lib().kb_moments(zeta.shape[0], zeta.shape[1], _ptr(zeta), _ptr(mean), _ptr(out), _stream()),
          "kb_moments")
    return out


def normalize(zeta, mean, inv_stdev, out=None):
    """(zeta - mean) * inv_stdev (descriptor.py:187-194)."""
    _need_cuda(zeta, mean, inv_stdev)
    zeta = zeta.contiguous()
    out = torch.empty_like(zeta) if out is None else out
    check(_lib.lib().kb_normalize(zeta.shape[0], zeta.shape[1], _ptr(zeta), _ptr(mean.contiguous()),
                                  _ptr(inv_stdev.contiguous()), _ptr(out), _stream()), "kb_normalize")
    return out


def probe_fp64_peak():
    v = C.c_double(0.0)
    check(_lib.lib().kb_probe_fp64_peak(C.byref(v), _stream()), "kb_probe_fp64_peak")
    return v.value


def probe_hbm_copy(nbytes=1 << 30):
    v = C.c_double(0.0)
    check(_lib.lib().kb_probe_hbm_copy(int(nbytes), C.byref(v), _stream()), "kb_probe_hbm_copy")
    return v.value


def pool_trim(keep_bytes=0):
    """Give the library's cached scratch (stream-ordered pool) back to the driver; synchronises the device."""
    check(_lib.lib().kb_pool_trim(int(keep_bytes)), "kb_pool_trim")


def launch_count(reset=False):
    """Kernels launched by libkliff_b200 in this process so far."""
    return int(_lib.lib().kb_launch_count(1 if reset else 0))


def kernel_timing(enable):
    """Collects the per-kernel device times recorded since the last call and switches recording
    on/off.  Returns {"prep": (ms, launches), "sweep": (...), "fused": (...), "scatter": (...),
    "stage": (...)} — "stage" brackets the whole two-stage descriptor path of a call (all chunks)."""
    ms = (C.c_double * 5)()
    cnt = (C.c_int32 * 5)()
    check(_lib.lib().kb_kernel_timing(int(bool(enable)), ms, cnt), "kb_kernel_timing")
    return {name: (ms[q], cnt[q]) for q, name in enumerate(("prep", "sweep", "fused", "scatter", "stage"))}
