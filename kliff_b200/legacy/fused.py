"""Fused training step for small tanh MLPs (csrc/train.cu) behind `Loss.minimize`.

What the reference does per batch — `CalculatorTorch.compute` (calculator_torch.py:161-228),
`Loss._get_loss_batch` (loss.py:843-920), `loss.backward()`, `optimizer.step()` — costs ~150 torch
micro-kernels for a 51-10-10-1 network (forward, double backward through dE/dzeta, index_add,
repeat_interleave, Adam ...).  Here one step is nine launches of hand-written kernels on flat buffers:

    forces.zero_ -> kb_mlp_energy_grad -> kb_force_scatter_fwd_seg (2 launches) -> kb_train_residual
                 -> kb_force_scatter_bwd -> kb_mlp_param_grad (3 launches) -> [all-reduce] -> kb_adam_step

and a whole epoch (every batch, gradient all-reduce included) is ONE CUDA graph.

Eligible models: `Linear -> Tanh -> ... -> Linear(out=1)` (<= 4 hidden layers of width <= 32, fan-in
<= 128; `Dropout(p=0)` layers are ignored), default residual function, Adam without weight decay /
amsgrad, energy and/or forces (no stress), fingerprints held as a packed GPU batch.  Anything else
stays on the torch route of `loss.py`.  The model's parameters become views into one flat buffer, so
`model.state_dict()`, `save()` and `write_kim_model()` see the trained values at any time.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops
from .._lib import check
from ..utils import capture_without_gc


def mlp_structure(model):
    """Hidden widths [H_1, ..., H_L] and fan-in when `model.layers` is a supported MLP, else None."""
    layers = [m for m in getattr(model, "layers", []) if not (isinstance(m, nn.Dropout) and m.p == 0.0)]
    if len(layers) < 3 or len(layers) % 2 == 0:
        return None
    widths, fan_in, lin = [], None, []
    for k, m in enumerate(layers):
        if k % 2 == 0:
            if not isinstance(m, nn.Linear) or m.bias is None:
                return None
            if fan_in is None:
                fan_in = m.in_features
            elif m.in_features != widths[-1]:
                return None
            widths.append(m.out_features)
            lin.append(m)
        elif not isinstance(m, nn.Tanh):
            return None
    if widths[-1] != 1:
        return None
    hidden = widths[:-1]
    if not (1 <= len(hidden) <= _lib.KB_MLP_MAX_HIDDEN and max(hidden) <= _lib.KB_MLP_MAX_WIDTH
            and fan_in <= _lib.KB_MLP_MAX_IN):
        return None
    return fan_in, hidden, lin


class FusedMLP:
    """Flat-buffer view of an eligible model + the per-step launches."""

    def __init__(self, model, linears, fan_in, hidden):
        self.model = model
        dev, dt = model.device, model.dtype
        self.dev, self.dt = dev, dt
        self.desc = _lib.kb_mlp_desc()
        self.desc.n_in, self.desc.n_hidden = fan_in, len(hidden)
        for l, h in enumerate(hidden):
            self.desc.width[l] = h
        self.desc.dtype = _lib.KB_F32 if dt == torch.float32 else _lib.KB_F64
        npar, avec = C.c_int32(0), C.c_int32(0)
        check(_lib.lib().kb_mlp_num_params(C.byref(self.desc), C.byref(npar), C.byref(avec)), "kb_mlp_num_params")
        self.P, self.avec = npar.value, avec.value
        # one flat parameter buffer; the module's parameters become views into it
        params = []
        for m in linears:
            params += [m.weight, m.bias]
        assert sum(p.numel() for p in params) == self.P
        self.theta = torch.empty(self.P, dtype=dt, device=dev)
        pos = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                self.theta[pos:pos + n].copy_(p.detach().reshape(-1))
                p.data = self.theta[pos:pos + n].view_as(p)
                pos += n
        self.params = params

    # -- kernels ------------------------------------------------------------------------------
    def energy_grad(self, x, e_atom, dEdx):
        check(_lib.lib().kb_mlp_energy_grad(C.byref(self.desc), x.shape[0], ops._ptr(x), ops._ptr(self.theta),
                                            ops._ptr(e_atom), ops._ptr(dEdx), ops._stream()), "kb_mlp_energy_grad")

    def param_grad(self, x, a, G, atom_vecs, partials, grad):
        check(_lib.lib().kb_mlp_param_grad(C.byref(self.desc), x.shape[0], ops._ptr(x), ops._ptr(self.theta),
                                           ops._ptr(a), ops._ptr(G), ops._ptr(atom_vecs), ops._ptr(partials),
                                           ops._ptr(grad), ops._stream()), "kb_mlp_param_grad")


class FusedTrainer:
    """Training passes of `LossNeuralNetworkModel` on the fused kernels.  `batches` is the list of this
    rank's (c0, c1, nglobal) configuration runs, one per global batch, in order."""

    def __init__(self, loss, lr, betas=(0.9, 0.999), eps=1e-8):
        calc = loss.calculator
        st = mlp_structure(calc.model)
        if st is None:
            raise ValueError("model is not a supported tanh MLP")
        fan_in, hidden, linears = st
        self.loss, self.calc = loss, calc
        self.net = FusedMLP(calc.model, linears, fan_in, hidden)
        dev, dt = self.net.dev, self.net.dt
        P = calc._packed
        if P.zeta_normalized.shape[1] != fan_in:
            raise ValueError("descriptor size does not match the first layer")
        self.x_all = P.zeta_normalized.to(dt).contiguous()
        self.cfg_ptr = torch.as_tensor(np.asarray(P.cfg_ptr, np.int64), device=dev)
        W = loss._weights_on_device()
        self.w_e = W["e"].to(torch.float64).contiguous()
        self.w_f = W["f"].to(torch.float64).contiguous()
        ref = calc._ref
        self.e_ref = ref["energy"].to(torch.float64).contiguous()
        self.f_ref = ref["forces"].to(torch.float64).contiguous() if calc.use_forces else None
        self.use_energy, self.use_forces = bool(calc.use_energy), bool(calc.use_forces)
        self.lr, self.b1, self.b2, self.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        npar = self.net.P
        self.m = torch.zeros(npar, dtype=dt, device=dev)
        self.v = torch.zeros(npar, dtype=dt, device=dev)
        self.step_count = torch.zeros((), dtype=torch.float64, device=dev)
        self.grad = torch.zeros(npar + 1, dtype=torch.float64, device=dev)  # [P] = this step's loss
        self.acc = torch.zeros((), dtype=torch.float64, device=dev)          # epoch loss
        self._sized = 0
        self._graphs = {}
        self._pool = None
        # four-launch form (csrc/train_fast.cu) when the network and the neighbor lists are inside its domain
        self.fast = (os.environ.get("KB_TRAIN_FAST", "1") != "0"
                     and _lib.lib().kb_train_fast_ok(C.byref(self.net.desc), int(P.maxn)) == 0)
        if self.fast:
            rows = 4 * torch.cuda.get_device_properties(dev).multi_processor_count
            self.partials_fast = torch.empty((rows, npar), dtype=torch.float64, device=dev)
            self.loss_cfg = torch.zeros(1, dtype=torch.float64, device=dev)
        self.peer = None
        if self.fast and os.environ.get("KB_PEER_ALLREDUCE", "1") != "0":
            self._setup_peer()

    def _setup_peer(self):
        """Workspaces of the peer-memory gradient all-reduce (csrc/peer.cu): one cudaMalloc'ed buffer per
        rank, mapped into every other rank of the node through CUDA IPC.  Collective; every rank ends up
        with the same answer (`self.peer` set, or None: the NCCL all-reduce stays)."""
        import socket
        from .loss import _dist
        dist = _dist()
        if dist is None or dist.get_backend() != "nccl":
            return
        W, me = dist.get_world_size(), dist.get_rank()
        L = _lib.lib()
        nbytes, own, handle = C.c_int64(0), C.c_void_p(0), C.create_string_buffer(64)
        ok = W <= 16 and L.kb_peer_ws_bytes(W, self.net.P + 1, C.byref(nbytes), None) == 0
        with torch.cuda.device(self.net.dev):
            ok = ok and L.kb_peer_alloc(nbytes.value, C.byref(own), handle) == 0
            gathered = [None] * W
            dist.all_gather_object(gathered, (bool(ok), handle.raw, socket.gethostname()))
            ok = all(g[0] for g in gathered) and len({g[2] for g in gathered}) == 1
            ptrs, opened = (C.c_void_p * W)(), []
            if ok:
                for r in range(W):
                    if r == me:
                        ptrs[r] = own.value
                        continue
                    p = C.c_void_p(0)
                    if L.kb_peer_open(gathered[r][1], C.byref(p)) != 0:
                        ok = False
                        break
                    ptrs[r] = p.value
                    opened.append(p)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.net.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag) == 1:
                self.peer = dict(ptrs=ptrs, world=W, rank=me, own=own, opened=opened)
                return
            for p in opened:
                L.kb_peer_close(p)
            if own.value:
                L.kb_peer_free(own)

    def release_peer(self):
        """Give the peer-memory workspaces back (collective over the ranks that set them up): importers
        unmap first, the owners free after a barrier — CUDA IPC forbids freeing an exported buffer that a
        peer still has open.  The captured graphs hold the raw pointers, so they go too."""
        if self.peer is None:
            return
        from .loss import _dist
        pr, self.peer = self.peer, None
        self._graphs = {}
        L = _lib.lib()
        with torch.cuda.device(self.net.dev):
            torch.cuda.synchronize()
            for p in pr["opened"]:
                L.kb_peer_close(p)
            dist = _dist()
            if dist is not None:
                dist.barrier()
            L.kb_peer_free(pr["own"])

    def _check_peer(self):
        if self.peer is None:
            return
        err = C.c_int32(0)
        check(_lib.lib().kb_peer_error(self.peer["own"], C.byref(err), ops._stream()), "kb_peer_error")
        if err.value:
            raise RuntimeError("kliff_b200: a rank did not arrive at the peer-memory gradient exchange within 4 s "
                               "(ranks out of step, or a peer died); KB_PEER_ALLREDUCE=0 selects the NCCL all-reduce")

    def _buffers(self, natoms):
        if natoms <= self._sized:
            return
        dev, D = self.net.dev, self.net.desc.n_in
        n = int(natoms)
        self.e_atom = torch.empty(n, dtype=torch.float64, device=dev)
        self.dEdx = torch.empty((n, D), dtype=torch.float64, device=dev)
        self.F = torch.empty(3 * n, dtype=torch.float64, device=dev)
        self.a = torch.empty(n, dtype=torch.float64, device=dev)
        self.gF = torch.empty(3 * n, dtype=torch.float64, device=dev)
        self.G = torch.empty((n, D), dtype=torch.float64, device=dev)
        self.avecs = torch.empty((n, self.net.avec), dtype=self.net.dt, device=dev)
        self.partials = torch.empty(((n + 31) // 32, self.net.P), dtype=torch.float64, device=dev)
        self._sized = n

    def _slot_buffer(self, nslots):
        """Scratch of the owner-sorted force assembly (3 doubles per slot of the largest batch)."""
        if nslots > getattr(self, "_slots_sized", 0):
            self.contrib = torch.empty(3 * int(nslots), dtype=torch.float64, device=self.net.dev)
            self._slots_sized = int(nslots)

    # ------------------------------------------------------------------------------------------
    def _step_fast(self, c0, c1, nglobal, train):
        """`step` as six launches (csrc/train_fast.cu): MLP forward, contraction, forces + residuals per
        configuration, adjoint, MLP backward + gradient partials, reduction (+ Adam when there is no all-reduce)."""
        P = self.calc._packed
        L = _lib.lib()
        net, g, uf = self.net, self.grad, self.use_forces
        lo, hi = int(P.cfg_ptr[c0]), int(P.cfg_ptr[c1])
        n, ncfg = hi - lo, c1 - c0
        self._buffers(max(n, 1))
        if ncfg > self.loss_cfg.shape[0]:
            self.loss_cfg = torch.zeros(ncfg, dtype=torch.float64, device=net.dev)
        st = ops._stream()
        pack = plan = None
        x = self.x_all[lo:hi]
        if n > 0:
            if uf:
                pack = P.pack_for(lo, hi)
                plan = pack["plan"]
                self._slot_buffer(max(plan["n_slots"], 1))
                check(L.kb_train_forward(C.byref(net.desc), n, ops._ptr(x), ops._ptr(net.theta),
                                         ops._ptr(pack.get("scale")), ops._ptr(pack["dz"]), ops._dz_code(pack["dz"]),
                                         ops._ptr(pack["dz_ptr"]), ops._ptr(plan["slot_ptr"]), ops._ptr(self.e_atom),
                                         ops._ptr(self.dEdx), ops._ptr(self.contrib), st), "kb_train_forward")
            else:
                check(L.kb_train_forward(C.byref(net.desc), n, ops._ptr(x), ops._ptr(net.theta), None, None, 0,
                                         None, None, ops._ptr(self.e_atom), None, None, st), "kb_train_forward")
        have_f = uf and n > 0
        check(L.kb_train_residual_seg(
            ncfg, ops._ptr(self.cfg_ptr, 8 * c0), lo, ops._ptr(self.e_atom),
            ops._ptr(plan["seg_ptr"]) if have_f else None, ops._ptr(plan["seg_src"]) if have_f else None,
            plan["slot_base"] if have_f else 0, ops._ptr(self.contrib) if have_f else None,
            ops._ptr(self.e_ref, 8 * c0), ops._ptr(self.f_ref, 24 * lo) if have_f else None,
            ops._ptr(self.w_e, 8 * c0), ops._ptr(self.w_f, 8 * c0) if have_f else None, 1.0 / nglobal,
            int(self.use_energy), int(have_f), ops._ptr(self.a), ops._ptr(self.gF) if have_f else None,
            ops._ptr(self.F) if have_f else None, ops._ptr(self.loss_cfg),
            ops._ptr(self.step_count) if train else None, st), "kb_train_residual_seg")
        rows = C.c_int32(0)
        if train and n > 0:
            check(L.kb_train_backward(
                C.byref(net.desc), n, ops._ptr(x), ops._ptr(net.theta),
                ops._ptr(pack.get("scale")) if uf else None, ops._ptr(pack["dz"]) if uf else None,
                ops._dz_code(pack["dz"]) if uf else 0, ops._ptr(pack["dz_ptr"]) if uf else None,
                ops._ptr(plan["slot_ptr"]) if uf else None, ops._ptr(self.contrib) if uf else None,
                ops._ptr(self.a), ops._ptr(self.G), ops._ptr(self.partials_fast),
                self.partials_fast.shape[0], C.byref(rows), st), "kb_train_backward")
        from .loss import _dist
        dist = _dist()
        single = dist is None
        if train and not single and self.peer is not None:
            # reduction of the partials, all-reduce over NVLink peer memory and Adam in ONE launch
            pr = self.peer
            check(L.kb_train_update_peer(net.P, rows.value, ops._ptr(self.partials_fast), ncfg,
                                         ops._ptr(self.loss_cfg), ops._ptr(g), 1, ops._ptr(net.theta),
                                         ops._ptr(self.m), ops._ptr(self.v), ops._ptr(self.step_count), self.lr,
                                         self.b1, self.b2, self.eps, net.desc.dtype, ops._ptr(self.acc), pr["ptrs"],
                                         pr["world"], pr["rank"], st), "kb_train_update_peer")
            return
        check(L.kb_train_update(C.byref(net.desc), rows.value, ops._ptr(self.partials_fast), ncfg,
                                ops._ptr(self.loss_cfg), ops._ptr(g), int(train and single), ops._ptr(net.theta),
                                ops._ptr(self.m), ops._ptr(self.v), ops._ptr(self.step_count), self.lr, self.b1,
                                self.b2, self.eps, ops._ptr(self.acc) if single else None, st), "kb_train_update")
        if not single:
            dist.all_reduce(g if train else g[net.P:], op=dist.ReduceOp.SUM)
            if train:
                check(L.kb_train_adam(net.P, ops._ptr(g), ops._ptr(net.theta), ops._ptr(self.m), ops._ptr(self.v),
                                      ops._ptr(self.step_count), self.lr, self.b1, self.b2, self.eps, net.desc.dtype,
                                      ops._ptr(self.acc), st), "kb_train_adam")
            else:
                self.acc.add_(g[net.P])

    def step(self, c0, c1, nglobal, train):
        """One batch: this rank's configurations [c0, c1) of a global batch of `nglobal`."""
        if self.fast:
            return self._step_fast(c0, c1, nglobal, train)
        P = self.calc._packed
        L = _lib.lib()
        lo, hi = int(P.cfg_ptr[c0]), int(P.cfg_ptr[c1])
        n = hi - lo
        self._buffers(max(n, 1))
        g = self.grad
        g.zero_()
        if n > 0:
            x = self.x_all[lo:hi]
            uf = self.use_forces
            self.net.energy_grad(x, self.e_atom, self.dEdx if uf else None)
            if uf:
                pack = P.pack_for(lo, hi)
                F = self.F[:3 * n]
                F.zero_()
                plan = pack["plan"]  # owner-sorted assembly: forces (and the loss) are bit-reproducible
                self._slot_buffer(max(plan["n_slots"], 1))
                check(L.kb_force_scatter_fwd_seg(
                    n, ops._ptr(pack["centers"]), ops._ptr(pack["offsets"]), pack["D"], ops._ptr(self.dEdx),
                    ops._ptr(pack.get("scale")), ops._ptr(pack["dz"]), ops._dz_code(pack["dz"]),
                    ops._ptr(pack["dz_ptr"]), ops._ptr(plan["slot_ptr"]), plan["n_slots"], plan["slot_base"], n,
                    ops._ptr(plan["seg_ptr"]), ops._ptr(plan["seg_src"]), ops._ptr(self.contrib), ops._ptr(F),
                    ops._stream()), "kb_force_scatter_fwd_seg")
            check(L.kb_train_residual(c1 - c0, ops._ptr(self.cfg_ptr, 8 * c0), lo, ops._ptr(self.e_atom),
                                      ops._ptr(self.F) if uf else None, ops._ptr(self.e_ref, 8 * c0),
                                      ops._ptr(self.f_ref, 24 * lo) if uf else None, ops._ptr(self.w_e, 8 * c0),
                                      ops._ptr(self.w_f, 8 * c0) if uf else None, 1.0 / nglobal,
                                      int(self.use_energy), int(uf), ops._ptr(self.a),
                                      ops._ptr(self.gF) if uf else None, ops._ptr(g, 8 * self.net.P), ops._stream()),
                  "kb_train_residual")
            if train:
                if uf:
                    check(L.kb_force_scatter_bwd(n, ops._ptr(pack["centers"]), ops._ptr(pack["offsets"]),
                                                 ops._ptr(pack["indices"]), ops._ptr(pack["image"]), pack["D"],
                                                 ops._ptr(self.gF, -24 * lo), ops._ptr(pack.get("scale")),
                                                 ops._ptr(pack["dz"]), ops._dz_code(pack["dz"]), ops._ptr(pack["dz_ptr"]),
                                                 ops._ptr(self.G), ops._stream()), "kb_force_scatter_bwd")
                self.net.param_grad(x, self.a, self.G if uf else None, self.avecs, self.partials, g)
        from .loss import _dist
        dist = _dist()
        if dist is not None:
            dist.all_reduce(g if train else g[self.net.P:], op=dist.ReduceOp.SUM)
        if train:
            check(L.kb_adam_step(self.net.P, ops._ptr(self.net.theta), ops._ptr(g), 1.0, ops._ptr(self.m),
                                 ops._ptr(self.v), ops._ptr(self.step_count), self.lr, self.b1, self.b2, self.eps,
                                 self.net.desc.dtype, ops._stream()), "kb_adam_step")
        self.acc.add_(g[self.net.P])

    def run_pass(self, batches, train, use_graph=True):
        """One pass over `batches`; returns the epoch loss (sum of batch losses, loss.py:789-806)."""
        key = (train, tuple(batches))
        self.acc.zero_()
        if not use_graph:
            for c0, c1, ng in batches:
                self.step(c0, c1, ng, train)
            total = float(self.acc)
            if train:
                self._check_peer()
            return total
        g = self._graphs.get(key)
        if g is None:
            self._buffers(max(int(self.calc._packed.cfg_ptr[c1] - self.calc._packed.cfg_ptr[c0])
                              for c0, c1, _ in batches) if batches else 1)
            if self.use_forces and batches:
                Pk = self.calc._packed
                sp = Pk.slot_ptr_host()
                self._slot_buffer(max(int(sp[int(Pk.cfg_ptr[c1])] - sp[int(Pk.cfg_ptr[c0])]) for c0, c1, _ in batches))
            if self._pool is None:
                # one eager warm-up step on a side stream (library handles, NCCL communicator), moving nothing
                side = torch.cuda.Stream(device=self.net.dev)
                side.wait_stream(torch.cuda.current_stream(self.net.dev))
                with torch.cuda.stream(side):
                    if batches:
                        self.step(*batches[0], train=False)
                torch.cuda.current_stream(self.net.dev).wait_stream(side)
                torch.cuda.synchronize(self.net.dev)
                self.acc.zero_()
                self._pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            with capture_without_gc(), torch.cuda.graph(g, pool=self._pool):
                for c0, c1, ng in batches:
                    self.step(c0, c1, ng, train)
            self._graphs[key] = g
        g.replay()
        total = float(self.acc)
        if train:
            self._check_peer()
        return total

    def state_dict(self):
        return {"theta": self.net.theta.clone(), "m": self.m.clone(), "v": self.v.clone(),
                "step": self.step_count.clone(), "lr": self.lr, "betas": (self.b1, self.b2), "eps": self.eps}
