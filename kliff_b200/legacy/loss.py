"""`Loss` / `LossNeuralNetworkModel` with the interface of kliff/legacy/loss.py:37-223, 637-935
(the neural-network half; physics-model losses are out of scope).

Semantics kept: per configuration  residual = config_weight * (pred - ref), residual[0] *=
energy_weight, residual[1:] *= forces_weight, residual /= natoms; loss = sum residual^2;
batch loss = sum / len(batch); epoch loss = sum of batch losses; `optimizer.step(closure)`;
the same `Epoch = ...  loss = ...` print-out and the extra loss pass after the last epoch.

With the default residual function and a calculator that holds packed fingerprints the batch
loss is evaluated in vectorised form on the GPU (one MLP pass, one force-scatter launch, one
reduction) instead of the reference's per-configuration Python loop (loss.py:867-877); any
user-supplied `residual_fn` takes the reference's per-configuration route.

Data-parallel training: when torch.distributed is initialised every rank takes a contiguous
1/world share of each batch, and the parameter gradients (641 numbers for 51-10-10-1) and the
loss are all-reduced (NCCL over NVLink on the GPU box, gloo in the CPU tests) before the
optimizer step, so all ranks follow the single-GPU trajectory up to summation order.
"""
import os

import numpy as np
import torch

from ..utils import capture_without_gc
from .calculators.calculator_torch import CalculatorTorch


class LossError(Exception):
    def __init__(self, msg):
        super().__init__(msg)
        self.msg = msg


def _w(value, like):
    """A weight in a form the residual tensor can be multiplied by: numbers pass through; arrays
    (`MagnitudeInverseWeight.forces_weight`, one value per force component, kliff/dataset/weight.py:178-186)
    become tensors beside `like` — the reference multiplies a CPU tensor by the ndarray directly, which a
    CUDA tensor cannot do."""
    if np.ndim(value) == 0:
        return value
    return torch.as_tensor(np.asarray(value), dtype=like.dtype, device=like.device)


def energy_forces_residual(identifier, natoms, weight, prediction, reference, data):
    """loss.py:37-110."""
    residual = weight.config_weight * (prediction - reference)
    residual[0] *= weight.energy_weight
    residual[1:] *= _w(weight.forces_weight, residual)
    if data["normalize_by_natoms"]:
        residual /= natoms
    return residual


def energy_residual(identifier, natoms, weight, prediction, reference, data):
    """loss.py:113-138."""
    residual = weight.config_weight * weight.energy_weight * (prediction - reference)
    if data["normalize_by_natoms"]:
        residual /= natoms
    return residual


def forces_residual(identifier, natoms, weight, prediction, reference, data):
    """loss.py:141-166."""
    residual = weight.config_weight * _w(weight.forces_weight, prediction) * (prediction - reference)
    if data["normalize_by_natoms"]:
        residual /= natoms
    return residual


def _check_residual_data(data, default):
    if data is not None:
        for key, value in data.items():
            if key not in default:
                raise LossError(
                    f"Expect the keys of `residual_data` to be one or combinations of "
                    f"{', '.join(default.keys())}; got {key}. ")
            default[key] = value
    return default


class _Share(list):
    """This rank's members of one global batch (sharded fingerprints); `nglobal` = size of the batch."""
    nglobal = 0


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class Loss:
    """Factory (loss.py:169-223): torch calculators get the neural-network loss."""

    def __new__(cls, calculator, nprocs=1, residual_fn=None, residual_data=None,
                log_per_atom_pred=False, log_per_atom_pred_path=None):
        if isinstance(calculator, CalculatorTorch):
            return LossNeuralNetworkModel(calculator, nprocs, residual_fn, residual_data,
                                          log_per_atom_pred, log_per_atom_pred_path)
        raise LossError("kliff_b200 only implements the loss of torch (neural-network) calculators; "
                        "physics-motivated models are outside the accelerated path.")


class LossNeuralNetworkModel:
    torch_minimize_methods = ["Adadelta", "Adagrad", "Adam", "SparseAdam", "Adamax", "ASGD",
                              "LBFGS", "RMSprop", "Rprop", "SGD"]

    def __init__(self, calculator, nprocs=1, residual_fn=None, residual_data=None,
                 log_per_atom_pred=False, log_per_atom_pred_path=None):
        self.residual_data = _check_residual_data(residual_data, {"normalize_by_natoms": True})
        self.calculator = calculator
        self.nprocs = nprocs
        self._default_residual = residual_fn is None
        self.residual_fn = energy_forces_residual if residual_fn is None else residual_fn
        self.optimizer = None
        self.optimizer_state_path = None
        if log_per_atom_pred:
            raise LossError("log_per_atom_pred needs lmdb, which is not part of the accelerated path")
        self.log_predictions = False
        self._epoch = 0
        self.vectorized = True
        self.epoch_losses = []
        # One CUDA graph per batch (captured the first time the batch is met, replayed afterwards):
        # a 100-configuration step is ~150 tiny launches whose host-side issue time (1.5 ms) is 5x
        # their device time.  Single-GPU vectorised route only; KB_NO_GRAPHS=1 switches it off.
        self.use_cuda_graphs = os.environ.get("KB_NO_GRAPHS", "0") != "1"
        self._graphs = {}
        self._graph_pool = None
        self._graph_acc = None
        # Fused training step (legacy/fused.py, csrc/train.cu): nine hand-written launches per batch on
        # flat buffers, one CUDA graph per epoch.  Taken when the model is a small tanh MLP and the
        # optimizer is plain Adam; `loss.use_fused = False` or KB_NO_FUSED=1 keeps the torch route.
        self.use_fused = os.environ.get("KB_NO_FUSED", "0") != "1"
        self._fused = None

    # ------------------------------------------------------------------ minimize (loss.py:740-828)
    def minimize(self, method="Adam", batch_size=100, num_epochs=1000, start_epoch=0, **kwargs):
        dev = self.calculator.model.device
        if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):  # CalculatorTorch(model, gpu=<index>): everything on that device
                return self.minimize(method, batch_size, num_epochs, start_epoch, **kwargs)
        if method not in self.torch_minimize_methods:
            raise LossError(f"Minimization method `{method}` not supported.")
        self.batch_size, self.num_epochs, self.start_epoch = batch_size, num_epochs, start_epoch
        try:
            self.optimizer = getattr(torch.optim, method)(self.calculator.model.parameters(), **kwargs)
            if self.optimizer_state_path is not None:
                self._load_optimizer_stat(self.optimizer_state_path)
        except TypeError as e:
            print(str(e))
            idx = str(e).index("argument '") + 10
            err_arg = str(e)[idx:].strip("'")
            raise LossError(f"Argument `{err_arg}` not supported by optimizer `{method}`.")

        loader = self.calculator.get_compute_arguments(batch_size)
        dist = _dist()
        chief = dist is None or dist.get_rank() == 0
        self.epoch_losses = []
        self._graphs, self._graph_pool, self._graph_acc = {}, None, None  # graphs belong to one optimizer
        if self._fused is not None:
            self._fused.release_peer()  # collective, like minimize() itself: every rank replaces its trainer here
        self._fused = self._prepare_fused(method, kwargs)
        graphed = (self._fused is None and self._graphs_possible(method, kwargs)
                   and self._prepare_graphs(method, kwargs))
        epoch = 0
        for epoch in range(self.start_epoch, self.start_epoch + self.num_epochs):
            self._epoch = epoch
            if self.start_epoch != 0 and epoch == self.start_epoch:
                epoch_loss = self._get_loss_epoch(loader)
            elif self._fused is not None:
                epoch_loss = self._fused.run_pass(self._fused_batches(), train=True,
                                                  use_graph=self.use_cuda_graphs)
                if chief:
                    self.calculator.save_model(epoch)
            elif graphed and self._graphs is not None:
                try:
                    epoch_loss = self._graphed_pass(train=True)
                except RuntimeError as e:
                    # a capture-time failure (a layer with a host sync, a collective that cannot be
                    # captured ...): no step of this epoch has been applied yet when the FIRST capture
                    # fails; drop the graphs and train on eagerly with the same (capturable) optimizer
                    if any(k[1] for k in self._graphs):
                        raise
                    print(f"kliff_b200: CUDA-graph capture failed ({e}); continuing with eager steps")
                    self._graphs, self._graph_pool = None, None
                    epoch_loss = self._eager_epoch(loader)
                if chief:
                    self.calculator.save_model(epoch)
            else:
                epoch_loss = self._eager_epoch(loader)
                if chief:
                    self.calculator.save_model(epoch)
            self.epoch_losses.append(epoch_loss)
            if chief:
                print("Epoch = {:<6d}  loss = {:.10e}".format(epoch, epoch_loss))
        epoch += 1
        epoch_loss = self._get_loss_epoch(loader)
        self.epoch_losses.append(epoch_loss)
        if chief:
            print("Epoch = {:<6d}  loss = {:.10e}".format(epoch, epoch_loss))
            self.calculator.save_model(epoch, force_save=True)

    def _eager_epoch(self, loader):
        """One optimisation epoch with one optimizer.step(closure) per batch (loss.py:771-790)."""
        epoch_loss = 0
        for batch in self._batches(loader):

            def closure():
                self.optimizer.zero_grad()
                loss = self._get_loss_batch(batch)
                loss.backward()
                return self._sync_gradients(loss)

            loss = self.optimizer.step(closure)
            epoch_loss += float(loss.detach())
        return epoch_loss

    def _batches(self, loader):
        """Batches of samples.  On the vectorised route the DataLoader's per-sample collate
        (numpy -> tensor for every value) is skipped: only configuration indices are needed."""
        shard = getattr(self.calculator, "_shard", None)
        if shard is not None and not self._fast_path():
            raise LossError("fingerprints sharded over ranks need the default residual function; call "
                            "`calc.create(..., shard_configs=False)` to keep every configuration on every rank")
        if self._fast_path():
            bs = self.batch_size if getattr(self, "batch_size", None) else loader.batch_size
            fp = self.calculator.fingerprints_dataset.fp
            if shard is not None:
                # global batch [lo, hi) -> this rank's configurations that fall inside it: a contiguous run
                # of the local list (members are kept in ascending global order), see _shard_below
                below = self.calculator._shard_below
                n = shard[2]
                for lo in range(0, n, bs):
                    hi = min(lo + bs, n)
                    share = _Share(fp[below(lo):below(hi)])
                    share.nglobal = hi - lo
                    yield share
                return
            n = len(self.calculator.fingerprints_dataset)
            for lo in range(0, n, bs):
                yield fp[lo:min(lo + bs, n)]
        else:
            for batch in loader:
                yield batch

    # ------------------------------------------------------------------ CUDA-graph stepping
    def _graphs_possible(self, method, kwargs):
        calc = self.calculator
        if not (self.use_cuda_graphs and self._fast_path()):
            return False
        if type(calc) is not CalculatorTorch or calc.model.device.type != "cuda":
            return False  # the per-species calculator sizes its index sets on the host (not capturable)
        if self.optimizer_state_path is not None or method not in ("Adam", "SGD"):
            return False
        if kwargs.get("weight_decay") or kwargs.get("amsgrad") or kwargs.get("momentum") or kwargs.get("fused"):
            return False  # the state-initialising dummy step below must leave the parameters untouched
        return True

    def _prepare_graphs(self, method, kwargs):
        """Rebuild the optimizer in its capturable form, give every parameter a static gradient,
        run one eager forward/backward (library handles, autograd threads) and initialise the
        optimizer state without moving the parameters.  Returns False (eager training) on failure."""
        calc = self.calculator
        params = [p for p in calc.model.parameters() if p.requires_grad]
        dev = calc.model.device
        try:
            if method == "Adam":
                self.optimizer = torch.optim.Adam(calc.model.parameters(), capturable=True, **kwargs)
            self._weights_on_device()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for p in params:
                    p.grad = torch.zeros_like(p)
                loss = self._get_loss_batch(next(iter(self._batches(None))))
                loss.backward()
                self._sync_gradients(loss)  # also brings up the NCCL communicator before any capture
                for p in params:
                    p.grad.zero_()
                before = [p.detach().clone() for p in params]
                self.optimizer.step()  # zero gradients: creates the state, moves nothing
                for group in self.optimizer.param_groups:
                    for p in group["params"]:
                        st = self.optimizer.state.get(p, {})
                        if torch.is_tensor(st.get("step")):
                            # float64 step counter: the bias corrections 1 - beta^t are then evaluated in
                            # double precision like the eager optimizer's Python floats (a float32
                            # counter moves an fp64 trajectory by 3e-7)
                            st["step"] = torch.zeros((), dtype=torch.float64, device=dev)
                same = all(torch.equal(a, p.detach()) for a, p in zip(before, params))
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            if not same:
                raise RuntimeError("state-initialising step moved the parameters")
            self._graphs = {}
            self._graph_pool = torch.cuda.graph_pool_handle()
            self._graph_acc = torch.zeros((), dtype=torch.float64, device=dev)
            return True
        except Exception as e:  # fall back to eager stepping; nothing has been updated
            print(f"kliff_b200: CUDA-graph training disabled ({type(e).__name__}: {e})")
            self.optimizer = getattr(torch.optim, method)(calc.model.parameters(), **kwargs)
            for p in params:
                p.grad = None
            return False

    def _graph_for(self, b, batch, train):
        key = (b, train)
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with capture_without_gc(), torch.cuda.graph(g, pool=self._graph_pool):
                with torch.enable_grad():
                    if train:
                        self.optimizer.zero_grad(set_to_none=False)
                    loss = self._get_loss_batch(batch)
                    if train:
                        loss.backward()
                        total = self._sync_gradients(loss)  # gradient + loss all-reduce (captured too)
                        self.optimizer.step()
                    else:
                        total = self._sync_loss(loss.detach())
                self._graph_acc.add_(total.detach().to(torch.float64))
            self._graphs[key] = g
        return g

    def _graphed_pass(self, train):
        """One pass over the data set, one graph replay per (global) batch; the epoch loss is
        accumulated on the device and read once."""
        self._graph_acc.zero_()
        for b, batch in enumerate(self._batches(None)):
            self._graph_for(b, batch, train).replay()
        return float(self._graph_acc)

    # ------------------------------------------------------------------ fused training step
    def _prepare_fused(self, method, kwargs):
        """FusedTrainer when this run is inside the fused kernels' domain, else None (torch route)."""
        calc = self.calculator
        if not (self.use_fused and self._fast_path() and method == "Adam" and type(calc) is CalculatorTorch):
            return None
        if calc.model.device.type != "cuda" or self.optimizer_state_path is not None:
            return None
        if set(kwargs) - {"lr", "betas", "eps", "weight_decay", "amsgrad"} or kwargs.get("weight_decay") \
                or kwargs.get("amsgrad"):
            return None
        if not (calc.use_energy or calc.use_forces) or calc._packed.zeta_normalized is None:
            return None
        from .fused import FusedTrainer, mlp_structure
        if mlp_structure(calc.model) is None:
            return None
        try:
            return FusedTrainer(self, lr=kwargs.get("lr", 1e-3), betas=kwargs.get("betas", (0.9, 0.999)),
                                eps=kwargs.get("eps", 1e-8))
        except Exception as e:  # nothing has been touched yet: the torch route takes over
            print(f"kliff_b200: fused training step disabled ({type(e).__name__}: {e})")
            return None

    def _fused_batches(self):
        """This rank's (first, last + 1, global batch size) configuration run of every global batch."""
        calc = self.calculator
        bs = self.batch_size
        shard = getattr(calc, "_shard", None)
        out = []
        if shard is not None:
            n = shard[2]
            for lo in range(0, n, bs):
                hi = min(lo + bs, n)
                out.append((calc._shard_below(lo), calc._shard_below(hi), hi - lo))
            return out
        n = len(calc.fingerprints_dataset)
        dist = _dist()
        r, w = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        for lo in range(0, n, bs):
            hi = min(lo + bs, n)
            m = hi - lo
            out.append((lo + (m * r) // w, lo + (m * (r + 1)) // w, m))
        return out

    def _get_loss_epoch(self, loader):
        if self._fused is not None:
            return self._fused.run_pass(self._fused_batches(), train=False, use_graph=self.use_cuda_graphs)
        if self._graph_acc is not None and self._graphs_live():
            return self._graphed_pass(train=False)
        total = 0.0
        with torch.enable_grad():
            for batch in self._batches(loader):
                loss = self._get_loss_batch(batch)
                total += float(self._sync_loss(loss.detach()))
        return total

    def _graphs_live(self):
        return self._graph_pool is not None and self.use_cuda_graphs

    # ------------------------------------------------------------------ data-parallel plumbing
    def _sync_gradients(self, loss):
        dist = _dist()
        if dist is None:
            return loss
        params = [p for p in self.calculator.model.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in params] + [loss.detach().reshape(1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        pos = 0
        for p in params:
            p.grad.copy_(flat[pos:pos + p.numel()].view_as(p.grad))
            pos += p.numel()
        return flat[pos]

    def _sync_loss(self, loss):
        dist = _dist()
        if dist is None:
            return loss
        v = loss.detach().clone().reshape(1)
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return v[0]

    def _my_share(self, batch):
        dist = _dist()
        if dist is None:
            return batch
        r, w = dist.get_rank(), dist.get_world_size()
        n = len(batch)
        lo, hi = (n * r) // w, (n * (r + 1)) // w
        return batch[lo:hi]

    # ------------------------------------------------------------------ batch loss (loss.py:843-877)
    def _fast_path(self):
        calc = self.calculator
        return (self.vectorized and self._default_residual and getattr(calc, "_packed", None) is not None
                and getattr(calc, "_ref", None) is not None and calc.use_energy and not calc.use_stress
                and self._scalar_weights())

    def _scalar_weights(self):
        """The vectorised / fused routes carry one energy and one forces weight per configuration; data-dependent
        weights with a value per force component (MagnitudeInverseWeight) take the per-configuration route."""
        fp = self.calculator.fingerprints_dataset.fp
        key = (id(fp), len(fp))
        if getattr(self, "_scalar_key", None) != key:
            self._scalar_ok = all(np.ndim(getattr(s["configuration"].weight, name)) == 0 for s in fp
                                  for name in ("config_weight", "energy_weight", "forces_weight"))
            self._scalar_key = key
        return self._scalar_ok

    def _get_loss_batch(self, batch, normalize=True):
        if isinstance(batch, _Share):   # sharded fingerprints: already this rank's members
            nglobal = batch.nglobal
        else:
            nglobal = len(batch)
            batch = self._my_share(batch)
        if len(batch) == 0:
            params = list(self.calculator.model.parameters())
            loss = sum((p * 0.0).sum() for p in params)
        elif self._fast_path() and self.calculator._packed_usable(batch):
            loss = self._loss_vectorized(batch)
        else:
            results = self.calculator.compute(batch)
            energy_batch = results["energy"]
            forces_batch = results["forces"] if results["forces"] is not None else [None] * len(batch)
            stress_batch = results["stress"] if results["stress"] is not None else [None] * len(batch)
            losses = [self._get_loss_single_config(s, e, f, st)
                      for s, e, f, st in zip(batch, energy_batch, forces_batch, stress_batch)]
            loss = torch.stack(losses).sum()
        if normalize:
            loss = loss / nglobal
        return loss

    def _weights_on_device(self):
        """Per-configuration loss weights and atom counts as device tensors, built once per data set
        (the reference reads them per configuration per step, loss.py:904-912)."""
        calc = self.calculator
        fp = calc.fingerprints_dataset.fp
        key = (id(fp), len(fp))
        if getattr(self, "_wcache_key", None) != key:
            dev, dt = calc.model.device, calc.model.dtype
            w = [s["configuration"].weight for s in fp]
            P = calc._packed
            nat = np.diff(P.cfg_ptr).astype(np.float64)
            wc = np.array([x.config_weight for x in w], np.float64)
            we = np.array([x.energy_weight for x in w], np.float64)
            wf = np.array([x.forces_weight if x.forces_weight is not None else 0.0 for x in w], np.float64)
            denom = nat if self.residual_data["normalize_by_natoms"] else np.ones_like(nat)
            self._wcache = dict(
                e=torch.as_tensor(wc * we / denom, device=dev).to(dt),
                f=torch.as_tensor(wc * wf / denom, device=dev).to(dt),
                n3=torch.as_tensor(3 * np.diff(P.cfg_ptr), device=dev))
            self._wcache_key = key
        return self._wcache

    def _loss_vectorized(self, batch):
        """Same arithmetic as `energy_forces_residual` + sum of squares, over the whole batch."""
        calc = self.calculator
        idx = [s["index"] for s in batch]
        energy, forces, natoms = calc.compute_packed(idx)
        dev = energy.device
        P, ref, W = calc._packed, calc._ref, self._weights_on_device()
        run = _is_run(idx)
        if run:
            sl = slice(idx[0], idx[-1] + 1)
            res_e = W["e"][sl] * (energy - ref["energy"][sl])
        else:
            idx_t = torch.as_tensor(idx, device=dev)
            res_e = W["e"].index_select(0, idx_t) * (energy - ref["energy"].index_select(0, idx_t))
        loss = (res_e * res_e).sum()
        if calc.use_forces:
            if run:
                fref = ref["forces"][3 * int(P.cfg_ptr[idx[0]]):3 * int(P.cfg_ptr[idx[-1] + 1])]
                per_comp = torch.repeat_interleave(W["f"][sl], W["n3"][sl], output_size=forces.shape[0])
            else:
                fref = torch.cat([ref["forces"][3 * int(P.cfg_ptr[i]):3 * int(P.cfg_ptr[i + 1])]
                                  for i in idx])
                per_comp = torch.repeat_interleave(W["f"].index_select(0, idx_t),
                                                   W["n3"].index_select(0, idx_t),
                                                   output_size=forces.shape[0])
            res_f = per_comp * (forces - fref)
            loss = loss + (res_f * res_f).sum()
        return loss

    def _get_loss_single_config(self, sample, pred_energy, pred_forces, pred_stress):
        """loss.py:879-920."""
        device = self.calculator.model.device
        calc = self.calculator
        if calc.use_energy:
            pred = pred_energy.reshape(-1)
            ref = _t(sample["energy"]).reshape(-1).to(device)
        if calc.use_forces:
            ref_forces = _t(sample["forces"]).to(device)
            if calc.use_energy:
                pred = torch.cat((pred, pred_forces.reshape(-1)))
                ref = torch.cat((ref, ref_forces.reshape(-1)))
            else:
                pred, ref = pred_forces.reshape(-1), ref_forces.reshape(-1)
        if calc.use_stress:
            ref_stress = _t(sample["stress"]).to(device)
            pred = torch.cat((pred, pred_stress.reshape(-1)))
            ref = torch.cat((ref, ref_stress.reshape(-1)))
        conf = sample["configuration"]
        residual = self.residual_fn(conf.identifier, conf.get_num_atoms(), conf.weight, pred, ref,
                                    self.residual_data)
        return torch.sum(torch.pow(residual, 2))

    # ------------------------------------------------------------------ optimizer state (loss.py:922-935)
    def save_optimizer_state(self, path="optimizer_state.pkl"):
        if self._fused is not None:  # flat Adam moments of the fused step (kliff_b200 format)
            torch.save({"kliff_b200_fused_adam": self._fused.state_dict()}, path)
            return
        torch.save(self.optimizer.state_dict(), path)

    def load_optimizer_state(self, path="optimizer_state.pkl"):
        self.optimizer_state_path = path

    def _load_optimizer_stat(self, path):
        self.optimizer.load_state_dict(torch.load(path))


def _t(x):
    return x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))


def _is_run(idx):
    return all(idx[k + 1] == idx[k] + 1 for k in range(len(idx) - 1))
