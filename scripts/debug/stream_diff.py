import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import configs_from_npz
from kliff_b200.legacy.descriptors.symmetry_function import SymmetryFunction
from kliff_b200.packed import PackedFingerprints
dev = torch.device("cuda:0")
desc = SymmetryFunction({"Si-Si": 5.0, "Si-C": 4.5, "C-C": 4.0}, "cos", "set51")
confs = configs_from_npz("sic_training_set.npz", limit=4) + configs_from_npz("si_varying_alat.npz", limit=3)
ref = desc.transform_packed(confs, grad=True, device=dev)
ref2 = desc.transform_packed(confs, grad=True, device=dev)
print("repeat equal zeta", torch.equal(ref.zeta, ref2.zeta), "dz", torch.equal(ref.dz, ref2.dz))
n = ref.n_centers
host = torch.empty((n, len(desc)), dtype=torch.float64).pin_memory()
p = PackedFingerprints.from_configs(len(desc), dev, confs, desc.species_code, 5.0)
p.compute(desc._params, grad=True, zeta_host=host, piece=50)
p.host_ready.synchronize()
dz_ = (p.zeta - ref.zeta).abs()
print("zeta equal", torch.equal(p.zeta, ref.zeta), float(dz_.max()), "nan", int(torch.isnan(p.zeta).sum()), int(torch.isnan(ref.zeta).sum()))
d = (p.dz - ref.dz).abs()
bad = torch.nonzero(d > 0).flatten()
print("dz equal", torch.equal(p.dz, ref.dz), "ndiff", bad.numel(), "max", float(d.max()), "nan", int(torch.isnan(p.dz).sum()), int(torch.isnan(ref.dz).sum()))
if bad.numel():
    ptr = ref.dz_ptr.cpu().numpy()
    b = bad.cpu().numpy()
    at = np.searchsorted(ptr, b, side="right") - 1
    print("atoms with diffs", np.unique(at)[:40], len(np.unique(at)))
    off = ref.offsets.cpu().numpy(); cen = ref.centers.cpu().numpy() if ref.centers is not None else np.arange(n)
    for a in np.unique(at)[:5]:
        nn = off[cen[a] + 1] - off[cen[a]]
        loc = b[at == a] - ptr[a]
        rows = loc // (3 * (nn + 1)); cols = loc % (3 * (nn + 1))
        print("atom", a, "n", nn, "rows", np.unique(rows)[:20], "cols", np.unique(cols)[:30], "vals", p.dz[b[at == a][:3]].tolist(), ref.dz[b[at == a][:3]].tolist())
