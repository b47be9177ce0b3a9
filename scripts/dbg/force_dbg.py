"""Debug aid: where does k_force differ from the chain oracle?  Pair terms (probe) vs sums."""
import sys, ctypes as C
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import pi_sph_fluid_b200 as pkg
from oracle import pyoracle

G = (0.0, -9.81)
g = np.load(ROOT / "tests/golden/drop_R0.075.npz")
R, snap = 0.075, 0
fluid, binit = g[f"fluid_{snap}"], g["boundary_init"]
prm = pkg.default_params(R)
sim = pkg.Simulation(prm); sim.upload(fluid, binit); sim.init_boundary(); sim.compute_accel(*G)
f, du, dv = sim.download()
o = pyoracle.Oracle(R=R, variant="chain")
of, ob = fluid.copy(), binit.copy()
gb = o.init_boundary(ob); gf = o.grid(len(of))
odu, odv = o.compute_accel(of, ob, gf, gb, *G)
bad = np.nonzero((du.view("u4") != odu.view("u4")) | (dv.view("u4") != odv.view("u4")))[0]
print("mismatching particles:", len(bad), "of", len(f), bad[:20])
off, flat = g[f"ff_off_{snap}"], g[f"ff_list_{snap}"]
prr = (of["p"] / (of["rho"] * of["rho"])).astype(np.float32)
for i in bad[:5]:
    nb = flat[off[i]:off[i + 1]]
    pairs = np.zeros((len(nb), 12), np.float32)
    pairs[:, 0], pairs[:, 1] = of["x"][i], of["y"][i]
    pairs[:, 2], pairs[:, 3] = of["x"][nb], of["y"][nb]
    pairs[:, 4], pairs[:, 5] = of["u"][i], of["v"][i]
    pairs[:, 6], pairs[:, 7] = of["u"][nb], of["v"][nb]
    pairs[:, 8], pairs[:, 9] = of["rho"][i], prr[i]
    pairs[:, 10], pairs[:, 11] = of["rho"][nb], prr[nb]
    for variant in (0, 1):
        t, sc = sim.probe_force_pair(pairs, variant)
        sx = np.float32(0); sy = np.float32(0)
        for k in range(len(nb)):
            sx = np.float32(sx + t[k, 0]); sy = np.float32(sy + t[k, 1])
        ax = np.float32(np.float32(G[0]) - sx); ay = np.float32(np.float32(G[1]) - sy)
        print(i, "variant", variant, "shortcuts", sc, "probe-sum a =", ax, ay, "| kernel", du[i], dv[i], "| oracle", odu[i], odv[i],
              "n_nb", len(nb))
