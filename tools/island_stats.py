#!/usr/bin/env python
"""Island-size and colour-size distribution of the 1M mixed bench scene after the settle steps (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes_np
from playrho_b200 import world as W
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sc = scenes_np.mixed_fast(nb, seed=12345)
n = len(sc["bodies"])
gw = W.World(sc, device=0, body_capacity=n + 16, proxy_capacity=n + 8192, contact_capacity=6 * n, record_solve_order=True)
conf = W.default_step_conf()
for _ in range(303):
    gw.step(conf)
ids, cols = gw.solve_order()
print("constraints", len(ids), "colour sizes", np.bincount(cols).tolist())
cs, ms = gw.contacts()
rows = gw.bodies()
dyn = (rows["inv_mass"] > 0)
touch = (cs["flags"] & 1) != 0
a, b = cs["body_a"][touch], cs["body_b"][touch]
both = dyn[a] & dyn[b]
g = coo_matrix((np.ones(both.sum()), (a[both], b[both])), shape=(n, n))
k, lab = connected_components(g, directed=False)
size = np.bincount(lab)
# constraints per island: label of the dynamic body of each touching contact
lab_c = np.where(dyn[a], lab[a], lab[b])
cper = np.bincount(lab_c, minlength=k)
isl = np.nonzero(dyn[np.argsort(lab, kind="stable")[np.cumsum(size) - 1]])[0] if False else None
print("touching contacts", int(touch.sum()), "dynamic bodies", int(dyn.sum()), "components", k)
for lim in (1, 2, 4, 8, 16, 64, 256, 1024, 4096, 16384, 1 << 30):
    m = size <= lim
    print("islands with <= %d bodies: %d islands, %d bodies, %d constraints (%.1f%% of constraints)" % (lim, int(m.sum()), int(size[m].sum()), int(cper[m].sum()), 100.0 * cper[m].sum() / max(1, cper.sum())))
print("largest islands (bodies):", np.sort(size)[-10:].tolist())
