"""Regenerates the golden vectors in this directory from the CPU oracle (oracle/oracle.cpp).

The reference (dineshadepu/prestige) has no implementation of this path and no fixtures, so these
vectors pin THIS repo's written contract, not reference outputs: "parity unpinned".
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc            # noqa: E402
from prestige_b200 import synth             # noqa: E402

b = synth.wcsph_block_3d(9, 8, 7).shuffled()
r = orc.wcsph(3, b.params, b.arrays)
pr, _ = orc.pairs(3, b.arrays["x"], b.arrays["y"], b.arrays["z"], b.arrays["h"])
np.savez_compressed(os.path.join(HERE, "wcsph3d_small.npz"), x=b.arrays["x"], pairs=pr, **r)

d = synth.dem_column_3d(6).shuffled()
f1, h1, _ = orc.dem(d.params, 12, d.arrays)
f2, h2, _ = orc.dem(d.params, 12, d.arrays, hist=h1)
np.savez_compressed(os.path.join(HERE, "dem3d_small.npz"), **f2, **h2)

c = synth.wcsph_dambreak_2d(dx=0.05).shuffled()
r = orc.wcsph(2, c.params, c.arrays)
np.savez_compressed(os.path.join(HERE, "wcsph2d_small.npz"), **r)
print("wcsph / dem golden vectors written")

# coupled SPH-DEM (tags 0 fluid, 1 boundary, 2 solid): second evaluation, so the contact history is exercised
k = synth.coupled_block_3d(9, 8, 9).shuffled()
r1, h1, _ = orc.coupled(k.params, k.max_contacts, k.arrays)
r2, h2, _ = orc.coupled(k.params, k.max_contacts, k.arrays, hist=h1)
tag = k.arrays["tag"]
nb, _ = orc.pairs(3, k.arrays["x"], k.arrays["y"], k.arrays["z"], k.arrays["h"])
nb = nb[(tag[nb[:, 0]] == 0) | (tag[nb[:, 1]] == 0)]
ct, _ = orc.pairs(3, k.arrays["x"], k.arrays["y"], k.arrays["z"], k.arrays["rad"], mode=1)
ct = ct[(tag[ct[:, 0]] == 2) & (tag[ct[:, 1]] != 0)]
np.savez_compressed(os.path.join(HERE, "coupled3d_small.npz"), x=k.arrays["x"], tag=tag, neighbours=nb, contacts=ct, hist_n=h2["hist_n"], **r2)
print("coupled golden vectors written")

# dummy-particle wall pressure: the small dam break with boundary_model = 1 (p, slaved wall density, rates)
c = synth.wcsph_dambreak_2d(dx=0.05).shuffled()
r = orc.wcsph(2, dict(c.params, boundary_model=1), c.arrays)
np.savez_compressed(os.path.join(HERE, "wall2d_small.npz"), **r)
print("wall pressure golden vectors written")
