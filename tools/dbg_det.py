import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import sph_b200 as S
from conftest import load_golden
g = load_golden("cube20_step200.npz")
s = S.default_settings()
pos, vel = g["pos0"].copy(), g["vel0"].copy()
if len(sys.argv) > 1 and sys.argv[1] == 'up':
    pos[:40, 1] += np.linspace(50, 4000, 40).astype(np.float32)
outs = []
for rep in range(3):
    sim = S.Sim(s, capacity=len(pos)); sim.upload(pos, vel); sim.step(1)
    outs.append(sim.download(S.ORDER_ID)); st = sim.stats(); print('grid', list(st.grid_dim), 'clamped', st.clamped); sim.close()
for k in ('density', 'force', 'pos'):
    print(k, 'rep0 vs rep1 differ rows:', int((outs[0][k].view(np.uint32) != outs[1][k].view(np.uint32)).reshape(len(pos), -1).any(1).sum()),
          'rep0 vs rep2:', int((outs[0][k].view(np.uint32) != outs[2][k].view(np.uint32)).reshape(len(pos), -1).any(1).sum()))
np.save('/tmp/dens_%s.npy' % os.environ.get('SPH_B200_MAX_CELLS', 'def'), outs[0]['density'])
np.save('/tmp/force_%s.npy' % os.environ.get('SPH_B200_MAX_CELLS', 'def'), outs[0]['force'])
