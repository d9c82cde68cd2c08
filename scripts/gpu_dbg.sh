set -x
cd $GRAFT_REPO_ROOT
export B200ANI_DEBUG_SEGV=1
timeout 300 python - <<'PY' 2>&1 | tail -60
import faulthandler; faulthandler.enable()
import numpy as np, pyani_b200
from pyani_b200 import synth
gs = synth.make_set(2, 2, 200_000, seed=7)[:3]
eng = pyani_b200.Engine(0)
print("created", flush=True)
gids = [eng.add_genome(g.residues, g.contig_off, g.name) for g in gs]
print("staged", gids, flush=True)
res = eng.anim(np.array([[0,1],[1,0],[0,2]], np.int32))
print("anim", res, flush=True)
print(eng.anib(np.array([[0,1]], np.int32)), flush=True)
print(eng.tetra_zscores([0])[0][0,:4], flush=True)
PY
