import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from nmfgpu_b200 import api
from nmfgpu_b200.workloads import dense_inputs
from oracle import binding as orc
def rel(a,b): return float(np.linalg.norm(a.astype(np.float64)-b)/np.linalg.norm(b))
m,n,k,it = 5000,1500,64,10
V,W0,H0 = dense_inputs(m,n,k,seed=5)
o = orc.run_nmf("mu", V, W0, H0, it)
L = api.Library(); L.set_verbosity(0); assert L.initialize()==0
r = L.compute(V,k,W0=W0,H0=H0,iterations=it)
W = r["W"].astype(np.float64); Wo = o["W"]
print("W err %.2e H err %.2e" % (rel(W,Wo), rel(r["H"],o["H"])))
norms = np.sqrt((W*W).sum(axis=0)); print("column norms: max |n-1| %.2e" % np.abs(norms-1).max())
s = (W*Wo).sum(axis=0)/(Wo*Wo).sum(axis=0); print("per-column scale: min %.6f max %.6f" % (s.min(), s.max()))
print("after removing the scale: %.2e" % rel(W/s, Wo))
D = np.abs(W-Wo); i = np.unravel_index(D.argmax(), D.shape); print("largest deviation at row %d col %d: %.3e (value %.3e)" % (i[0], i[1], D.max(), Wo[i]))
rows = np.sqrt((D*D).sum(axis=1)); top = np.argsort(rows)[-8:]; print("rows with the largest deviation:", top, rows[top])
print("share of the squared deviation in the worst 128 rows: %.3f" % (np.sort(rows**2)[-128:].sum()/ (rows**2).sum()))
L.finalize()
