"""Probe: relative difference of the Pose2MobileArm linearization (CUDA vs oracle) for the problem set of seed 63,
without any optional factor -- is the 3e-9 on Hoff a property of Pose2::LogmapDerivative's conditioning?"""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gpmp2_b200 as G
from gpmp2_b200 import synth
from oracle import oracle as O
O.build()
import test_gpu_parity as t
model, sdf, st, pr = t._mobile_setup(48, 63)
a = t._args(pr)
got = G.batch_linearize(model, sdf, *a, st)
ref = O.linearize(model, sdf, *a, st, want_dense=False)
for k in ("Hdiag", "Hoff", "g"):
    d = np.abs(got[k] - ref[k]).reshape(48, -1).max(axis=1) / np.abs(ref[k]).max()
    print(k, "max rel %.3e" % d.max(), "problem", int(d.argmax()))
p = int((np.abs(got["Hoff"] - ref["Hoff"]).reshape(48, -1).max(axis=1)).argmax())
th = pr["init_traj"][p].reshape(2, 11, 5)[0, :, 2]
print("heading steps of that problem:", np.diff(th))
