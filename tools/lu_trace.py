#!/usr/bin/env python
"""Reads the box trace of a -DLT_TRACE build (HYDRO_LT_TRACE=<prefix>): when each box of k_lu_tiled was claimed, started, finished."""
import sys
import numpy as np
L = open(sys.argv[1]).read().split("\n")
nb, nbi = map(int, L[0].split())
a = np.array([list(map(int, l.split())) for l in L[1:] if l.strip()], dtype=np.int64)
nbj = nb // nbi
t0 = a[:, 1].min()
claim = (a[:, 1] - t0).reshape(nbj, nbi) / 1e3; start = (a[:, 2] - t0).reshape(nbj, nbi) / 1e3; end = (a[:, 3] - t0).reshape(nbj, nbi) / 1e3
sm = a[:, 4].reshape(nbj, nbi)
print("boxes %d (%d x %d), total %.1f us" % (nb, nbi, nbj, end.max()))
np.set_printoptions(linewidth=250, precision=0, suppress=True)
print("start (us) [J rows, I columns], first 34 rows"); print(start[:34])
print("busy = end - start (us)"); print((end - start)[:34])
print("claim (us)"); print(claim[:34])
dJ = start[1:, :] - start[:-1, :]; dI = start[:, 1:] - start[:, :-1]
print("start step along J: mean %.1f us, median %.1f; along I: mean %.1f median %.1f" % (dJ.mean(), np.median(dJ), dI.mean(), np.median(dI)))
print("busy: mean %.1f us min %.1f max %.1f" % ((end - start).mean(), (end - start).min(), (end - start).max()))
