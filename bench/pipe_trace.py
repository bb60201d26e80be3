"""Debugging aid for the row-chunk pipelines of the streaming host paths (predict, transform, xfit): run with
JCB_PIPE_TRACE=1 to get, per chunk, the host time of every enqueue and the device completion time of the
three legs (host-to-device copy, kernel, device-to-host copy) on stderr."""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jchemo_b200 as jc
from oracle import synth
pk = importlib.import_module("jchemo_b200.plskern")
m, n, p, q, nlv = 1_000_000, 100_000, 500, 10, 25
fm = jc.plskern(synth.synth_matrix(1, n, p), synth.synth_matrix(2, n, q), nlv=nlv)
Xn = pk._out_empty((m, p)); Xn[:] = 0.5
for i in range(2):
    Pp = jc.predict(fm, Xn, nlv=range(0, nlv + 1)).pred
    print("predict done", file=sys.stderr)
for i in range(2):
    T = jc.transform(fm, Xn, nlv=nlv)
    print("transform done", file=sys.stderr)
for i in range(2):
    F = jc.xfit(fm, Xn, nlv=nlv)
    print("xfit done", file=sys.stderr)
