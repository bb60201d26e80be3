"""Small end-to-end run for compute-sanitizer (config C1 shapes + a weighted/scaled fit + sweep + gridscore)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jchemo_b200 as jc
from oracle import synth
X = synth.synth_matrix(1, 150, 200); Y = synth.synth_matrix(2, 150, 2); Xn = synth.synth_matrix(4, 50, 200)
fm = jc.plskern(X, Y, nlv=5)
jc.predict(fm, Xn, nlv=range(0, 6)); jc.transform(fm, Xn); jc.coef(fm, nlv=3)
w = synth.synth_weights(150, uniform=False)
fm = jc.plskern_bang(X.copy(order="F"), Y.copy(order="F"), w, nlv=4, scal=True)
X2 = synth.synth_matrix(1, 3001, 70); Y2 = synth.synth_matrix(2, 3001, 1)
fm2 = jc.plskern(X2, Y2, nlv=6)
jc.gridscorelv(X2[:2000], Y2[:2000], X2[2000:], Y2[2000:], score="rmsep", nlv=range(0, 7))
print("sanitize_small done")
