"""CPU oracle for the `plskern` path of Jchemo.jl (TEST INFRASTRUCTURE ONLY).

This package restates, in NumPy, the reference algorithm of
`/root/reference/src/plskern.jl` and the helpers of `src/utility.jl` that the
path uses.  It is the checker for the CUDA product in `jchemo.jl_b200/`; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it.  The product path never does.

PARITY UNPINNED: the reference ships no golden vectors or known-answer tests
for this path (`test/runtests.jl:1-2` holds two `using` lines) and Julia is not
installed in the build container, so the restatement could not be run against
the reference itself.  What pins it instead: (1) independent restatements
of three sibling algorithms of the reference that must give the same model —
NIPALS (`src/plsnipals.jl:70-96`), ROSA (`src/plsrosa.jl:32-96`) and, for one
response, SIMPLS (`src/plssimp.jl:28-88`) — (2) the algebraic invariants of a PLS fit, and
(3) extended-precision (longdouble) runs on small shapes.  See DESIGN.md.
"""
from .synth import synth_matrix, synth_weights, u01  # noqa: F401
from .plskern_ref import (  # noqa: F401
    Plsr, plskern, plskern_bang, transform, coef, predict, summary, xfit, xresid, sign_align,
)
from .nipals_ref import plsnipals  # noqa: F401
from .simpls_ref import plssimp  # noqa: F401
from .rosa_ref import plsrosa  # noqa: F401
from .gridscore_ref import gridscorelv, gridcvlv, locwlv  # noqa: F401
