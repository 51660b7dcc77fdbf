"""GPU parity of the open-boundary kernels written by hand after round 2's GPU budget was spent
(cfd/Boundary/Inlet/Inlet.cl::feed / rates, Outlet/Outlet.cl::rates / feed, Portal/Mirror.cl::mirror / unmirror /
teleport; presets cfd/inlet.xml, cfd/outlet.xml, cfd/portal.xml) through the Kernel-tool C-ABI,
against the oracle AND against the committed outputs of the reference's own scripts
(tests/golden/open_boundary_outputs.npz, made by make_golden_open_boundary.py where the reference tree is).

In a file of its own that sorts after every other GPU suite: its first run on a B200 is the driver's, and nothing
may hide behind it under -x.  The CPU halves are tests/test_oracle_vs_reference.py (restatement == scripts),
tests/test_oracle_golden.py (restatement == committed outputs) and tests/test_presets_host_emulation.py (the
CUDA kernel bodies and their launchers' arithmetic, compiled for the host == restatement)."""
import os

import numpy as np
import pytest

import open_boundary_common as ob
from aquagpusph_b200 import _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "open_boundary_outputs.npz")


@pytest.mark.parametrize("dims", [2, 3])
def test_open_boundary_kernels(oracle, dims):
    """Element-wise, no contraction (-fmad=false), IEEE + - * / only: bit-exact after every kernel."""
    import oracle.oracle as O
    case, v = ob.state(dims)
    D = O.make_defs(dims, case["h"])
    G = np.load(GOLD)
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    o = ob.args_of(v)
    d = {k: (ctx.array(x) if isinstance(x, np.ndarray) and k not in ob.SCALARS else x) for k, x in v.items()}
    d["refd"] = ctx.array(v["refd"])
    for step, key in enumerate(ob.STEPS):
        # (the Kernel tool's global size is N for every one of them; the inlet launcher trims it itself)
        ctx.launch(key[0], key[1], d)
        ob.oracle_step(oracle, D, dims, key, o)
        for k in ob.WRITES[key]:
            got = d[k].get()
            assert got.tobytes() == o[k].tobytes(), (key, k)
            assert got.tobytes() == G["%dD_step%d_%s" % (dims, step, k)].tobytes(), (key, k)
    final = {k: d[k].get() for k in o if isinstance(o[k], np.ndarray) and k not in ob.SCALARS}
    final["n_cells"], final["inlet_N"] = v["n_cells"], v["inlet_N"]
    final["N"], final["nbuffer"] = v["N"], v["nbuffer"]
    ob.checks(v, final, dims)
    # a starving flag of 0: no launch, nothing touched
    before = ctx.launch_count()
    d2 = dict(d)
    d2["inlet_starving"] = 0
    ctx.launch(ob.INLET, "feed", d2)
    assert ctx.launch_count() == before
    ctx.close()
