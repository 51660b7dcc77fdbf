"""GPU: run-time scripts (csrc/clc.cu): a user script in the reference's OpenCL dialect
(tests/scripts/user/Demo.cl, ours) is compiled by NVRTC for sm_100a at set-up, bound by argument NAME like
any kernel tool (Kernel.cpp:497-556) and launched -- through the C-ABI directly and through the host's
`kernel` tool inside a pipeline.  Arithmetic without contraction: bit-exact against numpy in fp32."""
import os

import numpy as np
import pytest

from aquagpusph_b200 import _lib, cases, casegen, host

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "scripts")
DEMO = os.path.join(ROOT, "user", "Demo.cl")


def _reflect_numpy(imove, r, u, rho, dt, n, dims):
    f = np.float32
    r, u, rho = r.copy(), u.copy(), rho.copy()
    speed = np.zeros(len(imove), f)
    mv = imove > 0
    uu = u[:, :dims]
    d = uu[:, 0] * n[0]
    for k in range(1, dims):
        d = (d + uu[:, k] * n[k]).astype(f)
    proj = (d[:, None] * n[None, :dims]).astype(f)
    new = (uu - (f(2.0) * proj).astype(f)).astype(f)
    u[mv, :dims] = new[mv]
    step = (f(dt) * u).astype(f)
    r[mv] = (r[mv] + step[mv]).astype(f)     # (the DEMO_GAIN term is exactly zero)
    s2 = u[:, 0] * u[:, 0]
    for k in range(1, dims):
        s2 = (s2 + u[:, k] * u[:, k]).astype(f)
    speed[mv] = np.sqrt(s2.astype(f))[mv]
    rho[mv] = np.maximum(np.minimum(rho[mv], f(1010.0)), f(990.0))
    return r, u, speed, rho


@pytest.mark.parametrize("dims", [2, 3])
def test_user_script_through_the_c_abi(oracle, dims):
    case = cases.dam_break(dims, 10 if dims == 3 else 40, 2.0)
    N, V = case["N"], (4 if dims == 3 else 2)
    rng = np.random.default_rng(3)
    imove = np.ascontiguousarray(case["imove"]).copy()
    r = np.ascontiguousarray(case["r"]).copy()
    u = rng.normal(size=(N, V)).astype(np.float32)
    if dims == 3:
        u[:, 3] = 0
    rho = rng.uniform(980, 1020, N).astype(np.float32)
    n = np.zeros(V, np.float32)
    n[0], n[1] = 0.6, 0.8
    ctx = _lib.Context(0, dims=dims, h=case["h"])
    kid = ctx.script_compile(DEMO, "reflect", ROOT, ("-DH=%rf" % float(case["h"]), "-DDEMO_GAIN=2.f"))
    assert ctx.script_compile(DEMO, "reflect", ROOT, ("-DH=%rf" % float(case["h"]), "-DDEMO_GAIN=2.f")) == kid
    d = dict(imove=ctx.array(imove), r=ctx.array(r), u=ctx.array(u), speed=ctx.array(np.full(N, 7, np.float32)),
             rho=ctx.array(rho), N=N, dt=1e-3, plane_n=n)
    ctx.launch(None, None, d, kid=kid)
    wr, wu, ws, wrho = _reflect_numpy(imove, r, u, rho, 1e-3, n, dims)
    assert np.array_equal(d["u"].get(), wu) and np.array_equal(d["r"].get(), wr)
    assert np.array_equal(d["speed"].get(), ws) and np.array_equal(d["rho"].get(), wrho)
    assert (ws[imove > 0] > 0).all() and (ws[imove <= 0] == 0).all()
    # a neighbour-list walk with a __local array and a barrier
    ll = oracle.linklist(case["r"], dims, 2.0, case["h"])
    kid2 = ctx.script_compile(DEMO, "cell_count", ROOT, ("-DH=%rf" % float(case["h"]),))
    icell = np.sort(ll["icell"])
    nw = int(ll["ncells"][3])
    ihoc = np.full(nw, N, np.uint32)
    first = np.flatnonzero(np.r_[True, icell[1:] != icell[:-1]])
    ihoc[icell[first]] = first
    d2 = dict(count=ctx.zeros(N, np.uint32), N=N, icell=ctx.array(icell), ihoc=ctx.array(ihoc), n_cells=ll["ncells"])
    ctx.launch(None, None, d2, kid=kid2)
    pop = np.bincount(icell, minlength=nw + 2)
    want = (pop[icell] + np.where(icell + 1 < nw, pop[np.minimum(icell + 1, nw)], 0)).astype(np.uint32)
    assert np.array_equal(d2["count"].get(), want)
    ctx.close()


def test_user_script_as_a_kernel_tool_of_a_pipeline(tmp_path):
    """<Tool type="kernel" path=".../Demo.cl" entry_point="reflect"/> in the lattice pipeline: not in the
    registry, so Kernel::setup compiles it (the problem's <Define>s included: H), binds imove, r, u, speed, rho,
    N, dt, plane_n by name and runs it every step; a script that does not compile fails the load with the
    compiler's message."""
    host.set_log_level(3)
    c = cases.lattice(10, 2.0)

    def with_tool(path):
        def transform(txt):
            txt = txt.replace("    </Variables>",
                              '        <Variable name="speed" type="float*" length="N" />\n'
                              '        <Variable name="plane_n" type="vec" value="0.0, 0.0, 1.0, 0.0" />\n'
                              "    </Variables>", 1)
            return casegen.add_tool_after(txt, "corrector",
                                          '<Tool action="add" name="user reflect" type="kernel" once="false" '
                                          'path="%s" entry_point="reflect" n="" />' % path)
        return transform
    os.environ["AQUAGPUSPH_ROOT"] = ROOT
    try:
        plain = casegen.load("lattice_3d", c, (c["N"],))
        sim = casegen.load("lattice_3d", c, (c["N"],), transform=with_tool(DEMO))
        plain.step(1)
        sim.step(1)
        u0 = plain.download("u", np.float32)
        u1 = sim.download("u", np.float32)
        # the user tool mirrored u_z after the corrector: same x, y, opposite z
        assert np.array_equal(u1[:, :2], u0[:, :2]) and np.array_equal(u1[:, 2], -u0[:, 2]) and np.abs(u0[:, 2]).max() > 0
        sp = sim.download("speed", np.float32)
        assert np.allclose(sp, np.sqrt((u1[:, :3].astype(np.float64) ** 2).sum(1)), rtol=1e-6)
        plain.close()
        sim.close()
        bad = tmp_path / "bad.cl"
        bad.write_text('#include "resources/Scripts/types/types.h"\n__kernel void reflect(__global vec* r, usize N)\n'
                       "{\n    r[get_global_id(0)] = undefined_name;\n}\n")
        with pytest.raises(host.HostError, match="undefined_name"):
            casegen.load("lattice_3d", c, (c["N"],), transform=with_tool(str(bad)))
    finally:
        del os.environ["AQUAGPUSPH_ROOT"]
