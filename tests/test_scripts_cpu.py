"""CPU (no device): the run-time script path (csrc/clc.cu) up to the cubin --
`aqc_script_check` compiles an OpenCL-dialect script for sm_100a with NVRTC behind the dialect header and
reports the argument list parsed from its signature (what clGetKernelArgInfo gave the reference,
Kernel.cpp:497-556).  Ours: tests/scripts/user/Demo.cl; the reference's own scripts where the tree is at hand."""
import glob
import os
import re

import pytest

from aquagpusph_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "scripts")
DEMO = os.path.join(ROOT, "user", "Demo.cl")
DEFS = ("-DH=0.1f", "-DDEMO_GAIN=2.f")


@pytest.mark.parametrize("dims", [2, 3])
def test_user_script_compiles_and_its_signature_is_parsed(dims):
    sig = _lib.script_check(DEMO, "reflect", dims, ROOT, DEFS)
    assert sig == ("int* imove; vec* r (out); vec* u (out); float* speed (out); float* rho (out); usize N; "
                   "float dt; vec plane_n; ")
    # LINKLIST_LOCAL_PARAMS expands to the three neighbour-list arguments (types.h:106-112)
    sig = _lib.script_check(DEMO, "cell_count", dims, ROOT, DEFS)
    assert sig == "unsigned int* count (out); usize N; usize* icell; usize* ihoc; svec4 n_cells; "


def test_the_cubin_holds_the_kernels(tmp_path):
    """(a dialect macro once emptied CUDA's __global__ attribute: everything compiled, no kernel was emitted)"""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("no cuobjdump")
    out = str(tmp_path / "demo.cubin")
    os.environ["AQC_SCRIPT_DUMP"] = out
    try:
        _lib.script_check(DEMO, "reflect", 3, ROOT, DEFS)
    finally:
        del os.environ["AQC_SCRIPT_DUMP"]
    txt = subprocess.run([exe, "-res-usage", out], capture_output=True, text=True).stdout
    assert "Function reflect:" in txt and "Function cell_count:" in txt


def test_errors_are_reported_not_swallowed(tmp_path):
    with pytest.raises(_lib.AquaError, match="cannot read"):
        _lib.script_check(str(tmp_path / "nothing.cl"), "entry", 2, ROOT)
    with pytest.raises(_lib.AquaError, match="no \"__kernel void nope"):
        _lib.script_check(DEMO, "nope", 2, ROOT, DEFS)
    bad = tmp_path / "bad.cl"
    bad.write_text('#include "resources/Scripts/types/types.h"\n'
                   "__kernel void entry(__global vec* r, usize N)\n{\n    r[get_global_id(0)] = undefined_name;\n}\n")
    with pytest.raises(_lib.AquaError, match="undefined_name"):
        _lib.script_check(str(bad), "entry", 2, ROOT)
    missing = tmp_path / "inc.cl"
    missing.write_text('#include "not/there.h"\n__kernel void entry(usize N) {}\n')
    with pytest.raises(_lib.AquaError, match="not/there.h"):
        _lib.script_check(str(missing), "entry", 2, ROOT)
    # H is a <Define> of the problem: without it the script does not compile
    with pytest.raises(_lib.AquaError, match="H"):
        _lib.script_check(DEMO, "reflect", 2, ROOT)


REF = "/root/reference"
REF_DEFS = ("-DH=0.04f", "-DCONW=1.f", "-DCONF=1.f", "-DSUPPORT=2.f", "-DKERNEL_NAME=Wendland",
            "-D__LAP_MONAGHAN__=1", "-D__LAP_MORRIS__=2", "-D__LAP_FORMULATION__=__LAP_MONAGHAN__")


@pytest.mark.skipif(not os.path.isdir(REF + "/resources/Scripts"), reason="needs the reference tree (build container only)")
def test_reference_scripts_compile_behind_the_dialect_header(tmp_path):
    """The scripts the 9 uncovered examples still miss (case-local ones included) compile as they lie, in the
    dimension of their example: Inlet / Outlet, Portal, the ideal-gas family with its Riemann solver, the BI
    pressure force, ...  And so do the hot-path scripts -- which the registry serves by hand-written kernels and
    this path never touches: a check of the dialect, not a product route."""
    root = tmp_path / "root"
    (root / "resources").mkdir(parents=True)
    os.symlink(REF + "/resources/Scripts", root / "resources" / "Scripts")

    def ok(path, entry, dims):
        _lib.script_check(path, entry, dims, str(root), ("-DDIMS=%d" % dims,) + REF_DEFS)

    local = {"2D/normal_impact_wall": ["init.cl"], "2D/taylor_green": ["Rescale.cl"],
             "2D/adiabatic_expansion": ["spring.cl"], "2D/cylinder_inside_channel": ["Initialization.cl"],
             "2D/shock_1d": ["bc.cl"], "2D/shock_point": ["bc.cl"], "3D/spheric_testcase2_dambreak": ["h_sensor.cl"]}
    n = 0
    for ex, files in local.items():
        for fn in files:
            path = "%s/examples/%s/src/templates/%s" % (REF, ex, fn)
            for entry in re.findall(r"__kernel\s+void\s+(\w+)", open(path).read()):
                ok(path, entry, int(ex[0]))
                n += 1
    assert n >= 14
    S = REF + "/resources/Scripts/"
    for script, entries, dims in (
            ("cfd/Boundary/Inlet/Inlet.cl", ("feed", "rates"), 2), ("cfd/Boundary/Outlet/Outlet.cl", ("feed", "rates"), 2),
            ("cfd/Boundary/Portal/Mirror.cl", ("mirror", "teleport", "unmirror"), 2),
            ("cfd/Boundary/Portal/Interactions.cl", ("entry",), 2), ("cfd/Boundary/Portal/Shepard.cl", ("entry",), 2),
            ("cfd/Forces/BI/PressureForces.cl", ("entry",), 2),
            ("cfd/ideal_gas/EOS.cl", ("entry",), 2), ("cfd/ideal_gas/Sort.cl", ("entry",), 2),
            ("cfd/ideal_gas/TimeStep.cl", ("entry",), 2), ("cfd/ideal_gas/Rates.cl", ("entry",), 2),
            ("cfd/ideal_gas/riemann/Interactions.cl", ("entry",), 2), ("cfd/ideal_gas/riemann/Rates.cl", ("entry",), 2),
            ("cfd/ideal_gas/time_scheme/midpoint.cl", ("predictor", "midpoint", "relax", "corrector"), 2),
            ("cfd/ideal_gas/symmetry/Mirror.cl", ("set",), 2),
            # hot-path scripts (dialect check only)
            ("cfd/Interactions.cl", ("entry",), 3), ("basic/MLS.cl", ("entry", "mls_inv"), 3),
            ("cfd/Boundary/BIe/ElasticBounce.cl", ("entry",), 3), ("cfd/MPI.cl", ("interactions", "gamma"), 3),
            ("cfd/Boundary/BI/Shepard.cl", ("compute",), 2), ("basic/time_scheme/adam_bashforth.cl", ("corrector",), 3)):
        for e in entries:
            ok(S + script, e, dims)
    # under <Define name="__LAP_FORMULATION__" value="__LAP_MORRIS__"/> (examples/2D/cylinder_inside_channel,
    # taylor_green) the host sends the scripts whose hand-written kernel holds the Monaghan branch only to this
    # path (Kernel::setup): they compile with that definition
    morris = tuple(d for d in REF_DEFS if not d.startswith("-D__LAP_FORMULATION__")) + ("-D__LAP_FORMULATION__=__LAP_MORRIS__",)
    for script, e, dims in (("cfd/Boundary/BI/LapU.cl", "freeslip", 2), ("cfd/Boundary/BI/NoSlip.cl", "entry", 2),
                            ("cfd/MPI.cl", "interactions", 3), ("cfd/Boundary/Portal/Interactions.cl", "entry", 2)):
        _lib.script_check(S + script, e, dims, str(root), ("-DDIMS=%d" % dims,) + morris)
    # the argument list equals what the reference's compiler reports (tests/golden/kernel_signatures.json)
    import json
    gold = json.load(open(os.path.join(HERE, "golden", "kernel_signatures.json")))
    sig = _lib.script_check(S + "cfd/Boundary/Inlet/Inlet.cl", "feed", 2, str(root), ("-DDIMS=2",) + REF_DEFS)
    got = [(a.split()[-1] if "(out)" not in a else a.split()[-2], "*" in a, "(out)" not in a and "*" in a)
           for a in sig.split("; ") if a.strip()]
    want = [(n_, p_, c_ and p_) for n_, p_, c_ in gold["cfd/Boundary/Inlet/Inlet.cl"]["feed"]]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert [g[1] for g in got] == [w[1] for w in want]
    assert [g[2] for g in got] == [w[2] for w in want]


def test_the_blast_rim_script_compiles():
    """tests/scripts/user/BlastRim.cl (this repository's wording of examples/2D/shock_point's bc.cl, used by
    tests/test_gpu_presets.py::test_shock_point_blast_pipeline): both kernels compile for sm_100a behind the
    dialect header with the problem's H and SUPPORT, and bind imove / r / N / R by name."""
    rim = os.path.join(ROOT, "user", "BlastRim.cl")
    defs = ("-DH=0.025f", "-DSUPPORT=2.f")
    assert _lib.script_check(rim, "set_fixed", 2, ROOT, defs) == "int* imove (out); vec* r; usize N; float R; "
    assert _lib.script_check(rim, "unset_fixed", 2, ROOT, defs) == "int* imove (out); usize N; "
