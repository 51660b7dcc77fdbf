"""Flattens a reference example (Main.xml + the presets it includes) into ONE
resolved XML template, using this repo's own XML front-end
(AQUAgpusph-b200 --resolve = State::parse + State::write).

Build-container only (reads /root/reference); the output is committed under
aquagpusph_b200/cases_xml/ so that the GPU box, which has no reference tree,
can run the unchanged pipeline.  The template keeps the example's {{KEY}}
placeholders; aquagpusph_b200.casegen fills them at run time.

    python tools/resolve_case.py
"""
import os
import re
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EXE = os.path.join(ROOT, "aquagpusph_b200", "AQUAgpusph-b200")
OUT = os.path.join(ROOT, "aquagpusph_b200", "cases_xml")

CASES = {
    # name: (example dir, dims[, main file])
    "spheric2_dambreak_3d": ("examples/3D/spheric_testcase2_dambreak/src/templates", 3),
    "spheric5_dambreak_2d": ("examples/2D/spheric_testcase5_dambreak/src/templates", 2),
    "spheric2_dambreak_mpi_3d": ("examples/3D/spheric_testcase2_dambreak_mpi/src/templates", 3),
    # BASELINE config 4: tuned liquid damper, BIe + forces + energy + motion presets (the two
    # `python` tools of cfd/motion.xml stay in the template; casegen.prescribed_roll replaces them)
    "spheric9_tld_2d": ("examples/2D/spheric_testcase9_tld/src/templates", 2),
    # lid-driven cavity (SPHERIC test 3): improved Euler, delta-SPH full, BI boundaries + BINoSlip
    "spheric3_liddriven_2d": ("examples/2D/spheric_testcase3_liddriven/src/templates", 2),
    # standing wave of Souto-Iglesias et al. 2012: improved Euler, delta-SPH full, BI bottom, two symmetry
    # planes (cfd/symmetry.xml twice: Symmetry/Mirror.cl) feeding on buffer particles, kinetic-energy report
    "souto2012_standingwave_2d": ("examples/2D/souto_etal_2012_standingwave/src/templates", 2),
    # circular blast of an ideal gas (examples/2D/shock_point): midpoint scheme with autostop / autorelax, the
    # ideal-gas presets (EOS, energy rates, energy time scheme) and a case-local bc.cl that freezes the rim
    "shock_point_2d": ("examples/2D/shock_point/src/templates", 2),
    # the reference's own multi-device parity test (tests/2D/MPI_plane)
    "mpi_plane_2d_serial": ("tests/2D/MPI_plane/cMake", 2, "main_serial.xml"),
    "mpi_plane_2d_mpi": ("tests/2D/MPI_plane/cMake", 2, "main_mpi.xml"),
    # BASELINE config 5: our own Main.xml (cases_xml/src/lattice_3d) over the reference's presets
    "lattice_3d": ("repo:aquagpusph_b200/cases_xml/src/lattice_3d", 3, "Lattice.xml"),
    "lattice_mpi_3d": ("repo:aquagpusph_b200/cases_xml/src/lattice_mpi_3d", 3, "Lattice.xml"),
    "lattice_ab_3d": ("repo:aquagpusph_b200/cases_xml/src/lattice_ab_3d", 3, "Lattice.xml"),
}


def installed_root(tmp):
    """resources/ laid out as `make install` does (Presets/src -> Presets)."""
    r = os.path.join(tmp, "root", "resources")
    os.makedirs(r)
    os.symlink(os.path.join(REF, "resources/Presets/src"), os.path.join(r, "Presets"))
    os.symlink(os.path.join(REF, "resources/Scripts"), os.path.join(r, "Scripts"))
    return os.path.join(tmp, "root")


def resolve(name, src, dims, main="Main.xml"):
    with tempfile.TemporaryDirectory() as tmp:
        root = installed_root(tmp)
        case = os.path.join(tmp, "case")
        shutil.copytree(os.path.join(ROOT, src[5:]) if src.startswith("repo:") else os.path.join(REF, src), case)
        keys = {}
        for fn in os.listdir(case):
            if not fn.endswith(".xml"):
                continue
            p = os.path.join(case, fn)
            txt = open(p).read()
            # the shipped 3-D Main.xml includes a root_path.xml that does not exist
            txt = "\n".join(l for l in txt.split("\n") if "root_path.xml" not in l)
            # CTest placeholders of the reference's tests (cMake/Test.cmake:3-19)
            txt = txt.replace("@RESOURCES_DIR@", os.path.join(root, "resources"))
            for k in set(re.findall(r"\{\{(\w+)\}\}", txt)):
                if k not in keys:
                    keys[k] = "9%06d" % (len(keys) + 1) if k in ("N", "N_SENSORS", "NBC", "NFLUID") \
                        else "0.9%05d1" % (len(keys) + 1)
                txt = txt.replace("{{%s}}" % k, keys[k])
            open(p, "w").write(txt)
        out = os.path.join(OUT, name + ".xml")
        subprocess.check_call([EXE, "-i", main, "-d", str(dims), "-l", "2", "--root", root,
                               "--resolve", out], cwd=case)
        txt = open(out).read()
        txt = txt.replace(root, "")
        for k, v in keys.items():
            txt = txt.replace(v, "{{%s}}" % k)
        hdr = ("<!-- Resolved by tools/resolve_case.py from %s of AQUAgpusph 5.0.4 with the\n"
               "     presets it includes; generated file, placeholders filled by casegen.py -->\n" % src)
        txt = txt.replace("<sphInput>\n", hdr + "<sphInput>\n", 1)
        open(out, "w").write(txt)
        n = txt.count("<Tool ")
        print("%s: %d tools, placeholders %s" % (name, n, sorted(keys)))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, spec in CASES.items():
        resolve(name, *spec)
