"""GPU: savers and the restart checkpoint (reference FileManager.cpp:146-164,
Particles.cpp:82-120 + 243-323, ASCII.cpp:240-333, State.cpp:195-226 + 1517-1908).

aqh_save starts every set's <Save> file (device-side un-sort, download on the side stream, a writer
thread per saver) and rewrites AQUAgpusph.save.N.xml; a Simulation made from that file alone resumes
the run.  The resumed run starts from the particles in their ORIGINAL order (the files are
un-sorted), the continuous one from the order the previous link-list left: same cells, another order
inside them, so the two agree to the fp32 rounding of the pair sums, like the other pipeline tests."""
import os
import re

import numpy as np
import pytest

from aquagpusph_b200 import cases, casegen, host

pytestmark = pytest.mark.gpu
FIELDS = "r, normal, tangent, u, dudt, rho, drhodt, m, imove"


def _with_saves(folder, fmt="FastASCII"):
    def transform(txt):
        k = [0]

        def add(m):
            k[0] += 1
            return ('        <Save format="%s" file="%s" fields="%s" />\n    </ParticlesSet>'
                    % (fmt, os.path.join(folder, "set%d" % (k[0] - 1)), FIELDS))
        return re.sub(r"    </ParticlesSet>", add, txt)
    return transform


def _read(path):
    rows = []
    for line in open(path):
        line = line.split("#")[0].strip()
        if line:
            rows.append([float(x) for x in re.split(r"[ ,;]+", line)])
    return np.array(rows, np.float64)


def test_restart_from_the_checkpoint_continues_the_run(tmp_path, monkeypatch):
    host.set_log_level(3)
    c = cases.spheric2_dam_break(6000, 3.0, seed=3)
    nset = (c["N"] - 8, 8)
    ov = {"iter_midpoint_max": 3}
    out = str(tmp_path / "out")
    os.makedirs(out)
    monkeypatch.chdir(tmp_path)
    # the continuous run
    A = casegen.load("spheric2_dambreak_3d", c, nset, ov, device=0)
    A.step(6)
    want = {k: A.download(k, np.float32, unsorted=True) for k in ("r", "u", "rho", "dudt")}
    t_a, dt_a = float(A.scalar("t")), float(A.scalar("dt"))
    A.close()
    # three steps, a save, and the process "ends"
    B = casegen.load("spheric2_dambreak_3d", c, nset, ov, device=0, transform=_with_saves(out))
    B.step(3)
    state = {k: B.download(k, np.int32 if k == "imove" else np.float32, unsorted=True)
             for k in [f.strip() for f in FIELDS.split(",")]}
    xml = B.save()
    B.step(1)            # the files are written while the run goes on ...
    B.wait_savers()      # ... and are complete here
    files = B.saver_files()
    assert [os.path.basename(f) for f in files] == ["set0.00000.dat", "set1.00000.dat"]
    assert os.path.basename(xml) == "AQUAgpusph.save.0.xml" and os.path.exists(xml)
    # the files hold the state at the time of the call (not one step later), row i = particle i of
    # the set in its original order, floats round-trip exactly
    first = 0
    for f, n in zip(files, nset):
        got = _read(f)
        assert got.shape[0] == n
        col = 0
        for k in [x.strip() for x in FIELDS.split(",")]:
            a = state[k][first:first + n]
            w = 1 if a.ndim == 1 else a.shape[1]
            b = got[:, col:col + w].astype(a.dtype).reshape(a.shape)
            assert np.array_equal(a, b), (f, k)
            col += w
        assert col == got.shape[1]
        first += n
    # a second save takes the next index and the same state file
    assert B.save(wait=True) == xml
    assert [os.path.basename(f) for f in B.saver_files()] == ["set0.00001.dat", "set1.00001.dat"]
    B.close()
    # ---- resume from the FIRST checkpoint?  it was overwritten by the second save (the reference
    # keeps one state file per run too): load it and step back to the continuous run's step count
    txt = open(xml).read()
    assert 'name="iter" type="unsigned int" value="4"' in txt
    assert "set0.00001.dat" in txt and 'format="FastASCII"' in txt
    R = host.Simulation(xml, dims=3, device=0)
    assert int(R.scalar("iter", np.uint32)) == 4
    R.step(2)
    assert int(R.scalar("iter", np.uint32)) == 6
    assert abs(float(R.scalar("t")) - t_a) <= 2e-6 * t_a
    assert abs(float(R.scalar("dt")) - dt_a) <= 1e-5 * dt_a
    fl = state["imove"] == 1
    for k, tol in (("r", 1e-6), ("u", 2e-5), ("rho", 2e-6), ("dudt", 5e-4)):
        a = want[k][fl].astype(np.float64)
        b = R.download(k, np.float32, unsorted=True)[fl].astype(np.float64)
        err = np.abs(a - b).max() / np.abs(a).max()
        assert err <= tol, (k, err)
    R.close()


def test_unknown_output_format_falls_back_to_ascii(tmp_path, monkeypatch):
    """The examples ask for format="VTK" (libvtk in the reference): the fields are written as
    FastASCII and the checkpoint says so, so the run still resumes."""
    host.set_log_level(3)
    monkeypatch.chdir(tmp_path)
    c = cases.lattice(8, 2.0)
    sim = casegen.load("lattice_3d", c, (c["N"],), transform=_with_saves(str(tmp_path), "VTK"))
    sim.step(1)
    xml = sim.save(wait=True)
    assert os.path.basename(sim.saver_files()[0]) == "set0.00000.dat"
    assert 'format="FastASCII"' in open(xml).read()
    r = sim.download("r", np.float32, unsorted=True)
    sim.close()
    R = host.Simulation(xml, dims=3, device=0)
    assert np.array_equal(R.download("r", np.float32, unsorted=True), r)
    R.close()
