"""type="installable" ("keep for compatibility"; CalcServer.cpp:375-400,
tests/ExternalTool/tool.hpp:28-40 in the reference): the host dlopen()s the library named by
`path`, takes its create_object(name, once) and runs the tool like any other.  The demo plugin of
tests/ExternalTool (ours, against this host's Tool class) is built here with g++."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from aquagpusph_b200 import cases, casegen, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "ExternalTool")


def build_plugin(dst):
    host.lib()   # (libaquahost.so exists)
    so = os.path.join(dst, "libdemotool.so")
    pkg = os.path.join(ROOT, "aquagpusph_b200")
    subprocess.check_call([shutil.which("g++") or "g++", "-std=c++17", "-O2", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "host"),
                           os.path.join(SRC, "tool.cpp"), "-o", so, "-L" + pkg, "-laquahost", "-laquacuda",
                           "-Wl,-rpath," + pkg])
    return so


def with_plugin(so):
    def transform(txt):
        txt = txt.replace("    </Variables>",
                          '        <Variable name="plugin_data" type="float*" length="N" />\n'
                          '        <Variable name="plugin_calls" type="unsigned int" value="0" />\n'
                          "    </Variables>", 1)
        return casegen.add_tool_after(txt, "corrector",
                                      '<Tool action="add" name="demo plugin" type="installable" once="false" '
                                      'path="%s" />' % so)
    return transform


def test_plugin_builds_and_the_front_end_accepts_the_tool(tmp_path):
    so = build_plugin(str(tmp_path))
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert re.search(r"\bT create_object\b", out), out
    c = cases.lattice(8, 2.0)
    txt = with_plugin(so)(casegen.instantiate("lattice_3d", c, (c["N"],)))
    p = tmp_path / "case.xml"
    p.write_text(txt)
    tools = host.Simulation(str(p), dims=3, parse_only=True).tools()
    assert ("demo plugin", "installable") in tools
    k = tools.index(("demo plugin", "installable"))
    assert tools[k - 1][0] == "corrector"


@pytest.mark.gpu
def test_installable_tool_runs_in_the_pipeline(tmp_path):
    so = build_plugin(str(tmp_path))
    host.set_log_level(3)
    c = cases.lattice(10, 2.0)
    sim = casegen.load("lattice_3d", c, (c["N"],), transform=with_plugin(so))
    data = np.arange(c["N"], dtype=np.float32) + 1.0
    sim.upload("plugin_data", data)
    sim.step(3)
    assert int(sim.scalar("plugin_calls", np.uint32)) == 3
    assert np.array_equal(sim.download("plugin_data"), data * 8.0)
    sim.close()
    # a library without the symbol / a missing library fail at load like the reference
    bad = casegen.instantiate("lattice_3d", c, (c["N"],))
    bad = with_plugin(os.path.join(ROOT, "aquagpusph_b200", "libaquacuda.so"))(bad)
    p = tmp_path / "bad.xml"
    p.write_text(bad)
    with pytest.raises(host.HostError, match="create_object"):
        host.Simulation(str(p), dims=3, device=0)
    p.write_text(with_plugin("/nonexistent/libnothing.so")(casegen.instantiate("lattice_3d", c, (c["N"],))))
    with pytest.raises(host.HostError):
        host.Simulation(str(p), dims=3, device=0)
