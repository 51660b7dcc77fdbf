"""`cfd motion data` script of the tuned-liquid-damper case (cfd/motion.xml:55-58) for this
repository: a prescribed roll theta(t) = theta0 sin(2 pi t / T) around motion_r, written against the
reference's python-tool API (aquagpusph.get / aquagpusph.set, main() -> bool).

It stands in for examples/2D/spheric_testcase9_tld/src/templates/Motion.py, whose mechanical model
needs the recorded mass position T_1-94_A100mm_water.dat that the reference tree does not ship.
The arithmetic is the one casegen.prescribed_roll writes as set_scalar expressions (double
precision, narrowed once), so both routes give the same motion bit for bit."""
import math

import numpy as np
import aquagpusph as aqua


def main():
    t = aqua.get("t")
    theta0 = aqua.get("motion_theta0")
    period = aqua.get("motion_period")
    w = 2 * math.pi / period
    ph = 2 * math.pi * t / period
    a = np.zeros(4, dtype=np.float32)
    a[2] = theta0 * math.sin(ph)
    aqua.set("motion_a", a)
    dadt = np.zeros(4, dtype=np.float32)
    dadt[2] = theta0 * w * math.cos(ph)
    aqua.set("motion_dadt", dadt)
    ddaddt = np.zeros(4, dtype=np.float32)
    ddaddt[2] = 0 - theta0 * w * w * math.sin(ph)
    aqua.set("motion_ddaddt", ddaddt)
    return True
