"""`cfd motion state` script (cfd/motion.xml:59-63) written for this repository with the behaviour of
the reference's resources/Scripts/cfd/Motions/State.py:30-57: per moving set, remember the motion
(motion_r, motion_a) applied last time and hand it to UnTransform as motion_r_in / motion_a_in.
On a set's first call there is no previous motion and the new one is handed over instead."""
import numpy as np
import aquagpusph as aqua

_last = {}


def main():
    iset = int(aqua.get("motion_iset"))
    now = (np.array(aqua.get("motion_r")), np.array(aqua.get("motion_a")))
    prev = _last.get(iset, now)
    aqua.set("motion_r_in", prev[0])
    aqua.set("motion_a_in", prev[1])
    _last[iset] = now
    return True
