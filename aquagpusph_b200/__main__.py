"""`python -m aquagpusph_b200` -- the command line of AQUAgpusph-b200 (aquagpusph_b200/host/main.cpp,
i.e. the reference's flags, aquagpusph/ArgumentsManager.cpp:56-175) run through the Python binding,
so that problems with type="python" tools work: the compiled CLI has no interpreter to hand their
scripts to, this process registers itself as the script runner (include/aquahost.h).

    python -m aquagpusph_b200 -i Main.xml -d 2 --root /path/with/resources [--steps N] [-l LEVEL]
    python -m aquagpusph_b200 -i Main.xml -d 3 --root ... --resolve flat.xml      (no GPU needed)
"""
import argparse
import sys

import numpy as np

from . import host


def parser():
    ap = argparse.ArgumentParser(prog="python -m aquagpusph_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("-i", "--input", default="Input.xml", help="XML definition input file")
    ap.add_argument("-l", "--log-level", default="1", help="0 debug .. 3 error")
    ap.add_argument("-q", "--queues", default=None, help="accepted for compatibility (one CUDA stream)")
    ap.add_argument("-d", "--dimensions", type=int, default=3, choices=[2, 3])
    ap.add_argument("--root", default=None, help="folder that contains resources/ (or AQUAGPUSPH_ROOT)")
    ap.add_argument("--resolve", default=None, metavar="OUT", help="write the resolved problem and exit")
    ap.add_argument("--steps", type=int, default=-1, help="run N steps instead of the <Timing> criteria")
    ap.add_argument("--device", type=int, default=-1, help="CUDA device (default: the XML's <Device>)")
    return ap


def main(argv=None):
    a = parser().parse_args(argv)
    host.set_log_level(int(str(a.log_level)[:1] or 1))
    try:
        if a.resolve:
            sim = host.Simulation(a.input, dims=a.dimensions, root=a.root, parse_only=True)
            sim.write_resolved(a.resolve)
            print("%d tools written to %s" % (len(sim.tools()), a.resolve))
            sim.close()
            return 0
        sim = host.Simulation(a.input, dims=a.dimensions, device=a.device, root=a.root)
        if a.steps >= 0:
            sim.step(a.steps)
            sim.sync()
        else:
            sim.run()
        print("Simulation finished: iter = %d, t = %g s, %d CUDA kernels launched"
              % (int(sim.scalar("iter", np.uint32)), float(sim.scalar("t")), sim.launch_count()))
        sim.close()
        return 0
    except host.HostError as e:
        sys.stderr.write("ERROR: %s\n" % e)
        return 1


if __name__ == "__main__":
    sys.exit(main())
