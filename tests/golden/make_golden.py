#!/usr/bin/env python3
"""Regenerates the fixtures in this directory from the reference checkout (run in the build container, where
/root/reference is mounted; the GPU box only sees the committed fixtures).

  HD155358.vels, TEST_2-1_COMPACT.vels   measurement data files of the reference (time[day] rv[m/s] err[m/s]), verbatim
  rvcurve_ben_2-1.txt, rvcurve_ben_3-1.txt   line 4 of plotArchive/Ben's 2-1/log_Ben-2-1 and Ben's 3-1/log_Ben-3-1:
                                             1000 times then 1000 REBOUND radial velocities logged by
                                             mcmc_benchmark_smala.py:49 (State.get_rv_plotting) -- KAT-3 / KAT-4

The scalar golden values (KAT-1 initial conditions, KAT-2 logp = -2.41616612321, KAT-5 Encounter vectors, KAT-6) are
printed notebook outputs; they are transcribed, with their notebook line numbers, in tests/rvtest.py and
tests/test_oracle_kat.py.
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    shutil.copyfile(os.path.join(REF, "HD155358.vels"), os.path.join(HERE, "HD155358.vels"))
    shutil.copyfile(os.path.join(REF, "plotArchive", "Ben's 2-1", "TEST_2-1_COMPACT.vels"),
                    os.path.join(HERE, "TEST_2-1_COMPACT.vels"))
    for sub, log, out in (("Ben's 2-1", "log_Ben-2-1", "rvcurve_ben_2-1.txt"), ("Ben's 3-1", "log_Ben-3-1", "rvcurve_ben_3-1.txt")):
        with open(os.path.join(REF, "plotArchive", sub, log)) as f:
            line4 = f.read().split("\n")[3]
        vals = line4.split()
        assert len(vals) == 2000, len(vals)
        with open(os.path.join(HERE, out), "w") as f:
            f.write(" ".join(vals) + "\n")
    print("fixtures regenerated from", REF)


if __name__ == "__main__":
    main()
