"""Observation containers -- Python-3 mirror of the reference's observations.py.

Same class names and attributes (tf, tb, rvf, rvb, errorf, errorb, Npoints, t, rv, err;
observations.py:6-16).  ``FakeObservation`` keeps the reference's numpy draw order
(observations.py:32-47) but gets its radial velocities from the GPU engine instead of rebound.
"""
import numpy as np

from . import _abi


class Observation(object):
    tf = None
    tb = None
    rvf = None
    rvb = None
    Npoints = 0
    errorf = None
    errorb = None
    t = None
    rv = None
    err = None

    # --- GPU residency (not in the reference): one rv_obs per (context) ---
    def _handle(self, ctx):
        cache = self.__dict__.setdefault("_rv_handles", {})
        key = (id(ctx), float(self.Npoints), len(self.tf), len(self.tb))
        h = cache.get(key)
        if h is None:
            h = _abi.ObsHandle(ctx, self.tf, self.rvf, self.errorf, self.tb, self.rvb, self.errorb, self.Npoints)
            cache.clear()
            cache[key] = h
        return h


class FakeObservation(Observation):
    def __init__(self, state, Npoints=30, error=0., errorVar=0., tmax=1.5):
        """Generates fake observations (observations.py:18-50).

        One simulation without an encounter distance visits obs.tf in order and then obs.tb in order
        (observations.py:38-46); numpy's global RNG is consumed in the reference's order: tf times,
        tb times, then (err, noise) per forward epoch, then per backward epoch.
        """
        self.Npoints = Npoints
        self.error = error
        self.errorVar = errorVar
        nh = int(self.Npoints / 2.)
        self.tf = np.append([0], np.sort(np.random.uniform(0., tmax / 2., nh)))
        self.tb = np.sort(np.random.uniform(0., -tmax / 2., nh))
        times = np.concatenate((self.tf, self.tb))
        vx = state._rv_no_encounter_check(times)
        self.rvf = np.zeros(nh + 1)
        self.rvb = np.zeros(nh)
        self.errorf = np.zeros(nh + 1)
        self.errorb = np.zeros(nh)
        for i in range(len(self.tf)):
            self.errorf[i] = error + np.random.normal(0., self.errorVar)
            self.rvf[i] = vx[i] + np.random.normal(0., self.errorf[i])
        for i in range(len(self.tb)):
            self.errorb[i] = error + np.random.normal(0., self.errorVar)
            self.rvb[i] = vx[nh + 1 + i] + np.random.normal(0., self.errorb[i])
        self.t = np.concatenate((self.tb, self.tf), axis=0)
        self.rv = np.concatenate((self.rvb, self.rvf), axis=0)
        self.err = np.concatenate((self.errorb, self.errorf), axis=0)


def parse_vels(filename):
    """Three space-delimited columns: time [day], rv [m/s], err [m/s] (observations.py:57-59)."""
    data = np.genfromtxt(filename, usecols=(0, 1, 2), dtype='d')
    data = np.atleast_2d(data)
    return data[:, 0].copy(), data[:, 1].copy(), data[:, 2].copy()


class Observation_FromFile(Observation):
    def __init__(self, filename='yourfile.txt', Npoints=30):
        """Load observations from a .vels or .txt file (observations.py:52-69): unit factors 0.01720 and
        3.355e-5, np.array_split into a backward and a forward half, shift so the last backward epoch is 0."""
        readtimes, readrvs, readerrors = parse_vels(filename)
        readb, readf = np.array_split(readtimes * 0.01720, 2)
        shift = readb[len(readb) - 1]
        self.Npoints = Npoints
        self.tf = readf - shift
        self.tb = readb - shift
        self.rvb, self.rvf = np.array_split(readrvs * 3.355e-5, 2)
        self.errorb, self.errorf = np.array_split(readerrors * 3.355e-5, 2)
        self.t = np.concatenate((self.tb, self.tf), axis=0)
        self.rv = np.concatenate((self.rvb, self.rvf), axis=0)
        self.err = np.concatenate((self.errorb, self.errorf), axis=0)
