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

    # --- GPU residency (not in the reference): the device copy lives in the Context's cache, keyed on the CONTENT of the
    # arrays, so edited / replaced data is re-uploaded (the reference reads obs on every call) and nothing outlives
    # its context ---
    def _handle(self, ctx):
        return ctx.obs_handle(self)


def _join_legs(obs):
    """Flat (backward, forward) views the plotting / saving helpers use (observations.py:48-50, 66-68)."""
    obs.t = np.concatenate((obs.tb, obs.tf))
    obs.rv = np.concatenate((obs.rvb, obs.rvf))
    obs.err = np.concatenate((obs.errorb, obs.errorf))


class FakeObservation(Observation):
    def __init__(self, state, Npoints=30, error=0., errorVar=0., tmax=1.5):
        """Synthetic data from a known State (observations.py:18-50): Npoints/2 + 1 forward epochs (t = 0 first) and
        Npoints/2 backward epochs, uniform in +-tmax/2 and sorted; per-epoch error bar error + N(0, errorVar) and
        velocity vx + N(0, error bar).

        numpy's legacy global RNG is consumed exactly as in the reference -- uniform(tf), uniform(tb), then one
        (error, noise) pair of normals per forward epoch, then per backward epoch -- here as ONE standard_normal block,
        which is the same stream (normal(m, s) is m + s * gauss()).  The noiseless velocities come from one GPU
        integration without an encounter distance that visits tf in order, then tb in order (observations.py:26-46)."""
        self.Npoints = Npoints
        self.error = error
        self.errorVar = errorVar
        half = int(Npoints / 2.)
        self.tf = np.concatenate(([0.], np.sort(np.random.uniform(0., tmax / 2., half))))
        self.tb = np.sort(np.random.uniform(0., -tmax / 2., half))
        vx = state._rv_no_encounter_check(np.concatenate((self.tf, self.tb)))
        z = np.random.standard_normal(2 * (2 * half + 1)).reshape(-1, 2)
        bars = error + self.errorVar * z[:, 0]
        noisy = vx + bars * z[:, 1]
        self.errorf, self.errorb = bars[:half + 1].copy(), bars[half + 1:].copy()
        self.rvf, self.rvb = noisy[:half + 1].copy(), noisy[half + 1:].copy()
        _join_legs(self)


def parse_vels(filename):
    """Three space-delimited columns: time [day], rv [m/s], err [m/s] (observations.py:57-59)."""
    data = np.atleast_2d(np.genfromtxt(filename, usecols=(0, 1, 2), dtype='d'))
    return data[:, 0].copy(), data[:, 1].copy(), data[:, 2].copy()


def write_vels(filename, obs):
    """Inverse of Observation_FromFile's unit conversion: code units back to day, m/s, m/s (time origin = last backward
    epoch).  The reference's driver.save_obs writes the velocity column twice instead of the errors (driver.py:222); this
    writes the three columns a .vels file is read with."""
    np.savetxt(filename, np.column_stack((obs.t / DAY_TO_CODE, obs.rv / MS_TO_CODE, obs.err / MS_TO_CODE)), fmt="%.10g")


DAY_TO_CODE = 0.01720      # day -> yr/2pi           (observations.py:60)
MS_TO_CODE = 3.355e-5      # m/s -> AU/(yr/2pi)      (observations.py:61-62)


class Observation_FromFile(Observation):
    def __init__(self, filename='yourfile.txt', Npoints=30):
        """Observations from a .vels / .txt file (observations.py:52-69): unit factors 0.01720 and 3.355e-5, the rows
        split by np.array_split into a backward and a forward half, times shifted so the last backward epoch is 0.
        Npoints is the caller's normaliser (driver.read_obs passes 100 whatever the row count)."""
        days, ms, ms_err = parse_vels(filename)
        self.Npoints = Npoints
        tb, tf = np.array_split(days * DAY_TO_CODE, 2)
        self.tb, self.tf = tb - tb[-1], tf - tb[-1]
        self.rvb, self.rvf = np.array_split(ms * MS_TO_CODE, 2)
        self.errorb, self.errorf = np.array_split(ms_err * MS_TO_CODE, 2)
        _join_legs(self)
