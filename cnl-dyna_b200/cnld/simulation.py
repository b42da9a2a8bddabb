"""
cnld.simulation -- time-domain contact simulation on top of the impulse-response database (mirror of
/root/reference/cnld/simulation.pyx: fir_conv_cy :51-77, contact laws :80-99, StateDB :102-294,
FixedStepSolver :297-555, drive signals :640-712).

The arithmetic runs in libcnld_b200 (csrc/firconv.cu): the patch-to-patch table, the total-pressure history and
every state array stay on one GPU.  Stepped one sample at a time, a step is one pass over the table plus the
fixed-point iterations of the current sample in a single CTA (`cnld_b200_td_*`); `step(n)` / `solve()` advance in
blocks of 16 samples per table pass (the history before the block is convolved once for all 16, the lags inside
the block are added per sample; the pass for the next block runs under the current block's samples) and never
return to the host in between; `step()` called sample by sample is served from a look-ahead of whole blocks.  This module keeps the reference's
objects and names around it: the constructor resamples the table and the drive exactly as the reference does
(scipy interp1d, host, once), `StateDB` holds the host copies the properties hand out.

Not mirrored: CompensationSolver (:558-637, marked experimental upstream; its per-patch callables come from
cnld/compensation.py, outside SURVEY.md section 8).
"""
import types

import numpy as np
import scipy.signal
from scipy.constants import epsilon_0 as e_0
from scipy.interpolate import interp1d

from . import _lib, database, impulse_response

_FIELDS = ('x', 'u', 'p_es', 'p_cont_spr', 'p_cont_dmp', 'p_tot')


def p_es_pp(v, x, g_eff):
    """Electrostatic pressure of a parallel-plate capacitor (:17-21)."""
    return -e_0 / 2 * v**2 / (x + g_eff)**2


def fir_conv_cy(fir, p, dt, offset):
    """x[j] = dt * sum_i sum_k fir[i, j, k + offset] * p[nsample - 1 - k, i], k < min(nsample, nfir - offset)
    (:51-77), on the device.  The table is uploaded per call; the solver keeps it resident instead."""
    tab = _lib.FirTable(np.asarray(fir, dtype=np.float64))
    try:
        return tab.conv(np.asarray(p, dtype=np.float64), dt, int(offset))
    finally:
        tab.close()


def fir_conv_py(fir, p, fs, offset):
    """The reference's pure-Python variant (:24-48) divides by the sampling rate instead of multiplying by dt."""
    return fir_conv_cy(fir, p, 1.0 / fs, offset)


def make_p_cont_spr(k, n, x0):
    """Contact spring k (x0 - x)^n below x0 (:80-88).  Float output for every input -- the reference's
    np.vectorize wrapper turns integer whenever its first element is out of contact; the device solver
    reproduces that inside FixedStepSolver, this helper does not."""
    def p_cont_spr(x):
        x = np.asarray(x, dtype=np.float64)
        pen = np.where(x < x0, x0 - x, 0.0)
        return np.where(x < x0, k * pen**n, 0.0)
    return p_cont_spr


def make_p_cont_dmp(lmbd, n, x0):
    """Contact damper -(lmbd (x0 - x)^n) xdot below x0 (:91-99); see make_p_cont_spr for the dtype remark."""
    def p_cont_dmp(x, xdot):
        x = np.asarray(x, dtype=np.float64)
        pen = np.where(x < x0, x0 - x, 0.0)
        return np.where(x < x0, -(lmbd * pen**n) * np.asarray(xdot, dtype=np.float64), 0.0)
    return p_cont_dmp


class StateDB:
    """Host copies of the state variables on the solver's time grid (:102-294): time t[nt], drive v[nt, npatch]
    (linear interpolation of t_v, end values held) and x, u, p_es, p_cont_spr, p_cont_dmp, p_tot [nt, npatch]."""

    class State(types.SimpleNamespace):
        pass

    def __init__(self, t_lim, t_v, npatch):
        v_t, v = t_v
        t_start, t_stop, t_step = t_lim
        self._t = np.arange(t_start, t_stop + t_step, t_step)
        v = np.asarray(v, dtype=np.float64)
        if v.ndim <= 1:
            v = np.tile(v, (npatch, 1)).T
        held = (v[0, :], v[-1, :])
        self._v = interp1d(v_t, v, axis=0, kind='linear', bounds_error=False, fill_value=held, assume_sorted=True)(self._t)
        self._i = np.arange(len(self._t))
        self._npatch = npatch
        self.clear()

    def clear(self):
        for name in _FIELDS:
            setattr(self, '_' + name, np.zeros((len(self._t), self._npatch)))

    i = property(lambda self: self._i.view(), doc='sample index')
    t = property(lambda self: self._t.view(), doc='time')
    v = property(lambda self: self._v.view(), doc='voltage')
    x = property(lambda self: self._x.view(), doc='displacement')
    u = property(lambda self: self._u.view(), doc='velocity')
    p_es = property(lambda self: self._p_es.view(), doc='electrostatic pressure')
    p_cont_spr = property(lambda self: self._p_cont_spr.view(), doc='contact spring pressure')
    p_cont_dmp = property(lambda self: self._p_cont_dmp.view(), doc='contact damper pressure')
    p_tot = property(lambda self: self._p_tot.view(), doc='total pressure')

    def get_state_i(self, i):
        rows = {name: getattr(self, '_' + name)[i, :] for name in _FIELDS}
        return self.State(i=self._i[i], t=self._t[i], v=self._v[i, :], **rows)

    def get_state_t(self, t):
        return self.get_state_i(int(np.argmin(np.abs(t - self._t))))

    def set_state_i(self, s):
        for name in _FIELDS:
            val = getattr(s, name, None)
            if val is not None:
                getattr(self, '_' + name)[s.i, :] = val

    def set_state_t(self, s):
        s.i = int(np.argmin(np.abs(s.t - self._t)))
        self.set_state_i(s)


class FixedStepSolver:
    """Fixed-step time-domain solver (:297-555) with the reference's constructor, properties and stepping
    interface; the state lives on the GPU (`_lib.TimeDomainSolver`) and is mirrored into `StateDB`."""

    Properties = types.SimpleNamespace

    def __init__(self, t_fir, t_v, gap, gap_eff, t_lim, k, n, x0, lmbd, atol=1e-10, maxiter=5, device=0):
        fir_t, fir = t_fir
        fir = np.asarray(fir, dtype=np.float64)
        t_step = t_lim[2]
        npatch = fir.shape[0]
        # the table on the solver's step: cubic resampling (:311-314)
        self._fir_t = np.arange(fir_t[0], fir_t[-1], t_step)
        self._fir = interp1d(fir_t, fir, axis=-1, kind='cubic', assume_sorted=True)(self._fir_t)
        self._gap = np.array(gap)
        self._gap_eff = np.array(gap_eff)
        self._db = StateDB(t_lim, t_v, npatch)
        self.atol = atol
        self.maxiter = maxiter
        self.npatch = npatch
        self.min_step = t_step
        self._p_cont_spr = make_p_cont_spr(k, n, x0)
        self._p_cont_dmp = make_p_cont_dmp(lmbd, n, x0)
        self._dev = _lib.TimeDomainSolver(self._fir, self._db.v, self._gap_eff, t_step, k, n, x0, lmbd, atol=atol,
                                          maxiter=maxiter, e0=e_0, device=device)
        self._start()

    @classmethod
    def from_array_and_db(cls, array, dbfile, t_v, t_lim, k, n, x0, lmbd, atol=1e-10, maxiter=5, calc_fir=False,
                          use_kkr=True, interp=4, device=0):
        """Solver from an array and its database (:338-365): impulse responses read from the file, or computed
        from its patch-to-patch frequency responses when calc_fir is set."""
        if calc_fir:
            freqs, ppfr = database.read_patch_to_patch_freq_resp(dbfile)
            fir_t, fir = impulse_response.fft_to_fir(freqs, ppfr, interp=interp, axis=-1, use_kkr=use_kkr)
        else:
            fir_t, fir = database.read_patch_to_patch_imp_resp(dbfile)
        gap, gap_eff = [], []
        for elem in array.elements:
            for mem in elem.membranes:
                for _ in mem.patches:
                    gap.append(mem.gap)
                    gap_eff.append(mem.gap + mem.isolation / mem.permittivity)
        return cls((fir_t, fir), t_v, gap, gap_eff, t_lim, k, n, x0, lmbd, atol, maxiter, device=device)

    # -- device -> host -------------------------------------------------------------------------------------
    # The device runs ahead of `current_step`: the drive is fixed at construction, so a call that needs one more
    # sample advances the device by LOOKAHEAD samples (whole blocks of 16 per table pass instead of one pass per
    # sample) into private arrays; rows become visible in the StateDB only when `step()` reaches them, exactly when
    # the reference would have computed them.
    LOOKAHEAD = 64

    def _start(self):
        nt = len(self._db.t)
        self._ahead = {name: np.zeros((nt, self.npatch)) for name in _FIELDS}
        self._error_all = np.zeros(nt)
        self._iters_all = np.zeros(nt, dtype=np.int32)
        self._done = 1                                   # samples final on the device (its current_step)
        self.current_step = 1
        self._dev.state(0, 1, into=self._ahead)
        self._reveal(0, 1)

    def _advance(self, upto):
        nt = len(self._db.t)
        need = upto - self._done
        if need <= 0:
            return
        n = need if self._done + need > nt else min(max(need, self.LOOKAHEAD), nt - self._done)
        self._dev.step(n)                                # raises past the end of the time grid, like the library
        lo, hi = self._done, self._done + n
        self._dev.state(lo, hi, into=self._ahead)
        self._error_all[lo:hi], self._iters_all[lo:hi] = self._dev.info(lo, hi)
        self._done = hi

    def _reveal(self, i0, i1):
        for name in _FIELDS:
            getattr(self._db, '_' + name)[i0:i1] = self._ahead[name][i0:i1]

    # -- the reference's read-only views ----------------------------------------------------------------------
    time = property(lambda self: self._db.t[:self.current_step])
    voltage = property(lambda self: self._db.v[:self.current_step, :])
    displacement = property(lambda self: self._db.x[:self.current_step, :])
    velocity = property(lambda self: self._db.u[:self.current_step, :])
    pressure_electrostatic = property(lambda self: self._db.p_es[:self.current_step, :])
    pressure_contact_spring = property(lambda self: self._db.p_cont_spr[:self.current_step, :])
    pressure_contact_damper = property(lambda self: self._db.p_cont_dmp[:self.current_step, :])
    pressure_total = property(lambda self: self._db.p_tot[:self.current_step, :])
    error = property(lambda self: np.array(self._error_all[1:self.current_step]))
    iters = property(lambda self: np.array(self._iters_all[1:self.current_step]))
    props = property(lambda self: self.Properties(gap=self._gap, gap_eff=self._gap_eff))

    def get_current_state(self):
        return self._db.get_state_i(self.current_step)

    def get_previous_state(self):
        return self._db.get_state_i(self.current_step - 1)

    def get_state_i(self, i):
        return self._db.get_state_i(i)

    def get_state_d(self, d):
        return self._db.get_state_i(self.current_step + d)

    def get_state_t(self, t):
        return self._db.get_state_t(t)

    # -- stepping ----------------------------------------------------------------------------------------------
    def step(self, nsteps=1):
        """Advance by one sample (:459-503), or by `nsteps` samples."""
        first = self.current_step
        self._advance(first + nsteps)
        self.current_step = first + nsteps
        self._reveal(first, self.current_step)

    def solve(self):
        """All remaining samples (:505-514; the reference's loop refers to a `_time` attribute that does not
        exist and fails -- this is its evident intent: step until current_step reaches len(t) - 1)."""
        stop = len(self._db.t) - 1
        if self.current_step < stop:
            self.step(stop - self.current_step)

    def reset(self):
        self._db.clear()
        self._dev.reset()
        self._start()

    def close(self):
        self._dev.close()

    def __iter__(self):
        return self

    def __next__(self):
        if self.current_step >= len(self._db.t) - 1:
            raise StopIteration
        self.step()
        return self.current_step


class CompensationSolver(FixedStepSolver):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('CompensationSolver (experimental upstream, simulation.pyx:558-637) depends on '
                                  'cnld/compensation.py, which is outside the path this library covers')


# ---- drive signals (:640-712) -----------------------------------------------------------------------------
def gaussian_pulse(fc, fbw, fs, td=0, tpr=-60, antisym=True):
    """Gaussian-modulated pulse cut at the tpr level; antisym picks the quadrature component."""
    cutoff = scipy.signal.gausspulse('cutoff', fc=fc, bw=fbw, tpr=tpr, bwr=-3)
    half = np.ceil(cutoff * fs) / fs
    t = np.arange(-half, half + 1 / fs / 2, 1 / fs)
    inphase, quad = scipy.signal.gausspulse(t, fc=fc, bw=fbw, retquad=True, bwr=-3)
    t += td
    sig = quad if antisym else inphase
    return t, sig / np.max(sig)


def logistic_ramp(tr, dt, td=0, tstop=None, tpr=-60):
    """DC ramp with rise time tr along the logistic function."""
    rate = 2 * np.log(10**(-tpr / 20)) / tr
    half = np.ceil(tr / 2 / dt) * dt
    t = np.arange(-half, (half if tstop is None else tstop - td) + dt / 2, dt)
    v = 1 / (1 + np.exp(-rate * t))
    t += td
    return t, v


def linear_ramp(tr, dt, td=0, tstop=None):
    """Linear DC ramp reaching 1 at tr."""
    if tstop is None:
        t = np.arange(0, np.ceil(tr / dt) * dt, dt)
    else:
        t = np.arange(0, tstop - td + dt / 2, dt)
    v = np.where(t > tr, 1.0, t / tr)
    t += td
    return t, v


def winsin(f, ncycle, dt, td=0):
    """ncycle periods of a sine, end samples forced to zero."""
    half = round(ncycle * 1 / f / 2 / dt) * dt
    t = np.arange(-half, half + dt / 2, dt)
    v = np.sin(2 * np.pi * f * t)
    v[0] = v[-1] = 0
    t += td
    return t, v


def sigadd(*args):
    """Sum of (t, v) signals on a common grid, each extended by its edge values."""
    starts = [t[0] for t, _ in args]
    tmin = min(starts)
    dt = args[0][0][1] - args[0][0][0]
    lead = [int(round((t0 - tmin) / dt)) for t0 in starts]
    total = max(a + len(v) for a, (_, v) in zip(lead, args))
    out = np.zeros(total)
    for a, (_, v) in zip(lead, args):
        out += np.pad(v, (a, total - a - len(v)), mode='edge')
    return np.arange(total) * dt + tmin, out
