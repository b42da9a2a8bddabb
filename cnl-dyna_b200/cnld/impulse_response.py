"""
cnld.impulse_response -- frequency response -> impulse response (mirror of
/root/reference/cnld/impulse_response.py: interp_fft :161-175, one_to_two :11-46,
kramers_kronig2 :115-125, fft_to_fir :191-206, fft_to_sir :209-224).  Two engines with the same
results: 'device' = libcnld_b200 (cnld_b200_fft_to_fir: interpolation kernel + one exactly-twiddled DFT
GEMM on the FP64 tensor cores, used by generate_db.postprocess and the field-pressure path whenever a GPU
is present) and 'host' = the numpy / scipy.fft restatement of the reference's own module below.
"""
import numpy as np
from scipy.fft import fftfreq, ifft
from scipy.signal import hilbert


def interp_fft(f, s, n, axis=-1):
    """Linear interpolation of magnitude and unwrapped phase on an n-times finer grid."""
    if n == 1:
        return f, s
    f = np.asarray(f, dtype=np.float64)
    s = np.moveaxis(np.asarray(s), axis, -1)
    fi = np.linspace(f[0], f[-1], (len(f) - 1) * n + 1)
    mag, phase = np.abs(s), np.unwrap(np.angle(s), axis=-1)
    lo = np.clip(np.searchsorted(f, fi, side='right') - 1, 0, len(f) - 2)
    slope_m = (mag[..., lo + 1] - mag[..., lo]) / (f[lo + 1] - f[lo])
    slope_p = (phase[..., lo + 1] - phase[..., lo]) / (f[lo + 1] - f[lo])
    mi = slope_m * (fi - f[lo]) + mag[..., lo]
    pi = slope_p * (fi - f[lo]) + phase[..., lo]
    return fi, np.moveaxis(mi * np.exp(1j * pi), -1, axis)


def one_to_two(f, s, axis=-1, odd=False):
    """One-sided spectrum of a real signal -> two-sided (conjugate mirror); energy not halved."""
    s = np.moveaxis(np.asarray(s, dtype=np.complex128), axis, -1)
    nf = s.shape[-1]
    nfft = 2 * nf - 1 if odd else (nf - 1) * 2
    out = np.zeros(s.shape[:-1] + (nfft,), dtype=np.complex128)
    if odd:
        out[..., :nfft // 2 + 1] = s
        out[..., nfft // 2 + 1:] = np.conj(s[..., -1:0:-1])
    else:
        out[..., :nfft // 2] = s[..., :-1]
        out[..., nfft // 2:] = np.conj(s[..., -1:0:-1])
    fs = (f[1] - f[0]) * nfft
    return fftfreq(nfft, 1 / fs), np.moveaxis(out, -1, axis)


def kramers_kronig2(f, s, axis=-1):
    """Rebuild the imaginary part from the real part (Hilbert transform); no front delay."""
    re = np.real(s)
    return re + 1j * (-np.imag(hilbert(re, axis=axis)))


def _to_time(f, s, interp, use_kkr, axis):
    fi, si = interp_fft(f, s, n=interp, axis=axis)
    f2s, s2s = one_to_two(fi, si, axis=axis, odd=True)
    if use_kkr:
        s2s = kramers_kronig2(f2s, s2s, axis=axis)
    nfft = len(f2s)
    fs = (f2s[1] - f2s[0]) * nfft
    return np.arange(nfft) / fs, np.real(ifft(s2s, axis=axis)) * fs


def _engine(engine):
    if engine in ('host', 'device'):
        return engine
    from . import _lib
    return 'device' if _lib.device_count() > 0 else 'host'


def _device_to_time(f, s, interp, use_kkr, axis, device):
    from . import _lib
    t, h = _lib.fft_to_fir(f, np.moveaxis(np.asarray(s), axis, -1), interp=interp, use_kkr=use_kkr, device=device)
    return t, np.moveaxis(h, -1, axis)


def fft_to_fir(f, s, interp=1, use_kkr=True, axis=-1, engine=None, device=0):
    """One-sided FFT -> impulse response of a causal LTI system."""
    if _engine(engine) == 'device':
        return _device_to_time(f, s, interp, use_kkr, axis, device)
    return _to_time(f, s, interp, use_kkr, axis)


def fft_to_sir(f, s, interp=1, use_kkr=False, axis=-1, engine=None, device=0):
    """One-sided FFT -> spatial impulse response."""
    if _engine(engine) == 'device':
        return _device_to_time(f, s, interp, use_kkr, axis, device)
    return _to_time(f, s, interp, use_kkr, axis)
