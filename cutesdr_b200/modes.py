"""Demodulator mode constants and per-mode filter limits / defaults.

Mirrors the reference's mode numbering (dsp/demodulator.h:20-28), the fixed
per-mode filter limits the GUI pushes through `SetDemod`
(gui/mainwindow.cpp:1004-1047) and the persisted-settings defaults
(gui/mainwindow.cpp:442-452). `tDemodInfo` is carried as a plain dict with the
reference's field names (minus the GUI-only QString / click resolution).
"""

DEMOD_AM, DEMOD_SAM, DEMOD_FM, DEMOD_USB, DEMOD_LSB, DEMOD_CWU, DEMOD_CWL = range(7)
NUM_DEMODS = 7
MODE_NAMES = ("AM", "SAM", "FM", "USB", "LSB", "CWU", "CWL")

INFO_FIELDS = ("HiCut", "HiCutmin", "HiCutmax", "LowCut", "LowCutmin", "LowCutmax", "Offset",
               "SquelchValue", "AgcSlope", "AgcThresh", "AgcManualGain", "AgcDecay", "AgcOn", "AgcHangOn")

# (HiCutmin, HiCutmax, LowCutmin, LowCutmax) -- gui/mainwindow.cpp:1004-1047
_LIMITS = {
    DEMOD_AM: (500, 10000, -10000, -500),
    DEMOD_SAM: (100, 10000, -10000, -100),
    DEMOD_FM: (5000, 15000, -15000, -5000),
    DEMOD_USB: (500, 20000, 0, 200),
    DEMOD_LSB: (-200, 0, -20000, -500),
    DEMOD_CWU: (50, 1000, -1000, -50),
    DEMOD_CWL: (50, 1000, -1000, -50),
}


def demod_info(mode, HiCut=5000, LowCut=-5000, Offset=0, SquelchValue=0, AgcSlope=0, AgcThresh=-100,
               AgcManualGain=30, AgcDecay=200, AgcOn=True, AgcHangOn=False):
    """tDemodInfo for `mode` with the reference's settings defaults."""
    himin, himax, lomin, lomax = _LIMITS[mode]
    return dict(HiCut=int(HiCut), HiCutmin=himin, HiCutmax=himax, LowCut=int(LowCut), LowCutmin=lomin,
                LowCutmax=lomax, Offset=int(Offset), SquelchValue=int(SquelchValue), AgcSlope=int(AgcSlope),
                AgcThresh=int(AgcThresh), AgcManualGain=int(AgcManualGain), AgcDecay=int(AgcDecay),
                AgcOn=int(bool(AgcOn)), AgcHangOn=int(bool(AgcHangOn)))


def max_bandwidth(mode, info):
    """The bandwidth `CDemodulator::SetDemod` hands to `SetDataRate` (dsp/demodulator.cpp:116-120)."""
    if mode in (DEMOD_LSB, DEMOD_CWL):
        return float(-info["LowCutmin"])
    return float(info["HiCutmax"])
