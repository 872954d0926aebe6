"""FDC -- B200-native drop-in for the `FDC` Python module of gr-FDC (python/__init__.py in the reference
imports the SWIG factories and the hier block under these same names)."""
from ._cabi import FDCError, LIB_PATH                                             # noqa: F401
from .blocks import (overlap_save, vector_cut_vxx, phase_shifting_windowing_vcc, fft_vcc, Channelizer,  # noqa: F401
                     opt_channelparams, psw_tables)
try:                                                                               # activity-gated blocks + hier block
    from .activity import PowerActivationChannel, SegmentDetection, activity_detection_channelizer_vcm  # noqa: F401
    from .FrequencyDomainChannelizer import FrequencyDomainChannelizer, FREQMODE, VERBOSEMODE            # noqa: F401
except ImportError:                                                                # pragma: no cover
    pass

from .waterfall import WaterfallMsgTagging                                         # noqa: F401

RECTANGULAR, HANN, RAMP = 0, 1, 2
