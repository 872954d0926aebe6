"""Python restatement of the hier block's frequency conversion and channel geometry
(python/FrequencyDomainChannelizer.py:37-40, 70-91, 322-345 in the reference), with the Python-2
semantics GNU Radio 3.7 runs it under (int/int floors; round() rounds half away from zero).

Used by the host mirror (gr-fdc_b200/python/FDC/FrequencyDomainChannelizer.py), the workloads and the tests;
tests/golden/make_golden.py checks it against the reference's own function source.
"""
import math


def nextpow2(k):
    if k < 1:
        raise ValueError('Cannot evaluate next power 2 of {}'.format(k))
    return 2 ** int(math.ceil(math.log2(k)))


def py2_round(x):
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def get_freq(f):          # normalized mode, :70
    return (f + 0.5) % 1.0


def get_bw(bw):           # :72
    return bw % 1.0


def get_opt_channelparams(blocksize, relinvovl, freq, bw):
    passsamps = blocksize * bw
    blocklen = nextpow2(passsamps)
    if blocklen < 1.2 * passsamps:
        blocklen *= 2
    passband = float(passsamps) / float(blocklen) * 1.1
    stopband = 1.0
    if passband >= 1.0:
        passband = 1.0
    elif passband < 0.7:
        stopband = passband + 0.25
    freqsamps = int(py2_round(freq * blocksize)) % blocksize
    freqsamps -= blocklen // 2
    if freqsamps < 0:
        freqsamps = (freqsamps + blocksize) % blocksize
    if freqsamps + blocklen > blocksize:
        freqsamps = blocksize - blocklen
    outputblocklen = int(blocklen) - int(blocklen) // relinvovl
    return int(freqsamps), int(blocklen), int(outputblocklen), float(passband), float(stopband)
