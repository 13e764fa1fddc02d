"""pyfr_b200: a B200-native (sm_100a) execution backend for the
flux-reconstruction right-hand-side evaluation of PyFR."""

__version__ = '0.1.0'
