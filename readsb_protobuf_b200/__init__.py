"""B200-native Mode S / ADS-B demodulator behind readsb's mag_buf / demodulate2400() boundary.

The product is the C-ABI library built from csrc/ (include/readsb_b200.h); this package is
the thin Python host layer used by the tests and bench.py.  See DESIGN.md.
"""
__all__ = ["build", "synth", "results"]
