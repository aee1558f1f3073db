"""TEST INFRASTRUCTURE ONLY: Python front-ends of the two parity checkers.

    oracle.port  -- the C restatement (oracle/modes_oracle.c), loaded through ctypes
    oracle.ref   -- the unmodified reference built from /root/reference (oracle/_ref/ref_demod)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under readsb_protobuf_b200/ does.
"""
