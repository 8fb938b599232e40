"""TEST INFRASTRUCTURE — ctypes access to the CPU checker (oracle/liboracle.so, oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (crdmodel_b200) never does.
"""
from .pyoracle import *  # noqa: F401,F403
