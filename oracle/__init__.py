"""ORACLE — test infrastructure only.

CPU restatements of the reference hot path used as the parity checker.  Nothing in the product
package (autoware_vision_pilot_b200/) may import from here; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do.
"""
