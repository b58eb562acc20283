"""CPU oracle for the isoneutral hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under veros_b200/ does.
"""
