"""CPU oracle for the F-16 longitudinal hot path. TEST INFRASTRUCTURE ONLY:
import from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs,
never from the product package."""
