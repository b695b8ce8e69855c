"""CPU oracle (TEST INFRASTRUCTURE ONLY): see oracle.hpp.  Importable only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs."""
