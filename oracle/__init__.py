"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/cone_oracle.h).
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from cvxpylayers_b200/."""
