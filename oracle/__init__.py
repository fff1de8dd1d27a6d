"""Test infrastructure only: the CPU oracle (sre_oracle.c), the reference build
recipe (_ref/), the lowering checker (lower_check.cpp) and the line-batch CPU
runner (cpu_baseline.py).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by sregex_b200."""
