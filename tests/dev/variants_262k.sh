python tests/dev/solve_timing.py C1:262144:f64 C2:131072:f64 2>&1 | grep -v "^$"
