for m in 0 1; do echo "== CILQR_LA_SERIAL=$m"; CILQR_LA_SERIAL=$m timeout 300 python tests/dev/lookahead_dbg.py C1:16384:f64 C2:2048:f64 C1:16384:f64 2>&1 | grep "la rep\|repeatable" ; done
