for m in 0 1; do echo "== CILQR_LA_COST_LATE=$m"; CILQR_LA_COST_LATE=$m python tests/dev/lookahead_ab.py C1:256:f64 C1:512:f64 C1:1024:f64 2>&1 | grep "lookahead=1" | awk 'NR%2==0'; done
python tests/dev/lookahead_ab.py C1:256:f64 C1:512:f64 C1:1024:f64 2>&1 | grep "lookahead=0\|same" | awk 'NR%3!=1'
