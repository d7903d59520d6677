timeout 300 python tests/dev/lookahead_dbg.py C1:4096:f64 C3:4096:f64 C1:4096:f32 C2:2048:f64 2>&1 | grep -v "^    " 
timeout 300 python tests/dev/lookahead_ab.py C1:4096:f64 C1:1024:f64 C3:4096:f64 C1:256:f64 2>&1
python tests/dev/la_stages.py C1:4096:f64 2>&1 | tail -1; python tests/dev/la_stages.py C1:1024:f64 2>&1 | tail -1
