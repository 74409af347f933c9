#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_metasnv_e2e.py -x -q -m gpu -k "hand_written or golden_fixture or metasnv or empty or two_splits or classic" --timeout 300 > gpurun_out/r2ai_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2ai_pytest.log | cut -c1-300
timeout 600 python tools/e2e_startup_ab.py --rounds 6 > gpurun_out/r2ai_ab_2gb.txt 2> gpurun_out/r2ai_ab_2gb.err
echo "ab rc=$?"; python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2ai_ab_2gb.txt')]
for e in ("1","0"):
    w=[r["wall_s"] for r in rows if r["early_decode"]==e]; t=[r["total_s"] for r in rows if r["early_decode"]==e]
    print("early_decode", e, "wall", [round(x,2) for x in w], "median", sorted(w)[len(w)//2], "in-process", [round(x,2) for x in t])
PY
