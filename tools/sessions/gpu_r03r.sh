#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r03r_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/r03r_pytest.log | tail -6
cat > /tmp/rs.py <<'PY'
import os, sys
sys.path[:0] = [".", "flucoma-core_b200"]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio
a = torch.from_numpy(make_audio(64, 130816, distinct=8)).cuda()
with fb.Plan(win=1024, hop=256, fft=1024) as plan:
    for _ in range(3):
        r = plan.bufnmf(a, 16, 5, seeds=np.arange(64), resynth=True)
        s = plan.stats()
    print(os.environ.get("FB200_ISTFT_CUFFT"), "ms_resynth", round(s["ms_resynth"], 2), float(r["resynth"].abs().sum()))
PY
FB200_ISTFT_CUFFT=1 python /tmp/rs.py
python /tmp/rs.py
