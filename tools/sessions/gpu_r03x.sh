#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_spectral.py -m gpu -q -x 2>&1 | tail -2
cat > /tmp/st.py <<'PY'
import os, sys
sys.path[:0] = [".", "flucoma-core_b200"]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio
a = torch.from_numpy(make_audio(1024, 130816, distinct=8)).cuda()
with fb.Plan(win=1024, hop=256, fft=1024) as plan:
    ts = []
    for _ in range(4):
        plan.bufnmf(a, 16, 1, seeds=np.arange(1024))
        ts.append(plan.stats()["ms_stft"])
    print("ms_stft", " ".join("%.3f" % t for t in ts))
PY
python /tmp/st.py
