#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r03p_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/r03p_pytest.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-400
python - <<'PY'
# resynthesis timing: config-2 shaped BufNMF with resynthesis on 64 buffers, fused inverse vs the cuFFT pipeline
import os, sys, time
sys.path[:0] = [".", "flucoma-core_b200"]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio
a = torch.from_numpy(make_audio(64, 130816, distinct=8)).cuda()
out = torch.empty((64, 16, 130816), device="cuda")
for env in (None, "1"):
    if env: os.environ["FB200_ISTFT_CUFFT"] = env
    with fb.Plan(win=1024, hop=256, fft=1024) as plan:
        for _ in range(3):
            plan.bufnmf(a, 16, 5, seeds=np.arange(64), resynth=True)
            s = plan.stats()
        print("cufft" if env else "fused", "ms_resynth", round(s["ms_resynth"], 2), "ms_total", round(s["ms_total"], 2))
PY
