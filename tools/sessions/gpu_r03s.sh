#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/rs.py <<'PY'
import os, sys
sys.path[:0] = [".", "flucoma-core_b200"]
import numpy as np, torch
import flucoma_b200 as fb
from bench import make_audio
a = torch.from_numpy(make_audio(64, 130816, distinct=8)).cuda()
with fb.Plan(win=1024, hop=256, fft=1024) as plan:
    for _ in range(2):
        r = plan.bufnmf(a, 16, 2, seeds=np.arange(64), resynth=True)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_istft_fused -s 1 -c 1 -o gpurun_out/r03s_istft python /tmp/rs.py > gpurun_out/r03s_ncu.log 2>&1; tail -1 gpurun_out/r03s_ncu.log
