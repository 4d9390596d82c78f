#!/bin/bash
cd $GRAFT_REPO_ROOT
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
    print(os.environ.get("FB200_LIB", "default")[-9:], os.environ.get("FB200_ISTFT_S"), "ms_resynth", round(s["ms_resynth"], 2))
PY
V=$GRAFT_REPO_ROOT/flucoma-core_b200/lib/variants
python /tmp/rs.py
for S in 8192 6144 4096; do FB200_LIB=$V/stft3.so FB200_ISTFT_S=$S python /tmp/rs.py; done
