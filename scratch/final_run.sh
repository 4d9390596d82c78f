#!/bin/bash
# Round-end validation on one B200: GPU tests, both bench arms, ncu launch list + full capture of the engine, STFT-stage
# kernels with their DRAM bytes.  Every step under its own timeout.
tag=${1:-r01e}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_nmf_tc -s 1 -c 1 -o gpurun_out/${tag}_nmf_tc python profiles/profile_step.py 200 1024 > gpurun_out/${tag}_full.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_magnitude|k_frame_window|vector_fft" -s 12 -c 3 --csv --log-file gpurun_out/${tag}_stft_kernels.csv python profiles/profile_step.py 1 1024 > /dev/null 2>&1
timeout 120 python -m pytest tests/test_gpu_parity.py -q -s -k "config5" 2>&1 | grep -E "config 5|passed|failed"
ls gpurun_out | grep $tag
