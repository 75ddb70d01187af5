#!/bin/bash
mkdir -p gpurun_out
echo "== bench fp32 + cudnn.benchmark"; timeout 900 python bench.py --steps 5 --warmup 4 --precision fp32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','skp_share_of_step')})" | tee gpurun_out/c_fp32.log
echo "== bench reference + cudnn.benchmark"; timeout 900 python bench.py --steps 5 --warmup 4 --precision reference --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','skp_share_of_step')})" | tee gpurun_out/c_ref.log
echo "== full parity"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full" --timeout 800 2>&1 | grep "full-size\|candidate\|passed\|failed" | tee gpurun_out/c_full.log
