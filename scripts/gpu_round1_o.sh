#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or conv or groupnorm" --timeout 300 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s -k "full or stage1" --timeout 800 2>&1 | grep -E "full-size|passed|failed|Error"
echo "== gemm bench"; timeout 300 python scripts/kernel_bench.py --only "gemm_nt_tc(kernel" 2>&1 | cut -c1-220
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','skp_kernel_ms_total')})"
timeout 600 python scripts/profile_step.py --table gpurun_out/o_step_table.json 2>&1 | tail -14 | cut -c1-150
