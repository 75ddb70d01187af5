#!/bin/bash
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or conv or gemm" --timeout 300 2>&1 | tail -4
echo "== pipeline fp32 sdpa"; timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s --timeout 800 2>&1 | grep -E "full-size|candidate|passed|failed|Error"
echo "== pipeline fp16 sdpa"; SKP_SELF_ATTN=fp16 timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -s --timeout 800 2>&1 | grep -E "full-size|candidate|passed|failed|Error|assert"
for sa in fp32 fp16; do
echo "== bench self-attn $sa"; SKP_SELF_ATTN=$sa timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','e2e','early_exit_images_per_s_1gpu','skp_kernel_ms_total')})"
done
timeout 600 python scripts/profile_step.py --table gpurun_out/l_step_table.json 2>&1 | tail -28 | cut -c1-150
