timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "capture or attn_store" 2>&1 | tail -3 | cut -c1-300
python scripts/attn_store_probe.py --impl 0 --cases cfg5,n500,n500s32,sd15,n100 2>&1 | tee gpurun_out/r4a_probe_fixed.jsonl
SKP_ATTN_STORE_DYNAMIC=1 SKP_ATTN_STORE_BIG=0 python scripts/attn_store_probe.py --impl 0 --cases cfg5,n500,n500s32 2>&1 | tee gpurun_out/r4a_probe_dyn.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:capture_store_reg -s 2 -c 1 -o gpurun_out/r4a_store_reg_cfg5 python scripts/attn_store_probe.py --impl 0 --cases cfg5 --reps 1 > gpurun_out/r4a_ncu.log 2>&1
