timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "capture" 2>&1 | tail -3 | cut -c1-300
python scripts/attn_store_probe.py --impl 0 --cases cfg5,n500,n500s32,sd15,n100 2>&1 | tee gpurun_out/r2y_probe_reg.jsonl
