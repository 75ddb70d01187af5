timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "capture or attn_store" 2>&1 | tail -3 | cut -c1-300
echo "rows4 px2"; python scripts/attn_store_probe.py --impl 0 --cases cfg5,sd15,n100 2>&1 | tee gpurun_out/r4j_probe_r4px2.jsonl | cut -c1-160
echo "rows4 px1"; SKP_ATTN_STORE_PX=1 python scripts/attn_store_probe.py --impl 0 --cases cfg5,sd15 2>&1 | tee gpurun_out/r4j_probe_r4px1.jsonl | cut -c1-160
echo "rows2 px2"; SKP_ATTN_STORE_ROWS=2 python scripts/attn_store_probe.py --impl 0 --cases cfg5,sd15 2>&1 | cut -c1-160
