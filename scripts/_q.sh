for i in 1 2 3 4; do timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "cfg5" -s 2>&1 | grep -E "cfg5 SDXL|passed|failed" | cut -c1-220; done
