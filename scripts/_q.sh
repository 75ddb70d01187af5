python scripts/capture_tc_trace.py 16 77 128 8 store 2>&1 | head -40
timeout 300 python scripts/capture_bench.py --json gpurun_out/r2f_capture_bench.json 2>&1 | cut -c1-420
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "capture" 2>&1 | tail -8
