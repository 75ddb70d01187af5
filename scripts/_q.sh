timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "capture" 2>&1 | tail -2 | cut -c1-200
python scripts/capture_bwd_probe.py --tokens 500; python scripts/capture_bwd_probe.py --tokens 77
