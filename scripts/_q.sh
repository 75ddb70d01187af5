python scripts/capture_tc_trace.py 16 77 128 8 mean 2>&1 | cut -c1-100 | awk '/producer/||/mma:/||/epilogue/||/it  [5-7]:/||/CTA 0/'
timeout 300 python scripts/capture_bench.py --json gpurun_out/r2j_capture_bench.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print({k:v for k,v in d.items() if k in ('case','tc_us','simt_us','tc_rel_err','tc_frac_hbm','simt_frac_hbm','tc_rowsum_err')})"
timeout 400 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "capture" 2>&1 | tail -4
