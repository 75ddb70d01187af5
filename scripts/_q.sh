timeout 300 python scripts/xattn_bench.py 2>&1 | tail -12 | cut -c1-300
python scripts/attn_store_probe.py --impl 2 --cases cfg5,sd15 2>&1 | tail -2
timeout 200 python scripts/capture_bench.py 2>&1 | tail -8 | cut -c1-300
