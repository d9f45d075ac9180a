# keyed UPGMA pruned scan: parity tests, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruned_scan.py -x -q 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_exact.py -x -q -k "widened or wide_range" 2>&1 | tail -3
timeout 600 python scripts/upgma_keyed_time.py --cells 10000 20000 > gpurun_out/r2_upgma_keyed.md 2>&1; cat gpurun_out/r2_upgma_keyed.md
