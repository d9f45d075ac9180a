# map sweep with the measured int8 tensor-pipe peak, up to 200k cells (C4 shape)
mkdir -p gpurun_out
timeout 1200 python scripts/map_sweep.py --ns 20000,50000,100000,200000 --ms 40 --Ss 50 > gpurun_out/r2_map_sweep.md 2>&1
timeout 600 python scripts/map_sweep.py --ns 20000,50000 --ms 60,200 --Ss 100 2>&1 | tail -8 >> gpurun_out/r2_map_sweep.md
cat gpurun_out/r2_map_sweep.md
