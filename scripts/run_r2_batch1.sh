# Round-2 follow-up checks on one B200: new parity tests, strict mode at the bench size, map sweep with the measured
# int8 peak (up to 200k cells), ncu capture of the final fixed-point tensor kernel.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_exact.py tests/test_gpu_scale.py -x -q 2>&1 | tail -5
python scripts/exact_time.py --cells 50000 --strict-max 50000 > gpurun_out/r2_exact_time_strict50k.md 2>&1; tail -4 gpurun_out/r2_exact_time_strict50k.md
python scripts/map_sweep.py --ns 50000,100000,200000 --ms 40 --Ss 50 > gpurun_out/r2_map_sweep.md 2>&1
python scripts/map_sweep.py --ns 20000,50000 --ms 60 --Ss 100 2>&1 | tail -4 >> gpurun_out/r2_map_sweep.md
cat gpurun_out/r2_map_sweep.md
bash scripts/prof_r2_gemm.sh
