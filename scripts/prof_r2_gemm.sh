set -e
python scripts/prof_gemm_w8.py > gpurun_out/r2_gemm_w8_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:whd_gemm_kernel -s 1 -c 1 -o gpurun_out/r2_gemm_w8 python scripts/prof_gemm_w8.py > gpurun_out/r2_gemm_w8_ncu.log 2>&1
cat gpurun_out/r2_gemm_w8_plain.log
