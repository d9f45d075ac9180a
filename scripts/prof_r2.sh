set -e
python scripts/prof_loop.py --cells 50000 --iters 5010 --mode exact > gpurun_out/r2_prof2_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 15000 -c 90 --csv --log-file gpurun_out/r2_launches_exact_k45k.csv python scripts/prof_loop.py --cells 50000 --iters 5010 --mode exact > gpurun_out/r2_prof2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"whd_exact_kernel|merge_kernel|scan_pruned|prune_rowtest" -s 0 -c 1 -o gpurun_out/r2_whd_exact python scripts/prof_loop.py --cells 50000 --iters 4 --mode exact > gpurun_out/r2_prof2_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"merge_kernel|scan_pruned|prune_rowtest" -s 15000 -c 3 -o gpurun_out/r2_loop_exact_v2 python scripts/prof_loop.py --cells 50000 --iters 5010 --mode exact > gpurun_out/r2_prof2_ncu3.log 2>&1
tail -2 gpurun_out/r2_prof2_plain.log
