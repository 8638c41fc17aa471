python bench.py > gpurun_out/r2_bench_h_C4_1gpu.json 2> gpurun_out/r2_bench_h.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-recipe --e2e-repeats 1 --frames-per-step 32"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_ncu_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_gate|k_rad|k_hb|k_label|k_cov' --launch-skip 7 --launch-count 7 -o gpurun_out/r2h_prof_full -f $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
