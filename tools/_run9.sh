B="python bench.py --steps 10 --warmup 3 --no-cpu --no-recipe --e2e-repeats 1"
$B > gpurun_out/p_sleep.json 2> gpurun_out/p_sleep.err
WE_PACK_SPIN_US=3000 $B > gpurun_out/p_spin.json 2> gpurun_out/p_spin.err
WE_PACK_SPIN_US=3000 WE_PACK_THREADS=12 $B > gpurun_out/p_spin12.json 2> gpurun_out/p_spin12.err
WE_PACK_SPIN_US=300 $B > gpurun_out/p_spin300.json 2> gpurun_out/p_spin300.err
