python -m pytest tests -m gpu -x -q -k "chunking or packed or full_size_production or pinned_ring or api_vs" > gpurun_out/gputest_sub.log 2>&1
tail -3 gpurun_out/gputest_sub.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-recipe --e2e-repeats 1"
WE_PACK_SPIN_US=3000 $B > gpurun_out/q_spin.json 2> gpurun_out/q_spin.err
$B > gpurun_out/q_sleep.json 2> gpurun_out/q_sleep.err
WE_GRAPHS=0 WE_PACK_SPIN_US=3000 $B > gpurun_out/q_nograph.json 2> gpurun_out/q_nograph.err
