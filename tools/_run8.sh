python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1
tail -5 gpurun_out/gputest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-recipe --e2e-repeats 1"
$B > gpurun_out/packed.json 2> gpurun_out/packed.err
WE_PACK_THREADS=8 $B > gpurun_out/packed8.json 2> gpurun_out/packed8.err
$B --pos-upload 1 > gpurun_out/whole.json 2> gpurun_out/whole.err
