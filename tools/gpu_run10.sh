mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_plugins.py -x -q -k "deconv_softargmax or depth_stationary" 2>&1 | tail -4) > gpurun_out/t_k.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b10.json 2> gpurun_out/b10.err
REDTAIL_TC_DEBUG=64 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b10_l1pf.json 2> gpurun_out/b10_l1pf.err
REDTAIL_TC_MAXN=128 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b10_maxn128.json 2> gpurun_out/b10_maxn128.err
REDTAIL_TC_MINSTAGES=3 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b10_min3.json 2> gpurun_out/b10_min3.err
cat gpurun_out/t_k.log; for f in b10 b10_l1pf b10_maxn128 b10_min3; do cut -c1-120 gpurun_out/$f.json; done
