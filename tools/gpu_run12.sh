mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_trailnet.py -x -q 2>&1 | tail -6) > gpurun_out/t_tn.log
(timeout 300 python tools/tnprof.py 256 2>&1 | tail -34 | head -8; timeout 100 python tools/tnprof.py 256 2>&1 | tail -1) > gpurun_out/tn_prof.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b12.json 2> gpurun_out/b12.err
REDTAIL_ENGINE_IM2COL_MINK=64 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b12_i2c.json 2> gpurun_out/b12_i2c.err
(REDTAIL_ENGINE_IM2COL_MINK=64 timeout 300 python -m pytest tests/test_gpu_net.py -x -q -s -k "nvsmall_parity or nvtiny_parity or synthetic" 2>&1 | tail -12) > gpurun_out/t_i2c.log
cat gpurun_out/t_tn.log gpurun_out/tn_prof.log gpurun_out/t_i2c.log; cut -c1-130 gpurun_out/b12.json; cut -c1-130 gpurun_out/b12_i2c.json
