mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_trailnet.py -x -q 2>&1 | tail -30) > gpurun_out/t_tn.log
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/t_all.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b11.json 2> gpurun_out/b11.err
(timeout 300 python tools/tnprof.py 256 2>&1 | tail -60) > gpurun_out/tn_prof.log
timeout 300 python bench.py --config trailnet --steps 5 --warmup 3 > gpurun_out/b11_trailnet.json 2> gpurun_out/b11_trailnet.err
cat gpurun_out/t_tn.log gpurun_out/t_all.log; tail -4 gpurun_out/tn_prof.log; cut -c1-200 gpurun_out/b11.json; cut -c1-300 gpurun_out/b11_trailnet.json; tail -3 gpurun_out/b11_trailnet.err
