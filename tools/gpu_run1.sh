mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_plugins.py -x -q -k "depth_stationary" 2>&1 | tail -25) > gpurun_out/t_ds.log
(timeout 300 python -m pytest tests/test_gpu_plugins.py -x -q -k "transpose" 2>&1 | tail -8) > gpurun_out/t_tr.log
(CONVBENCH_LAYERS=conv3D_2 timeout 200 python tools/convbench.py '{"REDTAIL_TC_DS":"0"}' 2>&1 | tail -5) > gpurun_out/t_cb.log
(timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/t_all.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b1.json 2> gpurun_out/b1.err
REDTAIL_TC_INTERLEAVE=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b1_noint.json 2> gpurun_out/b1_noint.err
cat gpurun_out/t_ds.log gpurun_out/t_tr.log gpurun_out/t_cb.log gpurun_out/t_all.log; cut -c1-300 gpurun_out/b1.json
