mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_plugins.py -x -q -k "deconv_softargmax" 2>&1 | tail -25) > gpurun_out/t_dsa.log
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12) > gpurun_out/t_all.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b2.json 2> gpurun_out/b2.err
(CONVBENCH_LAYERS=conv3D_2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv3d_ds -c 1 -o gpurun_out/prof_ds python tools/convbench.py 2>&1 | tail -3) > gpurun_out/t_ncu.log
cat gpurun_out/t_dsa.log gpurun_out/t_all.log gpurun_out/t_ncu.log; cut -c1-200 gpurun_out/b2.json; tail -3 gpurun_out/b2.err
