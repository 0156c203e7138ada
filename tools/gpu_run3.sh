mkdir -p gpurun_out
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12) > gpurun_out/t_all.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b3.json 2> gpurun_out/b3.err
for c in nvsmall_fp16 resnet18 resnet18_2d; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/b3_$c.json 2> gpurun_out/b3_$c.err; done
(CONVBENCH_LAYERS=conv3D_2 timeout 100 python tools/convbench.py 2>&1 | tail -3) > gpurun_out/t_cb.log
cat gpurun_out/t_all.log gpurun_out/t_cb.log; cut -c1-200 gpurun_out/b3.json; tail -3 gpurun_out/b3.err; for c in nvsmall_fp16 resnet18 resnet18_2d; do cut -c1-330 gpurun_out/b3_$c.json; tail -2 gpurun_out/b3_$c.err; done
