mkdir -p gpurun_out
: > gpurun_out/rn7.log
for i in 1 2 3 4 5 6 7 8; do for b in 4 8; do ( timeout 40 python tools/rn18dbg.py $b 2>&1 | tail -1 ) >> gpurun_out/rn7.log; done; done
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/t_all.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b7.json 2> gpurun_out/b7.err
timeout 300 python bench.py --config resnet18 --steps 5 --warmup 3 > gpurun_out/b7_resnet18.json 2> gpurun_out/b7_resnet18.err
cat gpurun_out/rn7.log gpurun_out/t_all.log; cut -c1-200 gpurun_out/b7.json; cut -c1-400 gpurun_out/b7_resnet18.json
