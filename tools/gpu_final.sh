# Round-end validation (run under gpurun): GPU test-suite, smoke(), the bench line of every BASELINE config.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/final_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --config nvsmall_fp16 --steps 10 --warmup 3 > gpurun_out/final_nvsmall_fp16.json 2> /dev/null
timeout 300 python bench.py --config resnet18_2d --steps 5 --warmup 3 > gpurun_out/final_resnet18_2d.json 2> /dev/null
timeout 300 python bench.py --config resnet18 --steps 5 --warmup 3 > gpurun_out/final_resnet18.json 2> /dev/null
timeout 300 python bench.py --config trailnet --steps 10 --warmup 3 > gpurun_out/final_trailnet.json 2> /dev/null
cat gpurun_out/final_tests.log gpurun_out/final_smoke.log; for f in final_bench final_nvsmall_fp16 final_resnet18_2d final_resnet18 final_trailnet; do cut -c1-150 gpurun_out/$f.json; done
