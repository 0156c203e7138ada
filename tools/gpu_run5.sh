mkdir -p gpurun_out
( REDTAIL_ENGINE_TRACE=1 timeout 50 python tools/rn18dbg.py 8 2>&1 | grep -v "resblock\|_conv[0-9]" | tail -45 ) > gpurun_out/rn_trace.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b5.json 2> gpurun_out/b5.err
REDTAIL_LIB_DIR=$PWD/redtail_b200/lib_scalar timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b5_scalar.json 2> gpurun_out/b5_scalar.err
(CONVBENCH_LAYERS=conv3D_2,conv3D_4 timeout 100 python tools/convbench.py 2>&1 | tail -3) > gpurun_out/t_cb.log
(REDTAIL_LIB_DIR=$PWD/redtail_b200/lib_scalar CONVBENCH_LAYERS=conv3D_2,conv3D_4 timeout 100 python tools/convbench.py 2>&1 | tail -3) >> gpurun_out/t_cb.log
cat gpurun_out/rn_trace.log gpurun_out/t_cb.log; for f in b5 b5_scalar; do cut -c1-200 gpurun_out/$f.json; done
