mkdir -p gpurun_out
( echo "== default b=2"; timeout 60 python tools/rn18dbg.py 2 2>&1 | tail -40; echo "rc=$?" ) > gpurun_out/rn_a.log
( echo "== default b=8"; timeout 60 python tools/rn18dbg.py 8 2>&1 | tail -40; echo "rc=$?" ) > gpurun_out/rn_b.log
if grep -q "3 steps" gpurun_out/rn_b.log; then echo "b=8 OK" >> gpurun_out/rn_b.log; else
( echo "== DS=0 b=8"; REDTAIL_TC_DS=0 timeout 60 python tools/rn18dbg.py 8 2>&1 | tail -40 ) > gpurun_out/rn_c.log
( echo "== DSA=0 b=8"; REDTAIL_ENGINE_DSA=0 timeout 60 python tools/rn18dbg.py 8 2>&1 | tail -40 ) > gpurun_out/rn_d.log
fi
(CONVBENCH_LAYERS=conv3D_2 timeout 100 python tools/convbench.py 2>&1 | tail -3) > gpurun_out/t_cb.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b4.json 2> gpurun_out/b4.err
REDTAIL_TC_DEBUG=16 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b4_nopf.json 2> gpurun_out/b4_nopf.err
REDTAIL_TC_DEBUG=48 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b4_old.json 2> gpurun_out/b4_old.err
cat gpurun_out/rn_*.log gpurun_out/t_cb.log; for f in b4 b4_nopf b4_old; do cut -c1-200 gpurun_out/$f.json; done
