mkdir -p gpurun_out
F='resblock\|_conv[0-9]\|tensor  *[0-9]* \(left\|right\)'
( REDTAIL_ENGINE_DUMP_PLAN=1 REDTAIL_ENGINE_TRACE=1 timeout 30 python tools/rn18dbg.py 8 2>&1 | grep -v "$F" | tail -120 ) > gpurun_out/rn6_default.log
( REDTAIL_ENGINE_DSA=0 REDTAIL_ENGINE_DUMP_PLAN=1 REDTAIL_ENGINE_TRACE=1 timeout 30 python tools/rn18dbg.py 8 2>&1 | grep -v "$F" | grep "tensor\|step 8" | tail -80 ) > gpurun_out/rn6_dsa0.log
( REDTAIL_TC_DEBUG=32 REDTAIL_ENGINE_TRACE=1 timeout 30 python tools/rn18dbg.py 8 2>&1 | grep -v "$F" | tail -8 ) > gpurun_out/rn6_dbg32.log
( REDTAIL_TC_EW=8 REDTAIL_ENGINE_TRACE=1 timeout 30 python tools/rn18dbg.py 8 2>&1 | grep -v "$F" | tail -8 ) > gpurun_out/rn6_ew8.log
( REDTAIL_ENGINE_TRACE=1 timeout 30 python tools/rn18dbg.py 4 2>&1 | grep -v "$F" | tail -8 ) > gpurun_out/rn6_b4.log
( REDTAIL_ENGINE_TRACE=1 timeout 240 compute-sanitizer --tool memcheck --print-limit 20 python tools/rn18dbg.py 8 2>&1 | grep -v "$F" | tail -60 ) > gpurun_out/rn6_memcheck.log
tail -5 gpurun_out/rn6_dbg32.log gpurun_out/rn6_ew8.log gpurun_out/rn6_b4.log; tail -40 gpurun_out/rn6_memcheck.log
