# Round-2 profile capture (run under gpurun): launch list with DRAM bytes of three NVSmall steps + ncu --set full of one step.
mkdir -p gpurun_out
REDTAIL_ENGINE_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_ncu.csv python tools/onestep.py 3 > gpurun_out/r02_launches.log 2>&1
REDTAIL_ENGINE_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none --launch-skip 23 -c 23 -f -o gpurun_out/prof_r02_step python tools/onestep.py 2 > gpurun_out/r02_full.log 2>&1
tail -3 gpurun_out/r02_launches.log gpurun_out/r02_full.log; ls -la gpurun_out/*.ncu-rep
