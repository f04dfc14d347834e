mkdir -p gpurun_out
( timeout 600 python tools/walker_trace.py 40000 2368 1 ) > gpurun_out/r9_k2368.log 2>&1
grep "^step" gpurun_out/r9_k2368.log
( timeout 900 python tools/walker_trace.py 100000 10000 1 ) > gpurun_out/r9_k10000.log 2>&1
grep "^step" gpurun_out/r9_k10000.log
