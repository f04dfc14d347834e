set -x
mkdir -p gpurun_out
export JARM_RTOL=1e-6
(GLEIPNIR_COOP_MIN_BLOCKS=3 timeout 300 python tools/jarm_check.py 20000 100 0) > gpurun_out/r1_min3.log 2>&1
(timeout 400 python tools/jarm_check.py 20000 11840 1) > gpurun_out/r1_min4.log 2>&1
timeout 600 python -m pytest tests/test_gpu_jarm.py tests/test_gpu_ode.py -x -q > gpurun_out/r1_tests.log 2>&1
tail -5 gpurun_out/r1_tests.log
