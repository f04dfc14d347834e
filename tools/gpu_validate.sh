#!/bin/bash
# Round-end validation on a GPU box: smoke, the GPU test suite, the bench the way the driver calls it, the reference arm,
# and the ncu launch list of a short bench run.  Outputs under gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_validate.sh [steps] [warmup]'
STEPS=${1:-20}; WARMUP=${2:-5}
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/v_smoke.log 2>&1; tail -4 gpurun_out/v_smoke.log
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/v_tests.log 2>&1; tail -5 gpurun_out/v_tests.log
( time timeout 1500 python bench.py --gpus 1 --steps $STEPS --warmup $WARMUP ) > gpurun_out/v_bench.log 2>&1; tail -c 300 gpurun_out/v_bench.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/v_bench_ref.log 2>&1; tail -c 400 gpurun_out/v_bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu > gpurun_out/v_ncu.log 2>&1
tail -c 200 gpurun_out/v_ncu.log
