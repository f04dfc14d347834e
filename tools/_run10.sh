mkdir -p gpurun_out
export JARM_RTOL=1e-6
for mb in 4 5; do
  echo "== MIN_BLOCKS $mb" >> gpurun_out/r10.log
  GLEIPNIR_COOP_MIN_BLOCKS=$mb timeout 200 python tools/jarm_check.py 40000 100 0 2>&1 | grep -E "batch|oracle" >> gpurun_out/r10.log
done
cat gpurun_out/r10.log
