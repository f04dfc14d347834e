mkdir -p gpurun_out
export JARM_RTOL=1e-6
for nf in 0 1; do
  echo "== NEG_FACTORS $nf" >> gpurun_out/r5.log
  GLEIPNIR_NEG_FACTORS=$nf timeout 200 python tools/jarm_check.py 40000 100 0 2>&1 | grep -E "batch|oracle" >> gpurun_out/r5.log
done
cat gpurun_out/r5.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_gaussian_dmma -c 1 -f -o gpurun_out/r02_i_dmma python tools/dmma_check.py > gpurun_out/r5_ncu.log 2>&1
tail -3 gpurun_out/r5_ncu.log
