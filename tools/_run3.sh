mkdir -p gpurun_out
export JARM_RTOL=1e-6
for lu in shfl smem; do for sf in select mask_t; do
  echo "== $lu $sf" >> gpurun_out/r3.log
  GLEIPNIR_LU_BROADCAST=$lu GLEIPNIR_SOLVE_FORM=$sf timeout 200 python tools/jarm_check.py 40000 100 0 2>&1 | grep -E "batch|oracle|rel " >> gpurun_out/r3.log
done; done
cat gpurun_out/r3.log
