mkdir -p gpurun_out
GLEIPNIR_DMMA_NO_PIPELINE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_gaussian_dmma_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r02_j_dmma python tools/dmma_check.py > gpurun_out/r8_ncu.log 2>&1
tail -3 gpurun_out/r8_ncu.log
