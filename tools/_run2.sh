set -x
mkdir -p gpurun_out
export JARM_RTOL=1e-6
timeout 200 python tools/jarm_profile.py 7104 > gpurun_out/r2_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ode_coop_eval -c 1 -f -o gpurun_out/r02_h_jarm_eval python tools/jarm_profile.py 7104 > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
