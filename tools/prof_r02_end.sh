# End-of-round evidence (one B200): launch list of the short bench with the final kernels (tree kernel after the rewrite of
# the chess primitives, gate kernel in the tier1_deep runs), and ncu --set full of the stand-alone gate kernel.  Each ncu
# command runs only after the same command exited 0 without ncu.
set -x
cd $GRAFT_REPO_ROOT
BENCH="python bench.py --steps 2 --warmup 3 --selfplay-seconds 0.05 --sustained-seconds 0.01 --no-cpu-baseline --no-config4"
$BENCH > gpurun_out/plain_bench5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches_v5.csv $BENCH > gpurun_out/ncu_bench5.log 2>&1
python tools/gate_kernel_once.py > gpurun_out/plain_gate.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:device_tier1_gates_kernel -c 1 -o gpurun_out/gate_kernel_v2 -f python tools/gate_kernel_once.py > gpurun_out/ncu_gate2.log 2>&1
tail -2 gpurun_out/ncu_gate2.log
