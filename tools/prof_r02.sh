# Round-2 evidence run (one B200): launch list of a short bench, ncu --set full of the three instantiations of the forward
# kernel, batch sweeps.  Each ncu command runs only after the same command exited 0 without ncu.
set -x
cd $GRAFT_REPO_ROOT
BENCH="python bench.py --steps 2 --warmup 3 --selfplay-seconds 0.05 --sustained-seconds 0.01 --no-cpu-baseline --no-config4"
$BENCH > gpurun_out/plain_bench4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_v4.csv $BENCH > gpurun_out/ncu_bench4.log 2>&1
cap() {  # name, target args...
    name=$1; shift
    python tools/ncu_target.py "$@" > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tower_r_umma -s 2 -c 1 -o /tmp/$name -f python tools/ncu_target.py "$@" > gpurun_out/ncu_$name.log 2>&1
    python tools/ncu_summary.py /tmp/$name.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:tower_r_umma -s 2 -c 1 python tools/ncu_target.py $*" > gpurun_out/r02_ncu_full_$name.json
    ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_$name.csv 2>/dev/null
}
cap tower_r_v2 1024 4
cp /tmp/tower_r_v2.ncu-rep gpurun_out/r02_tower_r_v2.ncu-rep
cap tower_r_b16 16 4
cap tower_r_split 1024 4 6 128 fp32
python tools/batch_sweep.py > gpurun_out/r02_batch_sweep_6x128.json 2> gpurun_out/sweep4.err
python tools/batch_sweep.py 6 128 fp32 > gpurun_out/r02_batch_sweep_6x128_fp32.json 2> gpurun_out/sweep5.err
du -sh gpurun_out
