# per-kernel instruction counts: bash scripts/gpu_inst.sh codec...
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes.sum --clock-control none --csv --log-file gpurun_out/inst_$c.csv python scripts/prof_one.py $c 0 both > gpurun_out/inst_$c.log 2>&1
  tail -1 gpurun_out/inst_$c.log
done
