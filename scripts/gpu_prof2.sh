# usage: bash scripts/gpu_prof2.sh <tag> <kernel-regex> <what> codec...   (debug sweep, timings, launch list + full capture of first codec)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; KRE=$2; WHAT=$3; shift 3
( timeout 600 python tests/gpu_debug.py ) > gpurun_out/debug.log 2>&1
grep -E "BAD|TOTAL|Error|error" gpurun_out/debug.log | head -20
for c in "$@"; do timeout 300 python scripts/prof_one.py $c 5 both; done
C=$1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof_one.py $C 2 $WHAT > gpurun_out/${TAG}_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 6 -c 8 -o gpurun_out/${TAG}_full python scripts/prof_one.py $C 2 $WHAT > gpurun_out/${TAG}_f.log 2>&1
tail -2 gpurun_out/${TAG}_f.log
