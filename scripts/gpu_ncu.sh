# usage: bash scripts/gpu_ncu.sh <tag> <codec> <kernel-regex>   (launch list + full capture, one codec, one enc+dec)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; CODEC=$2; KRE=$3
timeout 300 python scripts/prof_one.py $CODEC 5 both | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof_one.py $CODEC 1 both > gpurun_out/${TAG}_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 12 -c 10 -o gpurun_out/${TAG}_full python scripts/prof_one.py $CODEC 1 both > gpurun_out/${TAG}_f.log 2>&1
tail -2 gpurun_out/${TAG}_f.log
ls -la gpurun_out
