# usage: bash scripts/gpu_prof.sh <tag> <codec> <kernel-regex> [enc|dec|both]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; CODEC=$2; KRE=$3; WHAT=${4:-both}
python scripts/prof_one.py $CODEC 5 $WHAT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof_one.py $CODEC 2 $WHAT > gpurun_out/${TAG}_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 6 -c 6 -o gpurun_out/${TAG}_full python scripts/prof_one.py $CODEC 2 $WHAT > gpurun_out/${TAG}_f.log 2>&1
tail -3 gpurun_out/${TAG}_f.log
