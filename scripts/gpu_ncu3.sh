# usage: bash scripts/gpu_ncu3.sh <tag> then triples: <kernel-regex> <enc|dec> <codec> <skip> <count> ...   (full captures)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; shift
while [ $# -ge 5 ]; do
  KRE=$1; WHAT=$2; CODEC=$3; SKIP=$4; CNT=$5; shift 5
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_${KRE}_${CODEC} python scripts/prof_one.py $CODEC 0 $WHAT > gpurun_out/${TAG}_${KRE}_${CODEC}.log 2>&1
  tail -1 gpurun_out/${TAG}_${KRE}_${CODEC}.log
done
ls -la gpurun_out | tail -8
