# usage: bash scripts/gpu_ncu2.sh <tag> <kernel-regex> <enc|dec|both> codec...   (full capture of the matching kernels, one pass per codec)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; KRE=$2; WHAT=$3; shift 3
for CODEC in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 4 -o gpurun_out/${TAG}_${CODEC} python scripts/prof_one.py $CODEC 0 $WHAT > gpurun_out/${TAG}_${CODEC}.log 2>&1
  tail -1 gpurun_out/${TAG}_${CODEC}.log
done
ls -la gpurun_out | tail -8
