# ncu --set full capture of the kernels matching a regex for one codec.  usage: gpu_ncu.sh <tag> <codec> <regex> [enc|dec|both] [count]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=$1; C=$2; RE=$3; WHAT=${4:-both}; CNT=${5:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $CNT -c $CNT -f -o gpurun_out/${TAG}_full_$C python scripts/prof_one.py $C 0 $WHAT > gpurun_out/${TAG}_full_$C.log 2>&1
tail -2 gpurun_out/${TAG}_full_$C.log
ls -la gpurun_out/${TAG}_full_$C.ncu-rep
