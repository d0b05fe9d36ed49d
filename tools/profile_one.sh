# usage: bash tools/profile_one.sh <workload> <kernel-regex> <report-name> [skip-count]   (one GPU, under gpurun)
# plain run first (must exit 0), then ONE ncu --set full capture of one launch of the matching kernel
mkdir -p gpurun_out
ARGS="--workload $1 --steps 2 --warmup 3 --cands 1000 --train 1000 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/plain_$3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-3} -c 1 -o gpurun_out/$3 -f python bench.py $ARGS > gpurun_out/ncu_$3.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_$3.log
