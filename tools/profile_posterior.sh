# Launch list + one full capture of the default bench workload (posterior S^10), per B200_PROFILING.md.
set -x
CMD="python bench.py --steps 1 --warmup 1 --cands 151552 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_post.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_posterior_step.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_post2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:posterior_reduce -s 2 -c 1 -o gpurun_out/prof_r1b_posterior $CMD > gpurun_out/ncu_post.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain_post.log | cut -c1-400
