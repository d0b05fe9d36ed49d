mkdir -p gpurun_out
python bench.py --workload gram-h2 --steps 2 --warmup 3 --cands 1000 --train 1000 --no-cpu-baseline > gpurun_out/plain_h2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hyp_gram -s 3 -c 1 -o gpurun_out/prof_r1d_gram_h2 -f python bench.py --workload gram-h2 --steps 2 --warmup 3 --cands 1000 --train 1000 --no-cpu-baseline > gpurun_out/ncu_h2.log 2>&1
echo rc=$?; tail -2 gpurun_out/ncu_h2.log
