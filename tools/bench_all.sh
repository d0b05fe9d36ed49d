# usage (one GPU, under gpurun): bash tools/bench_all.sh <tag>
# every bench.py workload once, JSON lines into gpurun_out/bench_<tag>_<workload>.json; tools/bench_summary.py tabulates them
TAG=${1:-r}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err
for w in config1-s2 gram-s2 gram-so3 gram-s10 gram-t10 gram-h2 gram-spd2 gram-spd3 gram-h2-table; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
done
python bench.py --workload config2-bo-s3 --steps 2 --warmup 3 > gpurun_out/bench_${TAG}_config2-bo-s3.json 2> gpurun_out/bench_${TAG}_config2.err
python bench.py --workload tr-step --steps 2 --warmup 3 > gpurun_out/bench_${TAG}_tr-step.json 2> gpurun_out/bench_${TAG}_tr.err
tail -c 300 gpurun_out/bench_${TAG}_default.json; ls -la gpurun_out/bench_${TAG}_*.json | wc -l
