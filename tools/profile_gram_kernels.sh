set -x
python bench.py --workload config1-s2 --steps 10 --warmup 3 > gpurun_out/bench_config1.log 2>&1; tail -c 1500 gpurun_out/bench_config1.log
python bench.py --workload config2-bo-s3 --steps 2 --warmup 3 > gpurun_out/bench_config2.log 2>&1; tail -c 1800 gpurun_out/bench_config2.log
prof() { # name workload kernel-regex extra-args
  python bench.py --workload $2 --steps 2 --warmup 1 $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -o gpurun_out/prof_r1b_$1 python bench.py --workload $2 --steps 2 --warmup 1 $4 > gpurun_out/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
prof gram_s10 gram-s10 series_gram_fwd "--cands 200000"
prof gram_so3 gram-so3 series_gram_fwd "--cands 200000"
prof gram_s2 gram-s2 series_gram_fwd "--cands 200000"
prof gram_t10 gram-t10 series_gram_fwd "--cands 50000"
prof gram_h2 gram-h2 hyp_gram "--cands 1000 --train 1000"
prof gram_spd2 gram-spd2 spd2_lattice "--cands 1000 --train 1000"
ls -la gpurun_out/*.ncu-rep
