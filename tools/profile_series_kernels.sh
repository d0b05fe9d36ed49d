prof() {
  python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline $4 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o gpurun_out/prof_r1c_$1 python bench.py --workload $2 --steps 2 --warmup 1 --no-cpu-baseline $4 > gpurun_out/ncu_$1.log 2>&1
  echo "prof $1 rc=$?"
}
prof gram_s10 gram-s10 series_gram_fwd "--cands 200000"
prof gram_so3 gram-so3 series_gram_fwd "--cands 200000"
prof gram_s2 gram-s2 series_gram_fwd "--cands 200000"
prof gram_t10 gram-t10 series_gram_fwd "--cands 50000"
