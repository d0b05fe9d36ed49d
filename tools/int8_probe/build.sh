# builds tools/int8_probe/int8_gemm_probe (CUTLASS headers vendored in the image's flashinfer wheel; probe only)
set -e
CUT=$(python -c "import flashinfer,os;print(os.path.join(os.path.dirname(flashinfer.__file__),'data','cutlass'))")
cd "$(dirname "$0")"
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --expt-extended-lambda \
  -I$CUT/include -I$CUT/tools/util/include -DNDEBUG -o int8_gemm_probe int8_gemm_probe.cu
