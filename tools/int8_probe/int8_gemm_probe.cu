// Feasibility probe for DESIGN.md section 11 (error-free INT8 slicing of the posterior's N^2 contraction): the rate a
// LIBRARY int8 x int8 -> int32 GEMM (CUTLASS 4.x collective builder, tcgen05 kind::i8) reaches on this GPU at the shape
// of one slice product (M = candidate tile rows, N = K = training points).  A probe under tools/, not on any product
// path.  Build: see tools/int8_probe/build.sh.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/util/packed_stride.hpp"

using namespace cute;

using ElementA = int8_t;
using ElementB = int8_t;
using ElementC = int32_t;
using ElementAcc = int32_t;
using LayoutA = cutlass::layout::RowMajor;
using LayoutB = cutlass::layout::ColumnMajor;
using LayoutC = cutlass::layout::RowMajor;
constexpr int AlignA = 16, AlignB = 16, AlignC = 4;

using ArchTag = cutlass::arch::Sm100;
using OpClass = cutlass::arch::OpClassTensorOp;
#ifndef TILE_M
#define TILE_M _128
#define TILE_N _128
#define CL_M _1
#endif
#ifndef TILE_K
#define TILE_K _128
#endif
#ifndef CL_N
#define CL_N _1
#endif
using TileShape = Shape<TILE_M, TILE_N, TILE_K>;
using ClusterShape = Shape<CL_M, CL_N, _1>;

using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    ArchTag, OpClass, TileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc,
    ElementC, LayoutC, AlignC, ElementC, LayoutC, AlignC, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;

using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    ArchTag, OpClass, ElementA, LayoutA, AlignA, ElementB, LayoutB, AlignB, ElementAcc, TileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;

using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 16384, N = argc > 2 ? atoi(argv[2]) : 5120, K = argc > 3 ? atoi(argv[3]) : 5120;
  std::vector<int8_t> ha((size_t)M * K), hb((size_t)N * K);
  unsigned s = 12345u;
  for (auto& v : ha) { s = s * 1664525u + 1013904223u; v = (int8_t)((int)(s >> 24) - 128); }
  for (auto& v : hb) { s = s * 1664525u + 1013904223u; v = (int8_t)((int)(s >> 24) - 128); }
  int8_t *da, *db; int32_t* dc;
  CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size())); CK(cudaMalloc(&dc, sizeof(int32_t) * (size_t)M * N));
  CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));

  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {da, sa, db, sb},
                                {{1, 0}, dc, sc, dc, sd}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) { printf("cannot implement\n"); return 2; }
  size_t ws = Gemm::get_workspace_size(args);
  void* dws = nullptr;
  if (ws) CK(cudaMalloc(&dws, ws));
  if (gemm.initialize(args, dws) != cutlass::Status::kSuccess) { printf("initialize failed\n"); return 3; }
  for (int i = 0; i < 3; ++i) if (gemm.run() != cutlass::Status::kSuccess) { printf("run failed\n"); return 4; }
  CK(cudaDeviceSynchronize());
  // spot check against the host
  std::vector<int32_t> hc(64);
  CK(cudaMemcpy(hc.data(), dc, sizeof(int32_t) * 64, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int j = 0; j < 64; ++j) {
    long long r = 0;
    for (int k = 0; k < K; ++k) r += (int)ha[k] * (int)hb[(size_t)j * K + k];
    if ((int32_t)r != hc[j]) ++bad;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) gemm.run();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double ops = 2.0 * M * N * (double)K * reps;
  printf("{\"probe\": \"cutlass int8 gemm (tcgen05 kind::i8)\", \"M\": %d, \"N\": %d, \"K\": %d, \"ms_per_gemm\": %.4f, \"tops\": %.1f, \"mismatches_in_first_64\": %d}\n",
         M, N, K, ms / reps, ops / (ms * 1e-3) / 1e12, bad);
  return bad ? 5 : 0;
}
