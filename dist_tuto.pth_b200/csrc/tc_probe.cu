// Layout probes for the sm_100a tensor-core path (no GPU on the dev box: every assumption about TMA swizzle images and
// UMMA shared-memory descriptors is checked on hardware by tests/test_gpu_tc_probe.py before the training kernels rely on it).
//
//   tma_probe  : ONE cp.async.bulk.tensor.{2..5}d box load with a caller-defined tensor map (uint16 elements, so values are
//                exact) -> raw shared-memory image of the box, as TMA wrote it (swizzle included).
//   umma_probe : caller-provided shared-memory images of A and B, a caller-provided instruction descriptor and a list of
//                (A descriptor, B descriptor, TMEM column, accumulate) MMAs -> dump of TMEM [128 lanes x ncols] fp32.
//                Descriptors are given relative to the image (start address = byte offset in the image); the kernel
//                adds the shared-memory base.  Any operand major / swizzle / LBO / SBO hypothesis is testable from Python.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>
#include <string>

#include "tc_common.cuh"

namespace probe {

std::string g_err;

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, int c3, int c4, uint32_t bytes,
                 uint8_t* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) tile[i] = 0xEE;     // poison: bytes TMA did not write stay 0xEE
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  tc::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_expect_tx(&bar, bytes);
    const uint32_t dst = tc::smem_u32(tile), b = tc::smem_u32(&bar);
    if (rank == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&map), "r"(b), "r"(c0), "r"(c1) : "memory");
    else if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(dst), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else if (rank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(dst), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(dst), "l"(&map), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
  }
  tc::mbar_wait(&bar, 0);
  for (uint32_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

struct MmaOp {
  uint64_t adesc, bdesc;     // start-address fields relative to the A / B image
  uint32_t tmem_col, accumulate;
};
constexpr int kMaxOps = 64;
struct MmaList {
  MmaOp op[kMaxOps];
  int n;
};

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes, const uint8_t* __restrict__ b_img, uint32_t b_bytes,
                  uint32_t idesc, const __grid_constant__ MmaList ops, int ncols, float* __restrict__ dump) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sa = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sb = sa + ((a_bytes + 1023) & ~1023u);
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (uint32_t i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  if (threadIdx.x == 0) { tc::mbar_init(&done, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&tmem_slot);
  tc::fence_proxy_async();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = tmem_slot;
  // zero the dumped TMEM window first (lanes an instruction does not write must read back as 0)
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  if (threadIdx.x == 0) {
    const uint64_t abase = (uint64_t)((tc::smem_u32(sa) & 0x3FFFF) >> 4), bbase = (uint64_t)((tc::smem_u32(sb) & 0x3FFFF) >> 4);
    for (int i = 0; i < ops.n; ++i)
      tc::umma_bf16(tmem + ops.op[i].tmem_col, ops.op[i].adesc + abase, ops.op[i].bdesc + bbase, idesc, ops.op[i].accumulate);
    tc::commit(&done);
  }
  __syncwarp();
  tc::mbar_wait(&done, 0);
  tc::fence_after();
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) dump[(size_t)(warp * 32 + lane) * ncols + c0 + i] = __uint_as_float(r[i]);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) { tc::fence_after(); tc::tmem_dealloc<512>(tmem); }
}

}  // namespace probe

extern "C" {

const char* b2_probe_last_error() { return probe::g_err.c_str(); }

// dims/strides/box: innermost first; strides in BYTES for dims 1..rank-1 (stride of dim 0 is the element size, 2 B)
int b2_tma_probe(const void* tensor, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                 const unsigned int* box, int swizzle, const int* coords, unsigned int bytes, unsigned char* out,
                 cudaStream_t stream) {
  auto enc = probe::get_encode();
  if (!enc) { probe::g_err = "cuTensorMapEncodeTiled not available"; return -1; }
  if (rank < 2 || rank > 5) { probe::g_err = "rank must be 2..5"; return -2; }
  CUtensorMap map;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, (cuuint32_t)rank, const_cast<void*>(tensor), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { probe::g_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return -3; }
  const size_t smem = bytes + 1024;
  cudaFuncSetAttribute(probe::tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe::tma_probe_kernel<<<1, 128, smem, stream>>>(map, rank, coords[0], coords[1], rank > 2 ? coords[2] : 0,
                                                     rank > 3 ? coords[3] : 0, rank > 4 ? coords[4] : 0, bytes, out);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { probe::g_err = std::string("launch: ") + cudaGetErrorString(ce); return -4; }
  return 0;
}

// ops: n x {adesc, bdesc, tmem_col, accumulate} as 4 uint64 each
int b2_umma_probe(const unsigned char* a_img, unsigned int a_bytes, const unsigned char* b_img, unsigned int b_bytes,
                  unsigned int idesc, const unsigned long long* ops, int n_ops, int ncols, float* dump, cudaStream_t stream) {
  if (n_ops < 1 || n_ops > probe::kMaxOps || ncols < 16 || ncols > 512 || (ncols % 16) != 0 || (a_bytes % 16) || (b_bytes % 16)) {
    probe::g_err = "bad probe arguments";
    return -1;
  }
  probe::MmaList l;
  memset(&l, 0, sizeof(l));
  l.n = n_ops;
  for (int i = 0; i < n_ops; ++i) {
    l.op[i].adesc = ops[4 * i]; l.op[i].bdesc = ops[4 * i + 1];
    l.op[i].tmem_col = (uint32_t)ops[4 * i + 2]; l.op[i].accumulate = (uint32_t)ops[4 * i + 3];
  }
  const size_t smem = ((a_bytes + 1023) & ~1023u) + ((b_bytes + 1023) & ~1023u) + 1024;
  if (smem > 200 * 1024) { probe::g_err = "images too large"; return -2; }
  cudaFuncSetAttribute(probe::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe::umma_probe_kernel<<<1, 128, smem, stream>>>(a_img, a_bytes, b_img, b_bytes, idesc, l, ncols, dump);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) { probe::g_err = std::string("launch: ") + cudaGetErrorString(ce); return -4; }
  return 0;
}

}  // extern "C"
