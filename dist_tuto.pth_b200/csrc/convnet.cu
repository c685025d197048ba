// Fused MNIST-ConvNet training step for sm_100a: forward + loss + backward of the tutorial's `Net`
// (train_dist.py:53-71, nll_loss train_dist.py:120, backward :122) in ONE kernel.
//
// The reference runs ~35 library/ATen kernels per step (cuDNN convs, pools, relus, dropouts, cuBLAS
// linears, log_softmax, nll, and their backward twins) on [B,...] tensors that round-trip through
// HBM.  The network is per-sample independent and tiny (21,840 parameters, < 8 KB of activations per
// sample), so here one CTA carries a sample through the whole network and back with everything in
// shared memory / registers; weight gradients are accumulated in shared memory and flushed once per
// CTA with vectorised `red.global.add.v4.f32` into the flat gradient bucket -- which is the symmetric
// buffer the fused all-reduce + SGD kernel (sgd.cu) reads over NVSwitch.  HBM traffic per step is the
// input batch + one pass over the parameters; launches per step: 1 (+1 for all-reduce/SGD).
//
// Flat parameter layout (fp32, every tensor padded to 4 elements so all flushes are 16-byte vectors):
//   conv1.w 0 | conv1.b 252 | conv2.w 264 | conv2.b 5264 | fc1.w 5284 | fc1.b 21284 | fc2.w 21336 |
//   fc2.b 21836 | total 21848
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "tc_common.cuh"
#include "convnet_args.cuh"

namespace cn {

constexpr int T = 512;              // 16 warps per sample: the phases are latency-bound, so more warps per CTA

// The conv2 working set exists in two flavours that share one union:
//   SIMT path : fp32 weights in two layouts + split-K partial sums + the zero-padded conv2-output gradient
//   TC path   : bf16 UMMA operand tiles (K-major, 128B swizzle) for the two tcgen05 GEMMs of conv2
//               forward  D[64 pos x 32 co]   = im2col(p1)[64 x 256] * W2[32 x 256]^T
//               dgrad    D[64 pos x 256 k']  = dC[64 x 32 co]       * W2^T[256 x 32]^T      (k' = (ci/5)*128 + (ci%5)*25 + tap)
//               wgrad    D[64 co  x 256 k ]  = dC^T[64 x 64 pos]    * im2col(p1)^T[256 x 64]^T
struct SimtBufs {
  float w2f[250 * 20];      // [ci][ky][kx][co]          forward: 4 output channels per float4
  float w2b[500 * 16];      // [co][ky][kx][half][8]     backward-data: 5 input channels per (half)
  float part[6 * 1440];     // conv2 partial sums [5][20][64] (5*1280 used) / dgrad partials [6][10][144]
  float dc2pad[DC_SIZE];      // conv2-output gradient, zero padded [20][16][16]
};
struct TcBufs {
  unsigned char Bw[4 * 4096];   // W2 as B operand (fwd): 4 K-blocks x [32 rows x 128 B]
  unsigned char Bt[32768];      // W2^T as B operand (dgrad): [256 rows x 128 B] (K = co, 32 used)
  unsigned char A[4 * 8192];    // im2col(p1) as A operand: 4 K-blocks x [64 rows x 128 B]; later dA staging [64][128] fp32
  unsigned char Ad[8192];       // dC as A operand (dgrad) [64 rows x 128 B]; fwd: conv2 output staging [64][32] fp32
  unsigned char Adt[8192];      // dC^T as A operand (wgrad) [64 rows (co) x 64 positions]
};
union Scratch {
  SimtBufs simt;
  TcBufs tc;
};

struct __align__(1024) Smem {
  Scratch u;                // first member: 1024-byte aligned (UMMA SWIZZLE_128B tiles)
  float w1[252];
  float b1[12];
  float b2[20];
  float w4[500];
  float b4[12];
  float x[784];
  float p1[P1_SIZE];        // relu(pool(conv1))  [10][12][12], padded strides (see convnet_args.cuh)
  float p2[320];            // relu(pool(drop(conv2)))  [20][4][4]
  float g2[320];            // gradient at the pooled conv2 argmax
  float2 g1[1440];          // (gradient at the pooled conv1 argmax, input offset of that position as int bits)
  float h[52];              // fc1 activation after relu+dropout
  float hm[52];             // fc1 backward mask (relu' * dropout scale)
  float dh[52];
  float dlog[12];
  float m2[20];             // dropout2d channel scale
  float rnd[72];            // uniforms: [0,20) dropout2d, [20,70) dropout
  float g[NPAR];            // per-CTA gradient accumulators
  unsigned long long mma_bar;   // mbarrier: tcgen05.commit -> "accumulator ready"
  unsigned int tmem_slot;
  int work_ctr;             // dynamic work distribution inside a phase (warp-granular)
  short koff[256];          // im2col LUT: k=(ci,ky,kx) -> offset inside p1, -1 for the K padding
  unsigned char a1[1440];   // conv1 pool argmax (0..3)
  unsigned char a2[320];    // conv2 pool argmax (0..3)
  float loss_local;
  int correct_local;
};

template <bool TC>
__global__ void __launch_bounds__(T, 1) convnet_step_kernel(Args a) {
  // the kernel has no static shared memory, so the dynamic window starts at offset 0 of the CTA's (1024-byte aligned)
  // shared space: addresses stay compile-time constants (a run-time round-up costs an extra add on every access)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  if (TC && (tc::smem_u32(smem_raw) & 1023u) != 0u) __trap();   // UMMA SWIZZLE_128B tiles need 1024-byte alignment
  uint32_t mma_phase = 0;            // parity of the next "accumulator ready" wait (uniform across the CTA)
  uint32_t tmem = 0;
  const float* __restrict__ P = a.params;

  // ---------------------------------------------------------------- P0: stage weights, zero accumulators
  b2::pdl_launch_dependents();       // the all-reduce/SGD kernel may pre-launch; it parks in its own pdl_wait
  if (a.backward) {                  // everything that does not depend on the previous kernel happens before pdl_wait
    float4* g4 = reinterpret_cast<float4*>(s.g);
    for (int i = tid; i < NPAR / 4; i += T) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  b2::pdl_wait();                    // parameters / step counter written by the previous all-reduce+SGD kernel
  if (tid == T - 1) wait_input(a);   // (executor path) the H2D copy of this step's batch; the barrier that ends staging publishes it
  {
    // all global loads are issued before their first use (one L2 round trip instead of a dependent chain)
    const bool fast = (a.aux != nullptr) && !TC;   // conv2.weight already in both smem layouts (written by sgd.cu)
    float4 fa[3], fb[4];            // pre-arranged conv2.weight: every load is in flight before the first store
    if (fast) {
      const float4* __restrict__ af = reinterpret_cast<const float4*>(a.aux + AUX_W2F);
      const float4* __restrict__ ab = reinterpret_cast<const float4*>(a.aux + AUX_W2B);
#pragma unroll
      for (int k = 0; k < 3; ++k) fa[k] = (tid + k * T < 1250) ? __ldg(af + tid + k * T) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) fb[k] = (tid + k * T < 2000) ? __ldg(ab + tid + k * T) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4* __restrict__ P4w2 = reinterpret_cast<const float4*>(P + W2);   // 1250 float4, 16B aligned
    float4 v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i4 = tid + k * T;
      v[k] = (!fast && i4 < 1250) ? __ldg(P4w2 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float w1v = tid < 250 ? __ldg(P + W1 + tid) : 0.f;
    const float w4v = tid < 500 ? __ldg(P + W4 + tid) : 0.f;
    const float bv = tid < 10 ? __ldg(P + B1 + tid) : (tid < 20 ? __ldg(P + B4 + tid - 10) : (tid < 40 ? __ldg(P + B2 + tid - 20) : 0.f));
    if (tid < 250) s.w1[tid] = w1v;
    if (tid < 500) s.w4[tid] = w4v;
    if (tid < 10) s.b1[tid] = bv; else if (tid < 20) s.b4[tid - 10] = bv; else if (tid < 40) s.b2[tid - 20] = bv;
    if (TC) {
      // zero the bf16 operand tiles (row / K padding must be 0), build the im2col LUT, set up mbarrier + TMEM
      uint4* z = reinterpret_cast<uint4*>(s.u.tc.Bw);
      for (int i = tid; i < (16384 + 32768) / 16; i += T) z[i] = make_uint4(0u, 0u, 0u, 0u);
      if (tid < 256) s.koff[tid] = tid < 250 ? (short)p1_idx(tid / 25, (tid % 25) / 5, tid % 5) : (short)-1;
      if (tid == 0) { tc::mbar_init(reinterpret_cast<uint64_t*>(&s.mma_bar), 1); tc::mbar_fence_init(); }
      if ((tid >> 5) == 1) tc::tmem_alloc<512>(&s.tmem_slot);
      __syncthreads();
    }
    if (fast) {
      float4* df = reinterpret_cast<float4*>(s.u.simt.w2f);
      float4* db = reinterpret_cast<float4*>(s.u.simt.w2b);
#pragma unroll
      for (int k = 0; k < 3; ++k) if (tid + k * T < 1250) df[tid + k * T] = fa[k];
#pragma unroll
      for (int k = 0; k < 4; ++k) if (tid + k * T < 2000) db[tid + k * T] = fb[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i4 = tid + k * T;
      if (!fast && i4 < 1250) {
        const float w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {             // conv2.weight [co][ci][ky][kx]
          const int i = i4 * 4 + e;
          const int co = i / 250, r = i % 250, ci = r / 25, kk = r % 25;
          if (TC) {
            const __nv_bfloat16 wb = __float2bfloat16(w[e]);
            *reinterpret_cast<__nv_bfloat16*>(s.u.tc.Bw + (r >> 6) * 4096 + tc::sw128_offset(co, r & 63)) = wb;
            const int n = (ci / 5) * 128 + (ci % 5) * 25 + kk;      // dgrad output column (ci groups on 128 boundaries)
            *reinterpret_cast<__nv_bfloat16*>(s.u.tc.Bt + tc::sw128_offset(n, co)) = wb;
          } else {
            s.u.simt.w2f[(ci * 25 + kk) * 20 + co] = w[e];
            s.u.simt.w2b[((co * 25 + kk) * 2 + ci / 5) * 8 + ci % 5] = w[e];
          }
        }
      }
    }
    }
  if (tid == 0) { s.loss_local = 0.f; s.correct_local = 0; }
  const unsigned long long step = a.step ? *a.step : 0ull;
  const float keep_scale = 1.f / (1.f - a.p_drop);
  if (TC) { tc::fence_proxy_async(); tc::fence_before(); }
  __syncthreads();
  if (TC) { tc::fence_after(); tmem = s.tmem_slot; }

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    // -------------------------------------------------------------- S0: input, RNG, clear scratch
    if (a.x_u8) {
      const uint4* xs = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.x) + (size_t)b * 784);
      if (tid < 49) {                              // 784 bytes = 49 x 16
        const uint4 q = __ldcg(xs + tid);
        const unsigned int wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 16; ++e)
          s.x[tid * 16 + e] = ((float)((wv[e >> 2] >> ((e & 3) * 8)) & 0xffu) * (1.f / 255.f) - a.mean) * a.inv_std;
      }
    } else {
      const float4* xs = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + (size_t)b * 784);
      if (tid < 196) reinterpret_cast<float4*>(s.x)[tid] = __ldcg(xs + tid);
    }
    if (tid >= 256 && tid < 274) {
      const int q = tid - 256;
      uint4 r = b2::Philox::gen(a.seed, (unsigned long long)(a.sample_base + b), step * 32ull + q);
      const float k = 2.3283064365386963e-10f;   // 2^-32
      s.rnd[q * 4 + 0] = r.x * k; s.rnd[q * 4 + 1] = r.y * k;
      s.rnd[q * 4 + 2] = r.z * k; s.rnd[q * 4 + 3] = r.w * k;
    }
    if (!TC && a.backward)
      for (int i = tid; i < DC_SIZE / 4; i += T) reinterpret_cast<float4*>(s.u.simt.dc2pad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // -------------------------------------------------------------- S1: conv1 -> maxpool2 -> relu
    for (int o = tid; o < 1440; o += T) {
      const int c = o / 144, r = o % 144, py = r / 12, px = r % 12;
      float patch[6][6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {              // 8-byte aligned pairs: 18 LDS.64 instead of 36 LDS.32
          const float2 q = *reinterpret_cast<const float2*>(&s.x[(2 * py + i) * 28 + 2 * px + 2 * j]);
          patch[i][2 * j] = q.x; patch[i][2 * j + 1] = q.y;
        }
      const float bias = s.b1[c];
      float a00 = bias, a01 = bias, a10 = bias, a11 = bias;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const float w = s.w1[c * 25 + ky * 5 + kx];
          a00 = fmaf(w, patch[ky][kx], a00);
          a01 = fmaf(w, patch[ky][kx + 1], a01);
          a10 = fmaf(w, patch[ky + 1][kx], a10);
          a11 = fmaf(w, patch[ky + 1][kx + 1], a11);
        }
      float m = a00; int arg = 0;
      if (a01 > m) { m = a01; arg = 1; }
      if (a10 > m) { m = a10; arg = 2; }
      if (a11 > m) { m = a11; arg = 3; }
      s.p1[p1_idx(c, py, px)] = fmaxf(m, 0.f);
      s.a1[o] = (unsigned char)arg;
    }
    if (tid < 20)
      s.m2[tid] = a.training ? (s.rnd[tid] >= a.p_drop ? keep_scale : 0.f) : 1.f;
    __syncthreads();

    // -------------------------------------------------------------- S2: conv2 (TC: im2col + tcgen05 GEMM | SIMT: K split 5)
    if (TC) {
      // im2col(p1) -> bf16 A operand, 2048 16-byte chunks (row = output position, 8 consecutive k per chunk)
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int chunk = tid + m * T, r = chunk >> 5, c = chunk & 31;
        const int base = (r >> 3) * P1_ROW + (r & 7);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ko = s.koff[c * 8 + e];
          f[e] = ko >= 0 ? s.p1[ko + base] : 0.f;
        }
        const uint4 pk = make_uint4(b2::pack_bf16x2(f[0], f[1]), b2::pack_bf16x2(f[2], f[3]), b2::pack_bf16x2(f[4], f[5]),
                                    b2::pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(s.u.tc.A + (c >> 3) * 8192 + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 7) ^ (r & 7)) & 7) << 4)) = pk;
      }
      tc::fence_proxy_async();
      tc::fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after();
        constexpr uint32_t idesc = tc::idesc_bf16(64, 32);
        const uint32_t a0 = tc::smem_u32(s.u.tc.A), b0 = tc::smem_u32(s.u.tc.Bw);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(tmem, tc::smem_desc_sw128(a0 + kb * 8192 + k * 32), tc::smem_desc_sw128(b0 + kb * 4096 + k * 32),
                          idesc, (kb | k) != 0 ? 1u : 0u);
        tc::commit(reinterpret_cast<uint64_t*>(&s.mma_bar));
      }
      tc::mbar_wait(reinterpret_cast<uint64_t*>(&s.mma_bar), mma_phase);
      mma_phase ^= 1;
      tc::fence_after();
      if (tid < 128) {                            // UMMA_M = 64: rows 16q..16q+15 live in TMEM lanes 32q..32q+15
        const int q = tid >> 5, lane = tid & 31;
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16), r);
        tc::tmem_ld_wait();
        if (lane < 16) {
          const int row = q * 16 + lane;
          float4* dst = reinterpret_cast<float4*>(s.u.tc.Ad) + row * 8;      // conv2 output staging [64][32] fp32
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            dst[c4 ^ (row & 7)] = make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]),
                                             __uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
        }
      }
      tc::fence_before();
      __syncthreads();
    }
    if (!TC) {
    if (tid < 400) {
      const int cell = tid & 15, cg = (tid >> 4) % 5, ks = tid / 80;
      const int py = cell >> 2, px = cell & 3;
      const int ci0 = 2 * ks, ci1 = 2 * ks + 2;
      float acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;
      for (int ci = ci0; ci < ci1; ++ci) {
        float patch[6][6];
        const float* src = &s.p1[p1_idx(ci, 2 * py, 2 * px)];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) patch[i][j] = src[i * P1_ROW + j];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float4 w = *reinterpret_cast<const float4*>(&s.u.simt.w2f[((ci * 5 + ky) * 5 + kx) * 20 + cg * 4]);
            const float i00 = patch[ky][kx], i01 = patch[ky][kx + 1], i10 = patch[ky + 1][kx], i11 = patch[ky + 1][kx + 1];
            acc[0][0] = fmaf(w.x, i00, acc[0][0]); acc[0][1] = fmaf(w.y, i00, acc[0][1]);
            acc[0][2] = fmaf(w.z, i00, acc[0][2]); acc[0][3] = fmaf(w.w, i00, acc[0][3]);
            acc[1][0] = fmaf(w.x, i01, acc[1][0]); acc[1][1] = fmaf(w.y, i01, acc[1][1]);
            acc[1][2] = fmaf(w.z, i01, acc[1][2]); acc[1][3] = fmaf(w.w, i01, acc[1][3]);
            acc[2][0] = fmaf(w.x, i10, acc[2][0]); acc[2][1] = fmaf(w.y, i10, acc[2][1]);
            acc[2][2] = fmaf(w.z, i10, acc[2][2]); acc[2][3] = fmaf(w.w, i10, acc[2][3]);
            acc[3][0] = fmaf(w.x, i11, acc[3][0]); acc[3][1] = fmaf(w.y, i11, acc[3][1]);
            acc[3][2] = fmaf(w.z, i11, acc[3][2]); acc[3][3] = fmaf(w.w, i11, acc[3][3]);
          }
      }
      // part[ks][co][cell][pos]  (pos = dy*2+dx inside the pool window)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(&s.u.simt.part[ks * 1440 + ((cg * 4 + c) * 16 + cell) * 4]) =
            make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    }
    __syncthreads();
    }

    // -------------------------------------------------------------- S2b: +bias, dropout2d, maxpool2, relu
    for (int o = tid; o < 320; o += T) {
      const int co = o >> 4;
      float4 q;
      if (TC) {
        const int cell = o & 15, p00 = (2 * (cell >> 2)) * 8 + 2 * (cell & 3);
        const float* c2 = reinterpret_cast<const float*>(s.u.tc.Ad);
        const int cb = co >> 2, cl = co & 3;
        q.x = c2[(p00) * 32 + ((cb ^ ((p00) & 7)) << 2) + cl];
        q.y = c2[(p00 + 1) * 32 + ((cb ^ ((p00 + 1) & 7)) << 2) + cl];
        q.z = c2[(p00 + 8) * 32 + ((cb ^ ((p00 + 8) & 7)) << 2) + cl];
        q.w = c2[(p00 + 9) * 32 + ((cb ^ ((p00 + 9) & 7)) << 2) + cl];
      } else {
        q = *reinterpret_cast<const float4*>(&s.u.simt.part[o * 4]);
#pragma unroll
        for (int ks = 1; ks < 5; ++ks) {
          const float4 t = *reinterpret_cast<const float4*>(&s.u.simt.part[ks * 1440 + o * 4]);
          q.x += t.x; q.y += t.y; q.z += t.z; q.w += t.w;
        }
      }
      const float bias = s.b2[co], sc = s.m2[co];
      const float v0 = (q.x + bias) * sc, v1 = (q.y + bias) * sc, v2 = (q.z + bias) * sc, v3 = (q.w + bias) * sc;
      float m = v0; int arg = 0;
      if (v1 > m) { m = v1; arg = 1; }
      if (v2 > m) { m = v2; arg = 2; }
      if (v3 > m) { m = v3; arg = 3; }
      s.p2[o] = fmaxf(m, 0.f);
      s.a2[o] = (unsigned char)arg;
    }
    __syncthreads();

    // -------------------------------------------------------------- S3: fc1 + relu + dropout
    {
      const int j = tid >> 3, l8 = tid & 7;
      float sum = 0.f;
      if (j < 50) {
        const float4* wrow = reinterpret_cast<const float4*>(P + W3 + j * 320);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          const float4 w = __ldg(wrow + l8 + 8 * k);
          const float4 v = *reinterpret_cast<const float4*>(&s.p2[(l8 + 8 * k) * 4]);
          sum = fmaf(w.x, v.x, sum); sum = fmaf(w.y, v.y, sum);
          sum = fmaf(w.z, v.z, sum); sum = fmaf(w.w, v.w, sum);
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (j < 50 && l8 == 0) {
        const float pre = sum + __ldg(P + B3 + j);
        const float dm = a.training ? (s.rnd[20 + j] >= a.p_drop ? keep_scale : 0.f) : 1.f;
        s.h[j] = fmaxf(pre, 0.f) * dm;
        s.hm[j] = pre > 0.f ? dm : 0.f;
      }
    }
    __syncthreads();

    // -------------------------------------------------------------- S4: fc2 + log_softmax + nll
    if (tid < 32) {
      const long long y = __ldcg(a.target + b);
      float logit = -INFINITY;
      if (tid < 10) {
        float acc = s.b4[tid];
#pragma unroll 10
        for (int i = 0; i < 50; ++i) acc = fmaf(s.w4[tid * 50 + i], s.h[i], acc);
        logit = acc;
      }
      float mx = logit; int am = tid;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, mx, d);
        const int oi = __shfl_xor_sync(0xffffffffu, am, d);
        if (o > mx || (o == mx && oi < am)) { mx = o; am = oi; }
      }
      float e = tid < 10 ? __expf(logit - mx) : 0.f;
      float se = e;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
      const float lse = mx + __logf(se);
      if (tid < 10) {
        const float logp = logit - lse;
        if (a.out_logp) a.out_logp[(size_t)b * 10 + tid] = logp;
        s.dlog[tid] = (e / se - (tid == (int)y ? 1.f : 0.f)) * a.inv_bsz;
        if (tid == (int)y) s.loss_local += -logp;
      }
      if (tid == 0 && am == (int)y) s.correct_local += 1;
    }
    if (a.mask_out) {
      if (tid < 20) a.mask_out[(size_t)b * 70 + tid] = s.m2[tid];
      else if (tid < 70) a.mask_out[(size_t)b * 70 + tid] =
          a.training ? (s.rnd[tid] >= a.p_drop ? keep_scale : 0.f) : 1.f;
    }
    __syncthreads();
    if (!a.backward) continue;

    // -------------------------------------------------------------- S5: fc2 backward
    for (int e = tid; e < 500; e += T) s.g[W4 + e] += s.dlog[e / 50] * s.h[e % 50];
    if (tid < 10) s.g[B4 + tid] += s.dlog[tid];
    if (tid >= 64 && tid < 114) {
      const int i = tid - 64;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) d = fmaf(s.w4[k * 50 + i], s.dlog[k], d);
      s.dh[i] = d * s.hm[i];
    }
    __syncthreads();

    // -------------------------------------------------------------- S6: fc1 backward
    {
      float4* gw3 = reinterpret_cast<float4*>(&s.g[W3]);
      const float4* p24 = reinterpret_cast<const float4*>(s.p2);
      for (int e4 = tid; e4 < 4000; e4 += T) {    // fc1.weight gradient: 4 consecutive inputs per thread-iteration
        const int j = e4 / 80, i4 = e4 - j * 80;
        const float d = s.dh[j];
        const float4 pv = p24[i4];
        float4 gv = gw3[e4];
        gv.x = fmaf(d, pv.x, gv.x); gv.y = fmaf(d, pv.y, gv.y); gv.z = fmaf(d, pv.z, gv.z); gv.w = fmaf(d, pv.w, gv.w);
        gw3[e4] = gv;
      }
      if (tid >= 320 && tid < 370) s.g[B3 + tid - 320] += s.dh[tid - 320];
      if (tid == 511) s.work_ctr = 0;
      for (int o = tid; o < 320; o += T) {
        float d = 0.f;
#pragma unroll 10
        for (int jj = 0; jj < 50; ++jj) d = fmaf(__ldg(P + W3 + jj * 320 + o), s.dh[jj], d);
        const int co = o >> 4, cell = o & 15, arg = s.a2[o];
        const float gv = s.p2[o] > 0.f ? d * s.m2[co] : 0.f;
        s.g2[o] = gv;
        if (!TC) {
          const int y = 2 * (cell >> 2) + (arg >> 1), x = 2 * (cell & 3) + (arg & 1);
          s.u.simt.dc2pad[co * DC_PLANE + (y + 4) * DC_ROW + (x + 4)] = gv;
        }
      }
    }
    __syncthreads();

    // -------------------------------------------------------------- S7a: conv2 weight/bias gradient (sparse)
    // Work items are handed out 32 at a time per warp from a shared counter, so the warps that had no (or a short)
    // S7b tile start here immediately and the phase ends balanced.  item < 1000: (co, ci, ky) = 5 taps x 16 pooled
    // cells; item 1000..1019: bias gradient of channel item-1000.  (TC mode: weight items are done by tcgen05.)
    auto s7a = [&]() {
      for (;;) {
        int base = 0;
        if ((tid & 31) == 0) base = atomicAdd(&s.work_ctr, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= 1020) break;
        const int item = base + (tid & 31);
        if (item < 1000) {
          if (!TC) {
            const int co = item / 50, r = item - co * 50, ci = r / 5, ky = r - ci * 5;
            float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
            for (int cell = 0; cell < 16; ++cell) {
              const float gv = s.g2[co * 16 + cell];
              if (gv != 0.f) {
                const int arg = s.a2[co * 16 + cell];
                const int ay = 2 * (cell >> 2) + (arg >> 1), ax = 2 * (cell & 3) + (arg & 1);
                const float* src = &s.p1[p1_idx(ci, ay + ky, ax)];
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) acc[kx] = fmaf(gv, src[kx], acc[kx]);
              }
            }
            float* dst = &s.g[W2 + co * 250 + ci * 25 + ky * 5];
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) dst[kx] += acc[kx];
          }
        } else if (item < 1020) {
          const int co = item - 1000;
          float d = 0.f;
#pragma unroll
          for (int cell = 0; cell < 16; ++cell) d += s.g2[co * 16 + cell];
          s.g[B2 + co] += d;
        }
      }
    };
    // -------------------------------------------------------------- S7b/S8a: conv2 weight + data gradients on tcgen05
    if (TC) {
      // (1) dC[64 pos][co] (one non-zero per pool window and channel) as the bf16 A operand of the dgrad GEMM
      if (tid < 256) {
        const int r = tid >> 2, c8 = tid & 3, oy = r >> 3, ox = r & 7;
        const int cell = (oy >> 1) * 4 + (ox >> 1), sub = (oy & 1) * 2 + (ox & 1);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int co = c8 * 8 + e;
          f[e] = (co < 20 && s.a2[co * 16 + cell] == sub) ? s.g2[co * 16 + cell] : 0.f;
        }
        const uint4 pk = make_uint4(b2::pack_bf16x2(f[0], f[1]), b2::pack_bf16x2(f[2], f[3]), b2::pack_bf16x2(f[4], f[5]),
                                    b2::pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(s.u.tc.Ad + (r >> 3) * 1024 + (r & 7) * 128 + (((c8 ^ (r & 7)) & 7) << 4)) = pk;
      } else if (tid < 256 + 160) {
        // (2) dC^T[co][pos] as the A operand of the wgrad GEMM (rows >= 20 are never read back)
        const int q = tid - 256, co = q >> 3, oy = q & 7;              // chunk = 8 positions of output row oy
        float f[8];
#pragma unroll
        for (int ox = 0; ox < 8; ++ox) {
          const int cell = (oy >> 1) * 4 + (ox >> 1), sub = (oy & 1) * 2 + (ox & 1);
          f[ox] = s.a2[co * 16 + cell] == sub ? s.g2[co * 16 + cell] : 0.f;
        }
        const uint4 pk = make_uint4(b2::pack_bf16x2(f[0], f[1]), b2::pack_bf16x2(f[2], f[3]), b2::pack_bf16x2(f[4], f[5]),
                                    b2::pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(s.u.tc.Adt + (co >> 3) * 1024 + (co & 7) * 128 + (((oy ^ (co & 7)) & 7) << 4)) = pk;
      }
      // (3) im2col(p1)^T[k][pos] as the B operand of the wgrad GEMM: 256 rows x 8 chunks of 8 consecutive ox
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int chunk = tid + m * T, k = chunk >> 3, oy = chunk & 7;
        const int ko = s.koff[k];
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (ko >= 0) {
          const float* src = &s.p1[ko + oy * P1_ROW];
          pk = make_uint4(b2::pack_bf16x2(src[0], src[1]), b2::pack_bf16x2(src[2], src[3]), b2::pack_bf16x2(src[4], src[5]),
                          b2::pack_bf16x2(src[6], src[7]));
        }
        *reinterpret_cast<uint4*>(s.u.tc.A + (k >> 3) * 1024 + (k & 7) * 128 + (((oy ^ (k & 7)) & 7) << 4)) = pk;
      }
      tc::fence_proxy_async();
      tc::fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::fence_after();
        constexpr uint32_t idesc = tc::idesc_bf16(64, 256);
        const uint32_t ad = tc::smem_u32(s.u.tc.Ad), bt = tc::smem_u32(s.u.tc.Bt);
        const uint32_t at = tc::smem_u32(s.u.tc.Adt), bi = tc::smem_u32(s.u.tc.A);
        tc::umma_bf16(tmem, tc::smem_desc_sw128(ad), tc::smem_desc_sw128(bt), idesc, 0u);               // dgrad: K = co
        tc::umma_bf16(tmem, tc::smem_desc_sw128(ad + 32), tc::smem_desc_sw128(bt + 32), idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k)                                                                      // wgrad: K = pos
          tc::umma_bf16(tmem + 256, tc::smem_desc_sw128(at + k * 32), tc::smem_desc_sw128(bi + k * 32), idesc, k != 0 ? 1u : 0u);
        tc::commit(reinterpret_cast<uint64_t*>(&s.mma_bar));
      }
      s7a();                                                  // bias gradient (20 items) while the MMAs run
      tc::mbar_wait(reinterpret_cast<uint64_t*>(&s.mma_bar), mma_phase);
      mma_phase ^= 1;
      tc::fence_after();
      float* stage = reinterpret_cast<float*>(s.u.tc.A);      // [64 rows][128 cols] fp32, float4 index XOR-swizzled by row
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {                  // input channels 5*half .. 5*half+4  <->  TMEM columns 128*half ..
        if (tid < 128) {
          const int q = tid >> 5, lane = tid & 31, row = q * 16 + (lane & 15);
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t r[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128 + cc * 32), r);
            tc::tmem_ld_wait();
            if (lane < 16) {
              float4* dst = reinterpret_cast<float4*>(stage) + row * 32;
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4)
                dst[(cc * 8 + c4) ^ (row & 31)] = make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]),
                                                              __uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
            }
          }
        } else if (tid < 192) {
          // conv2.weight gradient: accumulator rows = co (rows 0..15 in lanes 0..15 of quadrant 0, rows 16..19 in quadrant 1)
          const int q = (tid >> 5) & 3, lane = tid & 31, co = q * 16 + lane;
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t r[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + half * 128 + cc * 32), r);
            tc::tmem_ld_wait();
            if (lane < 16 && co < 20) {
              float* dst = &s.g[W2 + co * 250 + half * 128 + cc * 32];
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (half * 128 + cc * 32 + i < 250) dst[i] += __uint_as_float(r[i]);
            }
          }
        }
        tc::fence_before();
        __syncthreads();
        // col2im gather: dp1[ci][y][x] = sum_{ky,kx} dA[(y-ky, x-kx)][ci, ky, kx], fused with relu'/pool routing of conv1
        for (int o5 = tid; o5 < 720; o5 += T) {
          const int cil = o5 / 144, rem = o5 - cil * 144, y = rem / 12, x = rem - y * 12;
          float d = 0.f;
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
            const int oy = y - ky;
            if ((unsigned)oy < 8u) {
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) {
                const int ox = x - kx;
                if ((unsigned)ox < 8u) {
                  const int row = oy * 8 + ox, j = cil * 25 + ky * 5 + kx;
                  d += stage[row * 128 + ((((j >> 2) ^ (row & 31)) & 31) << 2) + (j & 3)];
                }
              }
            }
          }
          const int o = (half * 5 + cil) * 144 + rem, arg = s.a1[o];
          const int off = (2 * y + (arg >> 1)) * 28 + 2 * x + (arg & 1);
          s.g1[o] = make_float2(s.p1[p1_of(o)] > 0.f ? d : 0.f, __int_as_float(off));
        }
        __syncthreads();
      }
    }
    if (!TC) {
    if (tid < 432) {
      const int tile = tid % 36, half = (tid / 36) & 1, ks = tid / 72;
      const int y0 = 2 * (tile / 6), x0 = 2 * (tile % 6);
      const int co0 = ks < 2 ? 4 * ks : 8 + 3 * (ks - 2), co1 = ks < 2 ? co0 + 4 : co0 + 3;   // 4,4,3,3,3,3
      float acc[4][5];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[p][c] = 0.f;
      for (int co = co0; co < co1; ++co) {
        if (s.m2[co] == 0.f) continue;            // channel dropped by Dropout2d: gradient plane is zero
        float patch[6][6];
        const float2* src = reinterpret_cast<const float2*>(&s.u.simt.dc2pad[co * DC_PLANE + y0 * DC_ROW + x0]);   // even offsets
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float2 q = src[i * (DC_ROW / 2) + j];
            patch[i][2 * j] = q.x; patch[i][2 * j + 1] = q.y;
          }
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float* wp = &s.u.simt.w2b[((co * 25 + ky * 5 + kx) * 2 + half) * 8];
            const float4 w = *reinterpret_cast<const float4*>(wp);
            const float w4 = wp[4];
            const float d00 = patch[4 - ky][4 - kx], d01 = patch[4 - ky][5 - kx];
            const float d10 = patch[5 - ky][4 - kx], d11 = patch[5 - ky][5 - kx];
            acc[0][0] = fmaf(w.x, d00, acc[0][0]); acc[0][1] = fmaf(w.y, d00, acc[0][1]);
            acc[0][2] = fmaf(w.z, d00, acc[0][2]); acc[0][3] = fmaf(w.w, d00, acc[0][3]); acc[0][4] = fmaf(w4, d00, acc[0][4]);
            acc[1][0] = fmaf(w.x, d01, acc[1][0]); acc[1][1] = fmaf(w.y, d01, acc[1][1]);
            acc[1][2] = fmaf(w.z, d01, acc[1][2]); acc[1][3] = fmaf(w.w, d01, acc[1][3]); acc[1][4] = fmaf(w4, d01, acc[1][4]);
            acc[2][0] = fmaf(w.x, d10, acc[2][0]); acc[2][1] = fmaf(w.y, d10, acc[2][1]);
            acc[2][2] = fmaf(w.z, d10, acc[2][2]); acc[2][3] = fmaf(w.w, d10, acc[2][3]); acc[2][4] = fmaf(w4, d10, acc[2][4]);
            acc[3][0] = fmaf(w.x, d11, acc[3][0]); acc[3][1] = fmaf(w.y, d11, acc[3][1]);
            acc[3][2] = fmaf(w.z, d11, acc[3][2]); acc[3][3] = fmaf(w.w, d11, acc[3][3]); acc[3][4] = fmaf(w4, d11, acc[3][4]);
          }
      }
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        float* dst = &s.u.simt.part[ks * 1440 + (half * 5 + c) * 144 + y0 * 12 + x0];
        dst[0] = acc[0][c]; dst[1] = acc[1][c]; dst[12] = acc[2][c]; dst[13] = acc[3][c];
      }
    }
    s7a();
    __syncthreads();

    // -------------------------------------------------------------- S8a: through relu+pool of conv1
    for (int o = tid; o < 1440; o += T) {
      float d = 0.f;
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) d += s.u.simt.part[ks * 1440 + o];
      const int cell = o % 144, arg = s.a1[o];
      const int off = (2 * (cell / 12) + (arg >> 1)) * 28 + 2 * (cell % 12) + (arg & 1);
      s.g1[o] = make_float2(s.p1[p1_of(o)] > 0.f ? d : 0.f, __int_as_float(off));
    }
    __syncthreads();
    }

    // -------------------------------------------------------------- S8b: conv1 weight/bias gradient (sparse)
    {
      const int out = tid >> 1, half = tid & 1;              // two lanes per tap: 72 cells each
      float acc = 0.f, gsum = 0.f;
      int k = 0;
      if (tid < 500) {
        const int c = out / 25;
        k = out - c * 25;
        const int koff = (k / 5) * 28 + (k % 5);
        const float2* gp = &s.g1[c * 144 + half * 72];
#pragma unroll 8
        for (int cell = 0; cell < 72; ++cell) {
          const float2 q = gp[cell];
          gsum += q.x;                                       // bias gradient rides along (used by the k == 0 lanes)
          acc = fmaf(q.x, s.x[__float_as_int(q.y) + koff], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);           // executed by every lane (no divergence at the shuffle)
      gsum += __shfl_xor_sync(0xffffffffu, gsum, 1);
      if (tid < 500 && half == 0) {
        s.g[W1 + out] += acc;
        if (k == 0) s.g[B1 + out / 25] += gsum;
      }
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ flush
  if (a.backward && blockIdx.x < a.B) {
    if (a.det_partials != nullptr) {        // deterministic mode: a private slot per CTA, summed in CTA order afterwards
      float4* slot = reinterpret_cast<float4*>(a.det_partials + (size_t)blockIdx.x * DET_STRIDE);
      for (int v = tid; v < NPAR / 4; v += T) slot[v] = *reinterpret_cast<const float4*>(&s.g[v * 4]);
    } else {
      float* gdst = a.grads + (size_t)(step & 1ull) * (size_t)a.grad_stride;   // double-buffered buckets: see sgd.cu
      for (int v = tid; v < NPAR / 4; v += T) {
        const float4 q = *reinterpret_cast<const float4*>(&s.g[v * 4]);
        red_add_v4(gdst + v * 4, q.x, q.y, q.z, q.w);
      }
    }
  }
  if (tid == 0 && a.loss_acc != nullptr && blockIdx.x < a.B) {
    if (a.det_partials != nullptr && a.backward) {       // deterministic mode: the slot's padding carries this CTA's loss terms
      a.det_partials[(size_t)blockIdx.x * DET_STRIDE + NPAR] = s.loss_local * a.inv_bsz;
      a.det_partials[(size_t)blockIdx.x * DET_STRIDE + NPAR + 1] = (float)s.correct_local;
    } else {
      atomicAdd(a.loss_acc, s.loss_local * a.inv_bsz);
      atomicAdd(a.loss_acc + 1, (float)s.correct_local);
    }
  }
  // ------------------------------------------------------------------ fused tail: gradient exchange + SGD in this kernel
  if (a.tail.enabled && a.backward) b2::fused_tail(a.tail, step, (int)gridDim.x, (int)blockIdx.x);   // grid <= B: every CTA flushed
  if (TC) {
    tc::fence_before();
    __syncthreads();
    if ((tid >> 5) == 1) { tc::fence_after(); tc::tmem_dealloc<512>(tmem); }
  }
}

}  // namespace cn

extern "C" {

size_t b2_convnet_smem_bytes() { return sizeof(cn::Smem) + 1024; }
int b2_convnet_npar() { return cn::NPAR; }

static int g_convnet_tc = -1;      // -1: read B200DIST_CONVNET_TC on first use
void b2_convnet_set_tc(int on) { g_convnet_tc = on ? 1 : 0; }
int b2_convnet_get_tc() {
  if (g_convnet_tc < 0) {
    const char* e = getenv("B200DIST_CONVNET_TC");
    g_convnet_tc = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_convnet_tc;
}

int b2_convnet_step_launch(const float* params, float* grads, const void* x, int x_u8, const long long* target,
                           float* loss_acc, float* out_logp, float* mask_out, const unsigned long long* step,
                           unsigned long long seed, long long sample_base, int B, int training, int backward,
                           float inv_bsz, float p_drop, int max_ctas, long long grad_stride, const float* aux,
                           const cn::FusedTailHost* tail, float* det_partials, const unsigned int* in_flag, unsigned int in_gen,
                           cudaStream_t stream) {
  static bool configured = false;
  const size_t smem = sizeof(cn::Smem) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cn::convnet_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(cn::convnet_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cn::Args a;
  a.params = params; a.grads = grads; a.x = x; a.target = target; a.loss_acc = loss_acc; a.out_logp = out_logp;
  a.mask_out = mask_out; a.step = step; a.seed = seed; a.sample_base = sample_base; a.B = B; a.x_u8 = x_u8;
  a.training = training; a.backward = backward; a.inv_bsz = inv_bsz; a.p_drop = p_drop;
  a.mean = 0.1307f; a.inv_std = 1.f / 0.3081f; a.grad_stride = grad_stride; a.aux = aux;
  cn::fill_tail(a.tail, backward ? tail : nullptr, grad_stride);
  a.det_partials = backward ? det_partials : nullptr;
  a.in_flag = in_flag; a.in_gen = in_gen;
  int grid = B;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (grid < 1) grid = 1;
  if (a.tail.enabled) {       // the tail's grid-wide check-in spins: every CTA must be resident (one CTA per SM)
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (grid > sms) grid = sms;                 // CTAs loop over samples (b += gridDim.x)
  }
  static const int pdl = [] { const char* e = getenv("B200DIST_PDL"); return (e == nullptr || e[0] != '0') ? 1 : 0; }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)cn::T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = b2_convnet_get_tc() ? cudaLaunchKernelEx(&cfg, cn::convnet_step_kernel<true>, a)
                                      : cudaLaunchKernelEx(&cfg, cn::convnet_step_kernel<false>, a);
  return (int)e;
}

}  // extern "C"
