// fp32 CUDA-core kernels of the sparse (CSR) exact path.  Internal layout is NODE-MAJOR: a graph signal
// with C channels is stored [R, N, C] (R = samples or samples*time) so that a neighbour gather reads
// C contiguous floats (one coalesced, vectorisable segment per neighbour).
//
// Reference op sites (Utils/graphML.py): shift z@S :123, tap contraction :134-135, bias :138-139,
// cell update :2402-2423, time-gate MLP :2362-2374, node-gate head :2383-2399, attention :586-625.
#pragma once
#include "common.cuh"

namespace gcrnn {
namespace k {

constexpr int MAX_SLABS = 24;
struct Slabs { const float* p[MAX_SLABS]; };
struct SlabsMut { float* p[MAX_SLABS]; };

static inline int grid1d(long long total, int block, int max_blocks = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ---- batched 2-D transpose: out[r][b][a] = in[r][a][b] (+ add[r][b][a]) --------------------------------
// r = r1*R2 + r2 with independent in/out strides so that [T,B,..] <-> [B,T,..] re-orderings fuse in.
constexpr int TRANSPOSE_TY = 8;     // 32 x 32 tiles per block along A: a block moves 32 KB instead of 4 KB (millions of 4 KB blocks
                                    // are bound by block scheduling, not by HBM: 1.7 TB/s measured at cfg5's H)
__global__ void transpose_k(const float* __restrict__ in, const float* add, float* out, int A, int Bd,
                            long long R1, long long R2, long long is1, long long is2, long long os1, long long os2) {
  __shared__ float tile[32][33];
  const long long R = R1 * R2;
  for (long long r = blockIdx.z; r < R; r += gridDim.z) {
    const long long r1 = r / R2, r2 = r % R2;
    const float* ip = in + r1 * is1 + r2 * is2;
    const long long oo = r1 * os1 + r2 * os2;
    const int b0 = blockIdx.x * 32;
    for (int ty = 0; ty < TRANSPOSE_TY; ++ty) {
      const int a0 = (blockIdx.y * TRANSPOSE_TY + ty) * 32;
      if (a0 >= A) break;                         // uniform across the block
      for (int i = threadIdx.y; i < 32; i += 8) {
        int a = a0 + i, b = b0 + threadIdx.x;
        if (a < A && b < Bd) tile[i][threadIdx.x] = ip[(long long)a * Bd + b];
      }
      __syncthreads();
      for (int i = threadIdx.y; i < 32; i += 8) {
        int b = b0 + i, a = a0 + threadIdx.x;
        if (a < A && b < Bd) {
          long long o = oo + (long long)b * A + a;
          float v = tile[threadIdx.x][i];
          if (add) v += add[o];
          out[o] = v;
        }
      }
      __syncthreads();
    }
  }
}

// node-major -> feature-major with 32 channels: in[r][a][32] -> out[r][32][a].  A block moves a 128 x 32 tile with 16 KB of
// float4 loads in flight (the generic 32 x 32 tiles leave HBM latency-bound at ~3 TB/s); same (r1, r2) stride interface.
// rowmap != nullptr: output column a comes from input row rowmap[a] (library-owned node renumbering, graph.cu: rowmap = old -> new);
// a row is one full 128-byte line, so the renumbering costs gathered lines on the read side and nothing on the write side.
__global__ void __launch_bounds__(256) transpose_c32_k(const float* __restrict__ in, float* __restrict__ out, int A,
                                                       long long R1, long long R2, long long is1, long long is2, long long os1, long long os2,
                                                       const int* __restrict__ rowmap) {
  __shared__ float tile[128][33];
  const long long R = R1 * R2;
  const int tiles_a = (A + 127) / 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long w = blockIdx.x; w < R * tiles_a; w += gridDim.x) {
    const long long r = w / tiles_a; const int a0 = (int)(w - r * tiles_a) * 128;
    const long long r1 = r / R2, r2 = r % R2;
    const float* ib = in + r1 * is1 + r2 * is2;
    float* op = out + r1 * os1 + r2 * os2 + a0;
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + 256 * q, a = a0 + (idx >> 3);
      if (a < A) {
        const long long row = rowmap ? __ldg(rowmap + a) : a;
        v[q] = __ldg(reinterpret_cast<const float4*>(ib + row * 32) + (idx & 7));
      } else {
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + 256 * q, node = idx >> 3, c = (idx & 7) * 4;
      tile[node][c] = v[q].x; tile[node][c + 1] = v[q].y; tile[node][c + 2] = v[q].z; tile[node][c + 3] = v[q].w;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = warp + 8 * j;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int node = pass * 32 + lane;
        if (a0 + node < A) op[(long long)f * A + node] = tile[node][f];
      }
    }
    __syncthreads();
  }
}

// feature-major -> node-major with 32 channels through a row map: in[r][32][a] -> out[r][rowmap[a]][32] (coalesced reads along a,
// one full 128-byte line written per node)
__global__ void __launch_bounds__(256) scatter_rows_c32_k(const float* __restrict__ in, float* __restrict__ out, int A,
                                                          long long R1, long long R2, long long is1, long long is2, long long os1, long long os2,
                                                          const int* __restrict__ rowmap) {
  __shared__ float tile[128][33];
  const long long R = R1 * R2;
  const int tiles_a = (A + 127) / 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long w = blockIdx.x; w < R * tiles_a; w += gridDim.x) {
    const long long r = w / tiles_a; const int a0 = (int)(w - r * tiles_a) * 128;
    const long long r1 = r / R2, r2 = r % R2;
    const float* ip = in + r1 * is1 + r2 * is2 + a0;
    float* ob = out + r1 * os1 + r2 * os2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = warp + 8 * j;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int node = pass * 32 + lane;
        tile[node][f] = (a0 + node < A) ? __ldg(ip + (long long)f * A + node) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = threadIdx.x + 256 * q, node = idx >> 3, c = (idx & 7) * 4, a = a0 + node;
      if (a < A) {
        const long long row = __ldg(rowmap + a);
        *(reinterpret_cast<float4*>(ob + row * 32) + (idx & 7)) = make_float4(tile[node][c], tile[node][c + 1], tile[node][c + 2], tile[node][c + 3]);
      }
    }
    __syncthreads();
  }
}

// reference layout <-> node-major through a node renumbering (perm: new -> old), same (r1, r2) stride interface:
//   IN : out[r][n][c] = in[r][c][perm[n]]        OUT: out[r][c][perm[n]] = in[r][n][c]
template <bool IN>
__global__ void permute_nodes_k(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ perm, int C, int N,
                                long long R1, long long R2, long long is1, long long is2, long long os1, long long os2) {
  const long long total = R1 * R2 * N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long rn = i / C;
    const int n = (int)(rn % N);
    const long long r = rn / N, r1 = r / R2, r2 = r % R2;
    const long long ref = (long long)c * N + __ldg(perm + n), nm = (long long)n * C + c;
    if (IN) out[r1 * os1 + r2 * os2 + nm] = in[r1 * is1 + r2 * is2 + ref];
    else    out[r1 * os1 + r2 * os2 + ref] = in[r1 * is1 + r2 * is2 + nm];
  }
}

// degenerate transposes (A == 1 or Bd == 1): a strided batch of contiguous copies, out[oo + i] = in[ip + i], i < L
__global__ void strided_copy_k(const float* __restrict__ in, const float* add, float* out, long long L,
                               long long R1, long long R2, long long is1, long long is2, long long os1, long long os2) {
  const long long R = R1 * R2;
  for (long long r = blockIdx.y; r < R; r += gridDim.y) {
    const long long r1 = r / R2, r2 = r % R2;
    const float* ip = in + r1 * is1 + r2 * is2;
    const long long oo = r1 * os1 + r2 * os2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
      float v = ip[i];
      if (add) v += add[oo + i];
      out[oo + i] = v;
    }
  }
}

// ---- sparse shift (SpMM), node-major: out[r,d,:] = add[r,d,:] + sum_p val[p] * in[r, idx[p], :] --------
template <int VEC>
__global__ void spmm_k(const int* __restrict__ ptr, const int* __restrict__ idx, const float* __restrict__ val,
                       const float* __restrict__ in, const float* __restrict__ add, float* __restrict__ out,
                       int N, int C, long long R) {
  const int CV = C / VEC;
  const long long total = R * N * CV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    const long long rn = i / CV;
    const int d = (int)(rn % N);
    const long long r = rn / N;
    const float* base = in + r * N * C + cv * VEC;
    const long long o = rn * C + cv * VEC;
    const int p0 = ptr[d], p1 = ptr[d + 1];
    if (VEC == 4) {
      float4 acc = add ? *reinterpret_cast<const float4*>(add + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = p0; p < p1; ++p) {
        const float v = __ldg(val + p);
        const float4 s = __ldg(reinterpret_cast<const float4*>(base + (long long)__ldg(idx + p) * C));
        acc.x = fmaf(v, s.x, acc.x); acc.y = fmaf(v, s.y, acc.y); acc.z = fmaf(v, s.z, acc.z); acc.w = fmaf(v, s.w, acc.w);
      }
      *reinterpret_cast<float4*>(out + o) = acc;
    } else {
      float acc = add ? add[o] : 0.f;
      for (int p = p0; p < p1; ++p) acc = fmaf(__ldg(val + p), __ldg(base + (long long)__ldg(idx + p) * C), acc);
      out[o] = acc;
    }
  }
}

// ---- tap contraction: y[rn,f] = epi( sum_{s,g} W[f,s,g] z_s[rn,g] + bias_scale*bias[f] ) ----------------
// EPI 0: plain.  EPI 1: tanh(acc + addb[(r % RB), n, f])  (ungated sub-cell state, T-invariant term addb).
template <int EPI>
__global__ void contract_fwd_k(Slabs z, const float* __restrict__ W, const float* __restrict__ bias, float bias_scale,
                               const float* __restrict__ addb, long long RB, float* __restrict__ y,
                               long long R, int N, int F, int S, int G) {
  extern __shared__ float Wt[];  // [S*G][F]
  const int SG = S * G;
  for (int i = threadIdx.x; i < F * SG; i += blockDim.x) { int f = i / SG, sg = i % SG; Wt[sg * F + f] = W[i]; }
  __syncthreads();
  const long long total = R * N * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long rn = i / F;
    float acc = bias ? bias[f] * bias_scale : 0.f;
    for (int s = 0; s < S; ++s) {
      const float* zp = z.p[s] + rn * G;
      const float* wp = Wt + (s * G) * F + f;
      for (int g = 0; g < G; ++g) acc = fmaf(wp[g * F], __ldg(zp + g), acc);
    }
    if (EPI == 1) {
      const long long r = rn / N; const int n = (int)(rn % N);
      acc = tanhf(acc + addb[((r % RB) * N + n) * F + f]);
    }
    y[i] = acc;
  }
}

// data gradient: dz_s[rn,g] (+)= sum_f W[f,s,g] dy[rn,f]
__global__ void contract_bwd_data_k(SlabsMut dz, const float* __restrict__ W, const float* __restrict__ dy,
                                    long long RN, int F, int S, int G, int accumulate) {
  extern __shared__ float Ws[];  // [F][S*G]
  const int SG = S * G;
  for (int i = threadIdx.x; i < F * SG; i += blockDim.x) Ws[i] = W[i];
  __syncthreads();
  const long long total = RN * SG;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int sg = (int)(i % SG);
    const long long rn = i / SG;
    const float* d = dy + rn * F;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc = fmaf(Ws[f * SG + sg], __ldg(d + f), acc);
    float* o = dz.p[sg / G] + rn * G + (sg % G);
    *o = accumulate ? *o + acc : acc;
  }
}

// weight gradient: dW[f,s,g] += sum_rn dy[rn,f] z_s[rn,g]   (persistent blocks, register accumulators)
constexpr int WG_ITEMS = 32;   // (r,n) items staged per iteration
constexpr int WG_OPT = 8;      // outputs per thread
__global__ void contract_wgrad_k(Slabs z, const float* __restrict__ dy, float* dW, long long RN, int F, int S, int G) {
  extern __shared__ float sm[];
  const int SG = S * G;
  float* dys = sm;                    // [WG_ITEMS][F]
  float* zs = sm + WG_ITEMS * F;      // [WG_ITEMS][SG]
  const int FSG = F * SG;
  const int o_base = blockIdx.y * blockDim.x * WG_OPT;
  float acc[WG_OPT];
#pragma unroll
  for (int j = 0; j < WG_OPT; ++j) acc[j] = 0.f;
  for (long long c = blockIdx.x; c * WG_ITEMS < RN; c += gridDim.x) {
    const long long rn0 = c * WG_ITEMS;
    for (int i = threadIdx.x; i < WG_ITEMS * F; i += blockDim.x) {
      long long rn = rn0 + i / F;
      dys[i] = rn < RN ? dy[rn * F + (i % F)] : 0.f;
    }
    for (int i = threadIdx.x; i < WG_ITEMS * SG; i += blockDim.x) {
      int it = i / SG, sg = i % SG;
      long long rn = rn0 + it;
      zs[i] = rn < RN ? z.p[sg / G][rn * G + (sg % G)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < WG_OPT; ++j) {
      const int o = o_base + j * blockDim.x + threadIdx.x;
      if (o < FSG) {
        const int f = o / SG, sg = o % SG;
        float a = acc[j];
#pragma unroll 8
        for (int it = 0; it < WG_ITEMS; ++it) a = fmaf(dys[it * F + f], zs[it * SG + sg], a);
        acc[j] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < WG_OPT; ++j) {
    const int o = o_base + j * blockDim.x + threadIdx.x;
    if (o < FSG && acc[j] != 0.f) atomicAdd(dW + o, acc[j]);
  }
}

// column sums: out[c] += scale * sum_r in[r, c]
__global__ void colsum_k(const float* __restrict__ in, float* out, long long rows, int C, float scale) {
  constexpr int ROWS = 64;
  for (long long c0 = blockIdx.x; c0 * ROWS < rows; c0 += gridDim.x) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
      const long long r1 = min(rows, (c0 + 1) * ROWS);
      for (long long r = c0 * ROWS; r < r1; ++r) a += in[r * C + c];
      atomicAdd(out + c, a * scale);
    }
  }
}

// ---- warp / block reductions --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  return v;  // valid in warp 0
}

// ---- cell update: h = tanh(gi*qi*ua + gf*qf*ur)   (graphML.py:2402-2423) --------------------------------
__global__ void combine_fwd_k(const float* __restrict__ ua, const float* __restrict__ ur,
                              const float* __restrict__ gi, const float* __restrict__ gf,
                              const float* __restrict__ qi, const float* __restrict__ qf,
                              float* __restrict__ h, long long B, int N, int F) {
  const long long total = B * N * F;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long bn = i / F, b = bn / N;
    float wi = gi ? gi[b] : 1.f, wf = gf ? gf[b] : 1.f;
    if (qi) { wi *= qi[bn]; wf *= qf[bn]; }
    h[i] = tanhf(fmaf(wi, ua[i], wf * ur[i]));
  }
}

// backward of the update for one step.  One thread per (b, n), block-level reduction for the scalar gates.
__global__ void combine_bwd_k(const float* __restrict__ dh, const float* __restrict__ h,
                              const float* __restrict__ ua, const float* __restrict__ ur,
                              const float* __restrict__ gi, const float* __restrict__ gf,
                              const float* __restrict__ qi, const float* __restrict__ qf,
                              float* __restrict__ dua, float* __restrict__ dur,
                              float* dgi, float* dgf, float* __restrict__ dqi, float* __restrict__ dqf,
                              long long B, int N, int F) {
  __shared__ float red[32];
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const float vgi = gi ? gi[b] : 1.f, vgf = gf ? gf[b] : 1.f;
    float si_tot = 0.f, sf_tot = 0.f;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
      const long long bn = b * N + n;
      const float vqi = qi ? qi[bn] : 1.f, vqf = qi ? qf[bn] : 1.f;
      const float wi = vgi * vqi, wf = vgf * vqf;
      float sa = 0.f, sr = 0.f;
      for (int f = 0; f < F; ++f) {
        const long long i = bn * F + f;
        const float hv = h[i];
        const float dp = dh[i] * (1.f - hv * hv);
        sa = fmaf(dp, ua[i], sa);
        sr = fmaf(dp, ur[i], sr);
        dua[i] = dp * wi;
        dur[i] = dp * wf;
      }
      if (dqi) { dqi[bn] = vgi * sa; dqf[bn] = vgf * sr; }
      si_tot += vqi * sa; sf_tot += vqf * sr;
    }
    if (dgi) {
      float a = block_sum(si_tot, red);
      if (threadIdx.x == 0) atomicAdd(dgi + b, a);
      float c = block_sum(sf_tot, red);
      if (threadIdx.x == 0) atomicAdd(dgf + b, c);
    }
  }
}

// ---- time gate: g[r] = sigmoid(sum_i Wg[i] u[r,i] + c)   (graphML.py:2364-2374) -------------------------
__global__ void gate_logit_k(const float* __restrict__ u, const float* __restrict__ Wg, const float* __restrict__ c,
                             float* __restrict__ g, long long R, long long NF) {
  __shared__ float red[32];
  for (long long r = blockIdx.x; r < R; r += gridDim.x) {
    const float* up = u + r * NF;
    float a = 0.f;
    for (long long i = threadIdx.x; i < NF; i += blockDim.x) a = fmaf(Wg[i], up[i], a);
    a = block_sum(a, red);
    if (threadIdx.x == 0) { a += c ? c[0] : 0.f; g[r] = 1.f / (1.f + expf(-a)); }
    __syncthreads();
  }
}
// dl[r] = dg[r] g[r] (1 - g[r]);  dc += sum_r dl[r]     (single block)
__global__ void gate_dlogit_k(const float* __restrict__ dg, const float* __restrict__ g, float* __restrict__ dl, float* dc, long long R) {
  __shared__ float red[32];
  float a = 0.f;
  for (long long r = threadIdx.x; r < R; r += blockDim.x) { float v = g[r]; float d = dg[r] * v * (1.f - v); dl[r] = d; a += d; }
  a = block_sum(a, red);
  if (threadIdx.x == 0 && dc) atomicAdd(dc, a);
}
// in place u -> dpre = dl[r] Wg[i] (1 - u^2);  dWg[i] += sum_r dl[r] u[r,i]
__global__ void gate_du_k(float* __restrict__ u, const float* __restrict__ Wg, const float* __restrict__ dl, float* dWg,
                          long long R, long long NF) {
  const long long rchunk = (R + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rchunk, r1 = min(R, r0 + rchunk);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < NF; i += (long long)gridDim.x * blockDim.x) {
    const float w = Wg[i];
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const float uv = u[r * NF + i], d = dl[r];
      a = fmaf(d, uv, a);
      u[r * NF + i] = d * w * (1.f - uv * uv);
    }
    atomicAdd(dWg + i, a);
  }
}
// out[b,i] = sum_t in[(t*B + b), i]
__global__ void reduce_t_k(const float* __restrict__ in, float* __restrict__ out, long long T, long long B, long long NF) {
  const long long total = B * NF;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (long long t = 0; t < T; ++t) a += in[t * total + i];
    out[i] = a;
  }
}

// ---- node gate head  GraphFilter(F -> 1)  evaluated contract-then-shift ---------------------------------
// p[s][rn] = sum_f w[s,f] u[rn,f]
__global__ void node_proj_fwd_k(const float* __restrict__ u, const float* __restrict__ w, float* __restrict__ p,
                                long long RN, int F, int S) {
  extern __shared__ float ws[];  // [S][F]
  for (int i = threadIdx.x; i < S * F; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  for (long long rn = blockIdx.x * (long long)blockDim.x + threadIdx.x; rn < RN; rn += (long long)gridDim.x * blockDim.x) {
    const float* up = u + rn * F;
    for (int s = 0; s < S; ++s) {
      float a = 0.f;
      for (int f = 0; f < F; ++f) a = fmaf(ws[s * F + f], up[f], a);
      p[(long long)s * RN + rn] = a;
    }
  }
}
// q = sigmoid(lin + c)
__global__ void sigmoid_bias_k(const float* __restrict__ lin, const float* __restrict__ c, float* __restrict__ q, long long n) {
  const float cv = c ? c[0] : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    q[i] = 1.f / (1.f + expf(-(lin[i] + cv)));
}
// dlin = dq q (1-q);  dc += sum dlin
__global__ void dsigmoid_k(const float* __restrict__ dq, const float* __restrict__ q, float* __restrict__ dlin, float* dc, long long n) {
  __shared__ float red[32];
  float a = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = q[i]; float d = dq[i] * v * (1.f - v); dlin[i] = d; a += d;
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0 && dc) atomicAdd(dc, a);
}
// in place u -> dpre = (sum_s w[s,f] dp[s][rn]) (1-u^2);  dw[s,f] += sum_rn dp[s][rn] u[rn,f]
__global__ void node_proj_bwd_k(float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ dp,
                                float* dw, long long RN, int F, int S) {
  extern __shared__ float sm[];
  float* ws = sm;            // [S][F]
  float* dws = sm + S * F;   // [S][F]
  for (int i = threadIdx.x; i < S * F; i += blockDim.x) { ws[i] = w[i]; dws[i] = 0.f; }
  __syncthreads();
  for (long long rn = blockIdx.x * (long long)blockDim.x + threadIdx.x; rn < RN; rn += (long long)gridDim.x * blockDim.x) {
    float* up = u + rn * F;
    for (int f = 0; f < F; ++f) {
      const float uv = up[f];
      float du = 0.f;
      for (int s = 0; s < S; ++s) {
        const float d = dp[(long long)s * RN + rn];
        du = fmaf(ws[s * F + f], d, du);
        atomicAdd(dws + s * F + f, d * uv);
      }
      up[f] = du * (1.f - uv * uv);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * F; i += blockDim.x) if (dws[i] != 0.f) atomicAdd(dw + i, dws[i]);
}
__global__ void add_inplace_k(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

// ---- graph attention, edge-wise over the pattern of S' = S + I  (graphML.py:586-625) ---------------------
// rc[rn] = (r = a1.Wu[rn,:], c = a2.Wu[rn,:])
__global__ void gat_scores_k(const float* __restrict__ wu, const float* __restrict__ mixer, float2* __restrict__ rc, long long RN, int F) {
  for (long long rn = blockIdx.x * (long long)blockDim.x + threadIdx.x; rn < RN; rn += (long long)gridDim.x * blockDim.x) {
    const float* p = wu + rn * F;
    float r = 0.f, c = 0.f;
    for (int f = 0; f < F; ++f) { const float v = p[f]; r = fmaf(__ldg(mixer + f), v, r); c = fmaf(__ldg(mixer + F + f), v, c); }
    rc[rn] = make_float2(r, c);
  }
}
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.2f * x; }
// alpha[r][e] = softmax over the edges e of row i of leaky(c_i + r_j)
__global__ void gat_softmax_k(const int* __restrict__ rptr, const int* __restrict__ col, const float2* __restrict__ rc,
                              float* __restrict__ alpha, long long R, int N, long long nnz) {
  const long long total = R * N;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / N; const int i = (int)(t % N);
    const float2* rcr = rc + r * N;
    const float ci = rcr[i].y;
    const int p0 = rptr[i], p1 = rptr[i + 1];
    float m = -INFINITY;
    for (int p = p0; p < p1; ++p) m = fmaxf(m, leaky(ci + rcr[col[p]].x));
    float den = 0.f;
    for (int p = p0; p < p1; ++p) den += expf(leaky(ci + rcr[col[p]].x) - m);
    const float inv = 1.f / den;
    float* ar = alpha + r * nnz;
    for (int p = p0; p < p1; ++p) ar[p] = expf(leaky(ci + rcr[col[p]].x) - m) * inv;
  }
}
// y[r,j,f] = relu( sum_{e in column j} Wu[r, row(e), f] * val[e] * alpha[r][e] )
__global__ void gat_aggregate_k(const int* __restrict__ cptr, const int* __restrict__ crow, const int* __restrict__ ceid,
                                const float* __restrict__ val, const float* __restrict__ alpha, const float* __restrict__ wu,
                                float* __restrict__ y, long long R, int N, int F, long long nnz) {
  const long long total = R * N * F;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(t % F);
    const long long rn = t / F; const int j = (int)(rn % N); const long long r = rn / N;
    const float* ar = alpha + r * nnz;
    const float* wr = wu + r * N * F + f;
    float acc = 0.f;
    for (int q = cptr[j]; q < cptr[j + 1]; ++q) {
      const int e = ceid[q];
      acc = fmaf(__ldg(val + e) * ar[e], wr[(long long)crow[q] * F], acc);
    }
    y[t] = fmaxf(acc, 0.f);
  }
}
// dyr = dy * (y > 0) ;  in place on a copy
__global__ void relu_mask_k(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// per row i: dalpha_e = val_e * <dyr[r,j(e),:], Wu[r,i,:]>; softmax + leaky backward -> ds[r][e]; dc[r,i] = sum_e ds
__global__ void gat_bwd_rows_k(const int* __restrict__ rptr, const int* __restrict__ col, const float* __restrict__ val,
                               const float2* __restrict__ rc, const float* __restrict__ alpha, const float* __restrict__ wu,
                               const float* __restrict__ dyr, float* __restrict__ ds, float* __restrict__ dc,
                               long long R, int N, int F, long long nnz) {
  const long long total = R * N;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / N; const int i = (int)(t % N);
    const float* wi = wu + t * F;
    const float* dr = dyr + r * N * F;
    const float* ar = alpha + r * nnz;
    float* dsr = ds + r * nnz;
    const float2* rcr = rc + r * N;
    const int p0 = rptr[i], p1 = rptr[i + 1];
    float dot = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float* dj = dr + (long long)col[p] * F;
      float a = 0.f;
      for (int f = 0; f < F; ++f) a = fmaf(dj[f], wi[f], a);
      a *= val[p];
      dsr[p] = a;                  // dalpha, temporarily
      dot = fmaf(ar[p], a, dot);
    }
    const float ci = rcr[i].y;
    float dci = 0.f;
    for (int p = p0; p < p1; ++p) {
      float de = ar[p] * (dsr[p] - dot);
      const float s = ci + rcr[col[p]].x;
      de *= (s > 0.f) ? 1.f : 0.2f;
      dsr[p] = de;
      dci += de;
    }
    dc[t] = dci;
  }
}
// dWu[r,i,f] = sum_{e in row i} dyr[r,j(e),f] val_e alpha_e + a2[f] dc[r,i] + a1[f] dr[r,i],
// with dr[r,i] = sum_{e in column i} ds[r][e]
__global__ void gat_bwd_dwu_k(const int* __restrict__ rptr, const int* __restrict__ col, const float* __restrict__ val,
                              const int* __restrict__ cptr, const int* __restrict__ ceid,
                              const float* __restrict__ alpha, const float* __restrict__ ds, const float* __restrict__ dc,
                              const float* __restrict__ dyr, const float* __restrict__ mixer,
                              float* __restrict__ dwu, float2* __restrict__ drc, long long R, int N, int F, long long nnz) {
  const long long total = R * N * F;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(t % F);
    const long long rn = t / F; const int i = (int)(rn % N); const long long r = rn / N;
    const float* ar = alpha + r * nnz;
    const float* dsr = ds + r * nnz;
    const float* dr = dyr + r * N * F + f;
    float acc = 0.f;
    for (int p = rptr[i]; p < rptr[i + 1]; ++p) acc = fmaf(__ldg(val + p) * ar[p], dr[(long long)col[p] * F], acc);
    float dri = 0.f;
    for (int q = cptr[i]; q < cptr[i + 1]; ++q) dri += dsr[ceid[q]];
    const float dci = dc[rn];
    if (f == 0) drc[rn] = make_float2(dri, dci);
    dwu[t] = acc + __ldg(mixer + F + f) * dci + __ldg(mixer + f) * dri;
  }
}
// dmixer[f] += sum_rn drc.x Wu[rn,f];  dmixer[F+f] += sum_rn drc.y Wu[rn,f]
__global__ void gat_bwd_mixer_k(const float2* __restrict__ drc, const float* __restrict__ wu, float* dmixer, long long RN, int F) {
  constexpr int ROWS = 64;
  for (long long c0 = blockIdx.x; c0 * ROWS < RN; c0 += gridDim.x) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
      float a = 0.f, b = 0.f;
      const long long r1 = min(RN, (c0 + 1) * ROWS);
      for (long long rn = c0 * ROWS; rn < r1; ++rn) { const float w = wu[rn * F + f]; const float2 d = drc[rn]; a = fmaf(d.x, w, a); b = fmaf(d.y, w, b); }
      atomicAdd(dmixer + f, a);
      atomicAdd(dmixer + F + f, b);
    }
  }
}

}  // namespace k
}  // namespace gcrnn
