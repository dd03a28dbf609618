// Windowed correlation kernels: Einstein MSD and Green-Kubo autocorrelation.
//
// Replaces
//   einstein_diffusion_coefficients.py:168-190, 230-244   (per-window squared_difference + sums)
//   green_kubo_self_diffusion_coefficients.py:191-199     (tfp.stats.auto_correlation per window)
//   green_kubo_ionic_conductivity.py:201-203
//   data_manager.py:309-339                               (sliding "ensembles" inside a batch)
// All windows of a batch are processed by one launch.  Trajectory rows are atom-major
// [A][T][3] fp32 (time contiguous per atom), so a (atom, time-range) tile is one contiguous,
// coalesced read that is staged in shared memory and reused for every (origin, lag) pair.
// Differences / products are formed in fp32 (inputs are exact fp32 values, so each term
// carries one rounding), short runs are summed in fp32 and folded into fp64 accumulators.
#include "mdk_common.cuh"

#include <cstdlib>

namespace mdk {

constexpr int DYN_NT = 128;  // threads per CTA: one lag per thread and pass
constexpr int DYN_RL = 4;    // lag passes per thread (lags tid + r*NT)

// ---- Einstein MSD -----------------------------------------------------------------------
// grid.x: window chunk, grid.y: atom group.  Each CTA loops over the atoms of its group,
// stages x[a][t_begin .. t_begin+len) (raw xyz interleaved: bank = 3t+d is conflict free)
// and lets thread k sweep all window origins of the chunk for its lags.
__global__ void __launch_bounds__(DYN_NT)
msd_windowed_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                    int atoms_per_cta, long long t0, int W, int ct, const int* __restrict__ tau,
                    int n_tau, int span, int Wc, double* __restrict__ msd_sum) {
  extern __shared__ float s_tile[];  // 3 * len floats
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * Wc;
  const int w1 = min(W, w0 + Wc);
  if (w0 >= w1) return;
  const int nw = w1 - w0;
  const int len = (nw - 1) * ct + span;
  const long long t_begin = t0 + (long long)w0 * ct;
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);

  for (int kb = 0; kb < n_tau; kb += DYN_NT * DYN_RL) {
    int lag3[DYN_RL];
    bool ok[DYN_RL];
    double acc64[DYN_RL];
#pragma unroll
    for (int r = 0; r < DYN_RL; ++r) {
      const int k = kb + r * DYN_NT + tid;
      ok[r] = k < n_tau;
      lag3[r] = ok[r] ? 3 * __ldg(tau + k) : 0;
      acc64[r] = 0.0;
    }
    for (long long a = a0; a < a1; ++a) {
      const float* __restrict__ src = traj + ((size_t)a * T + t_begin) * 3;
      __syncthreads();
      for (int e = tid; e < 3 * len; e += DYN_NT) s_tile[e] = __ldg(src + e);
      __syncthreads();
      float acc32[DYN_RL];
#pragma unroll
      for (int r = 0; r < DYN_RL; ++r) acc32[r] = 0.f;
      for (int w = 0; w < nw; ++w) {
        const float* __restrict__ o = s_tile + 3 * w * ct;
        const float x0 = o[0], y0 = o[1], z0 = o[2];
#pragma unroll
        for (int r = 0; r < DYN_RL; ++r) {
          const float* __restrict__ q = o + lag3[r];
          const float dx = q[0] - x0, dy = q[1] - y0, dz = q[2] - z0;
          acc32[r] = fmaf(dx, dx, acc32[r]);
          acc32[r] = fmaf(dy, dy, acc32[r]);
          acc32[r] = fmaf(dz, dz, acc32[r]);
        }
        if ((w & 15) == 15) {
#pragma unroll
          for (int r = 0; r < DYN_RL; ++r) {
            acc64[r] += (double)acc32[r];
            acc32[r] = 0.f;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < DYN_RL; ++r) acc64[r] += (double)acc32[r];
    }
#pragma unroll
    for (int r = 0; r < DYN_RL; ++r) {
      const int k = kb + r * DYN_NT + tid;
      if (ok[r]) atomicAdd(msd_sum + k, acc64[r]);
    }
  }
}

// ---- Einstein MSD, dense lags (tau = 0 .. n_lags-1, correlation_time 1) ----------------------
// A thread owns R consecutive lags and slides over window origins: the R positions x(w + lag) it
// needs at origin w are kept in a register ring, so one origin costs ONE new position load (plus
// the broadcast origin) for R updates instead of one load per update.  Threads are arranged as
// G = ceil(lags / R) lag-threads x NG = NT / G window groups, so that short lag ranges (the
// reference's default data_range is 100) still fill the CTA: group g sweeps its own slice of the
// window chunk.  The lane stride is R = 9 positions: odd, hence bank-conflict free (a skewed
// index for R = 8 was measured slower).
// Two atoms are swept together: differences and squares run on FADD2/FFMA2 for {xA,yA},
// {xB,yB} and {zA,zB} -- six packed instructions for two updates (a scalar FADD/FFMA on z costs
// the same pipe slot as a packed one); fp32 partial sums over MD_FOLD * R origins of both atoms
// are folded into fp64.
constexpr int MD_NT = 64;
constexpr int MD_FOLD = 3;   // outer iterations (of R origins) between fp64 folds
constexpr int MD2_FOLD = 2;  // same for the two-atom kernel (twice the terms per iteration)

constexpr int MD_R = 9;  // odd: the lane stride of MD_R positions is bank-conflict free

// msd_dense_kernel: one atom per sweep ({x,y} packed, z scalar) -- kept for short lag ranges, where
// window groups split the CTA and the two-atom variant below needs too many registers.
template <bool GROUPS, int R>
__global__ void __launch_bounds__(MD_NT)
msd_dense_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                 int atoms_per_cta, long long t0, int W, int n_lags, int Wc, int len_alloc,
                 double* __restrict__ msd_sum) {
  extern __shared__ __align__(16) float md_smem[];
  // layout: xy (float2 x len_alloc) | z (len_alloc floats) | origin copies for lag blocks > 0
  float2* s_xy = reinterpret_cast<float2*>(md_smem);
  float* s_z = md_smem + 2 * len_alloc;
  float* s_oxy = s_z + len_alloc;  // xy (2*Wc) then z (Wc)
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * Wc;
  const int w1 = min(W, w0 + Wc);
  if (w0 >= w1) return;
  const int nw = w1 - w0;
  const int lag_blk0 = blockIdx.z * MD_NT * R;
  const int lags_here = min(n_lags - lag_blk0, MD_NT * R);
  // GROUPS: G lag-threads x NG window groups; otherwise every thread is a lag-thread and sweeps
  // the whole window chunk (the lag range fills the CTA)
  const int G = GROUPS ? (lags_here + R - 1) / R : MD_NT;
  const int NG = GROUPS ? MD_NT / G : 1;
  const int k = GROUPS ? tid % G : tid, wg = GROUPS ? tid / G : 0;
  const bool active = wg < NG;
  const int nwg = GROUPS ? (nw + NG - 1) / NG : nw;  // origins per group
  const int ws_lo = GROUPS ? min(wg * nwg, nw) : 0;
  const int ws_hi = GROUPS ? (active ? min(nw, ws_lo + nwg) : ws_lo) : nw;
  const int len = nw + lags_here - 1;             // frames staged per atom
  const long long t_begin = t0 + w0 + lag_blk0;   // slab index 0 <-> this frame
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);
  const int kR = k * R;

  double acc64[R];
#pragma unroll
  for (int b = 0; b < R; ++b) acc64[b] = 0.0;

  for (long long a = a0; a < a1; ++a) {
    const float* __restrict__ row = traj + (size_t)a * T * 3;
    __syncthreads();
    {
      const float* __restrict__ src = row + (size_t)t_begin * 3;
      for (int e = tid; e < 3 * len; e += MD_NT) {
        const float v = __ldg(src + e);
        const int t = e / 3, d = e - 3 * t;
        if (d < 2) md_smem[2 * t + d] = v; else s_z[t] = v;
      }
      if (lag_blk0 > 0) {
        const float* __restrict__ so = row + (size_t)(t0 + w0) * 3;
        for (int e = tid; e < 3 * nw; e += MD_NT) {
          const float v = __ldg(so + e);
          const int t = e / 3, d = e - 3 * t;
          if (d < 2) s_oxy[2 * t + d] = v; else s_oxy[2 * Wc + t] = v;
        }
      }
    }
    __syncthreads();
    const float2* __restrict__ o_xy =
        lag_blk0 > 0 ? reinterpret_cast<const float2*>(s_oxy) : s_xy;
    const float* __restrict__ o_z = lag_blk0 > 0 ? s_oxy + 2 * Wc : s_z;

    float2 qxy[R];
    float qz[R];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int p = min(ws_lo + kR + b, len_alloc - 1);
      qxy[b] = s_xy[p];
      qz[b] = s_z[p];
    }
    float2 axy[R];
    float az[R];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      axy[b] = make_float2(0.f, 0.f);
      az[b] = 0.f;
    }
    int fold = 0;
    for (int w = ws_lo; w < ws_lo + nwg; w += R) {
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int ws = w + s;
        if (ws < ws_hi) {
          const float2 oxy = o_xy[ws];
          const float2 noxy = make_float2(-oxy.x, -oxy.y);
          const float noz = -o_z[ws];
#pragma unroll
          for (int b = 0; b < R; ++b) {
            const int slot = (s + b) % R;
            const float2 d = __fadd2_rn(qxy[slot], noxy);
            const float dz = qz[slot] + noz;
            axy[b] = __ffma2_rn(d, d, axy[b]);
            az[b] = fmaf(dz, dz, az[b]);
          }
          // slot s held frame ws + lag: dead now; refill with frame ws + lag + R
          const int nx = min(ws + kR + R, len_alloc - 1);
          qxy[s] = s_xy[nx];
          qz[s] = s_z[nx];
        }
      }
      if (++fold == MD_FOLD) {
        fold = 0;
#pragma unroll
        for (int b = 0; b < R; ++b) {
          acc64[b] += (double)((axy[b].x + axy[b].y) + az[b]);
          axy[b] = make_float2(0.f, 0.f);
          az[b] = 0.f;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < R; ++b) acc64[b] += (double)((axy[b].x + axy[b].y) + az[b]);
  }
  if (active) {
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int lag = lag_blk0 + kR + b;
      if (kR + b < lags_here && lag < n_lags) atomicAdd(msd_sum + lag, acc64[b]);
    }
  }
}

// msd_dense2_kernel: two atoms per sweep, z components packed (lag ranges that fill the CTA).

// R (lags per thread, odd) is picked by the launcher so that the lag range fills the 64 threads:
// 5 for up to 320 lags, 7 for up to 448, 9 beyond (several lag blocks above 576).
template <bool GROUPS, int R>
__global__ void __launch_bounds__(MD_NT)
msd_dense2_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                 int atoms_per_cta, long long t0, int W, int n_lags, int Wc, int len_alloc,
                 double* __restrict__ msd_sum) {
  extern __shared__ __align__(16) float md_smem[];
  // Two atoms (A, B) are swept together so that their z components share packed instructions:
  // layout xyA (float2 x len_alloc) | xyB | zAB (float2 {zA, zB} x len_alloc) | origin copies
  // for lag blocks > 0 (same three arrays, Wc entries each)
  float2* s_xyA = reinterpret_cast<float2*>(md_smem);
  float2* s_xyB = s_xyA + len_alloc;
  float2* s_z2 = s_xyB + len_alloc;
  float2* s_oA = s_z2 + len_alloc;
  float2* s_oB = s_oA + Wc;
  float2* s_oz = s_oB + Wc;
  const int tid = threadIdx.x;
  const int w0 = blockIdx.x * Wc;
  const int w1 = min(W, w0 + Wc);
  if (w0 >= w1) return;
  const int nw = w1 - w0;
  const int lag_blk0 = blockIdx.z * MD_NT * R;
  const int lags_here = min(n_lags - lag_blk0, MD_NT * R);
  // GROUPS: G lag-threads x NG window groups; otherwise every thread is a lag-thread and sweeps
  // the whole window chunk (the lag range fills the CTA)
  const int G = GROUPS ? (lags_here + R - 1) / R : MD_NT;
  const int NG = GROUPS ? MD_NT / G : 1;
  const int k = GROUPS ? tid % G : tid, wg = GROUPS ? tid / G : 0;
  const bool active = wg < NG;
  const int nwg = GROUPS ? (nw + NG - 1) / NG : nw;  // origins per group
  const int ws_lo = GROUPS ? min(wg * nwg, nw) : 0;
  const int ws_hi = GROUPS ? (active ? min(nw, ws_lo + nwg) : ws_lo) : nw;
  const int len = nw + lags_here - 1;             // frames staged per atom
  const long long t_begin = t0 + w0 + lag_blk0;   // slab index 0 <-> this frame
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);
  const int kR = k * R;

  double acc64[R];
#pragma unroll
  for (int b = 0; b < R; ++b) acc64[b] = 0.0;

  for (long long a = a0; a < a1; a += 2) {
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // atom B of an odd tail is staged as zeros: every difference is 0 and adds nothing
      const bool have = a + h < a1;
      const float* __restrict__ row = traj + (size_t)(a + (have ? h : 0)) * T * 3;
      float* __restrict__ dxy = reinterpret_cast<float*>(h ? s_xyB : s_xyA);
      float* __restrict__ dz = reinterpret_cast<float*>(s_z2) + h;
      const float* __restrict__ src = row + (size_t)t_begin * 3;
      for (int e = tid; e < 3 * len; e += MD_NT) {
        const float v = have ? __ldg(src + e) : 0.f;
        const int t = e / 3, d = e - 3 * t;
        if (d < 2) dxy[2 * t + d] = v; else dz[2 * t] = v;
      }
      if (lag_blk0 > 0) {
        float* __restrict__ oxy = reinterpret_cast<float*>(h ? s_oB : s_oA);
        float* __restrict__ oz = reinterpret_cast<float*>(s_oz) + h;
        const float* __restrict__ so = row + (size_t)(t0 + w0) * 3;
        for (int e = tid; e < 3 * nw; e += MD_NT) {
          const float v = have ? __ldg(so + e) : 0.f;
          const int t = e / 3, d = e - 3 * t;
          if (d < 2) oxy[2 * t + d] = v; else oz[2 * t] = v;
        }
      }
    }
    __syncthreads();
    const float2* __restrict__ o_A = lag_blk0 > 0 ? s_oA : s_xyA;
    const float2* __restrict__ o_B = lag_blk0 > 0 ? s_oB : s_xyB;
    const float2* __restrict__ o_z = lag_blk0 > 0 ? s_oz : s_z2;

    float2 qA[R], qB[R], qz[R];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int p = min(ws_lo + kR + b, len_alloc - 1);
      qA[b] = s_xyA[p];
      qB[b] = s_xyB[p];
      qz[b] = s_z2[p];
    }
    float2 axy[R], az[R];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      axy[b] = make_float2(0.f, 0.f);
      az[b] = make_float2(0.f, 0.f);
    }
    int fold = 0;
    for (int w = ws_lo; w < ws_lo + nwg; w += R) {
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int ws = w + s;
        if (ws < ws_hi) {
          const float2 oA = o_A[ws], oB = o_B[ws], oz = o_z[ws];
          const float2 noA = make_float2(-oA.x, -oA.y);
          const float2 noB = make_float2(-oB.x, -oB.y);
          const float2 noz = make_float2(-oz.x, -oz.y);
#pragma unroll
          for (int b = 0; b < R; ++b) {
            const int slot = (s + b) % R;
            const float2 dA = __fadd2_rn(qA[slot], noA);
            const float2 dB = __fadd2_rn(qB[slot], noB);
            const float2 dz = __fadd2_rn(qz[slot], noz);
            axy[b] = __ffma2_rn(dA, dA, axy[b]);
            az[b] = __ffma2_rn(dz, dz, az[b]);
            axy[b] = __ffma2_rn(dB, dB, axy[b]);
          }
          // slot s held frame ws + lag: dead now; refill with frame ws + lag + R
          const int nx = min(ws + kR + R, len_alloc - 1);
          qA[s] = s_xyA[nx];
          qB[s] = s_xyB[nx];
          qz[s] = s_z2[nx];
        }
      }
      if (++fold == MD2_FOLD) {
        fold = 0;
#pragma unroll
        for (int b = 0; b < R; ++b) {
          acc64[b] += (double)((axy[b].x + axy[b].y) + (az[b].x + az[b].y));
          axy[b] = make_float2(0.f, 0.f);
          az[b] = make_float2(0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < R; ++b) acc64[b] += (double)((axy[b].x + axy[b].y) + (az[b].x + az[b].y));
  }
  if (active) {
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int lag = lag_blk0 + kR + b;
      if (kR + b < lags_here && lag < n_lags) atomicAdd(msd_sum + lag, acc64[b]);
    }
  }
}

// ---- Einstein MSD, short dense lag ranges (n_lags <= 16): HBM-streaming kernel ------------------
// With few lags per origin the work per byte is small and the trajectory read itself is the
// bound.  One warp streams one atom's row exactly once, in chunks of 128 frames (1536 contiguous
// bytes, three coalesced 512-byte LDG.128 rows), into a two-chunk ring in shared memory that
// keeps the global xyz interleaving (three STS.128 per lane); lane l owns the origins 32 s + l
// (s < 4) of the chunk and reads its lag partners x(origin + k) from the ring (lane stride 3
// words: conflict free).  The next chunk is prefetched in registers while the current one is
// processed.
constexpr int MS_WARPS = 4;
constexpr int MS_CH = 128;           // frames per chunk
constexpr int MS_RING = 2 * MS_CH;   // frames in the ring

template <int NL, bool VEC>
__global__ void __launch_bounds__(32 * MS_WARPS)
msd_stream_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                  long long t0, int W, int n_lags, double* __restrict__ msd_sum) {
  __shared__ __align__(16) float s_ring[MS_WARPS][3 * MS_RING];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* __restrict__ ring = s_ring[warp];
  const long long warps_total = (long long)gridDim.x * MS_WARPS;
  const long long n_fr = (long long)W + n_lags - 1;   // frames [t0, t0 + n_fr) are touched
  const long long n_el = n_fr * 3;
  const int n_chunks = (int)((W + MS_CH - 1) / MS_CH);
  double total = 0.0;  // lane k accumulates lag k

  auto load_chunk = [&](const float* __restrict__ src, int c, float4 (&r)[3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const long long e = (long long)c * (3 * MS_CH) + j * 128 + 4 * lane;
      if (VEC && e + 3 < n_el) {
        r[j] = __ldg(reinterpret_cast<const float4*>(src + e));
      } else {
        r[j].x = e + 0 < n_el ? __ldg(src + e + 0) : 0.f;
        r[j].y = e + 1 < n_el ? __ldg(src + e + 1) : 0.f;
        r[j].z = e + 2 < n_el ? __ldg(src + e + 2) : 0.f;
        r[j].w = e + 3 < n_el ? __ldg(src + e + 3) : 0.f;
      }
    }
  };
  auto store_chunk = [&](int c, const float4 (&r)[3]) {
    float* dst = ring + (c & 1) * (3 * MS_CH);
#pragma unroll
    for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(dst + j * 128 + 4 * lane) = r[j];
  };

  for (long long a = a_lo + (long long)blockIdx.x * MS_WARPS + warp; a < a_hi; a += warps_total) {
    const float* __restrict__ src = traj + ((size_t)a * T + t0) * 3;
    float acc[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) acc[k] = 0.f;
    float4 pre[3];
    load_chunk(src, 0, pre);
    __syncwarp();
    store_chunk(0, pre);
    load_chunk(src, 1, pre);
    for (int c = 0; c < n_chunks; ++c) {
      store_chunk(c + 1, pre);          // chunk c + 1 (lag partners of the last origins of c)
      load_chunk(src, c + 2, pre);      // in flight while chunk c is processed
      __syncwarp();
      const int base = (c & 1) * MS_CH;
#pragma unroll
      for (int sr = 0; sr < MS_CH / 32; ++sr) {
        const int w_loc = 32 * sr + lane;
        const bool valid = (long long)c * MS_CH + w_loc < W;
        const int o_idx = base + w_loc;
        const float ox = ring[3 * o_idx], oy = ring[3 * o_idx + 1], oz = ring[3 * o_idx + 2];
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          if (k < n_lags) {
            const int p = 3 * ((o_idx + k) & (MS_RING - 1));
            const float dx = ring[p] - ox, dy = ring[p + 1] - oy, dz = ring[p + 2] - oz;
            const float v = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            acc[k] += valid ? v : 0.f;
          }
        }
      }
      __syncwarp();
      if ((c & 63) == 63 || c + 1 == n_chunks) {
        // fold the lane-private fp32 sums: warp-reduce each lag, lane k keeps lag k in fp64
#pragma unroll
        for (int k = 0; k < NL; ++k) {
          float v = acc[k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == k) total += (double)v;
          acc[k] = 0.f;
        }
      }
    }
  }
  if (lane < n_lags) atomicAdd(msd_sum + lane, total);
}

// ---- Einstein MSD, medium dense lag ranges (8 .. 128 lags): register-window kernel --------------
// Between the HBM-bound streaming kernel (one shared-memory read per update) and the register
// ring (needs ~300 lags to fill a CTA) neither roofline was reached (round-1 review: 18 % of HBM
// and 20 % of FP32 at 16 lags).  Here one warp owns one atom and walks its row in chunks of
// 32 x RW_F origins: lane l owns the RW_F CONSECUTIVE origins 7 l .. 7 l + 6 and, per pass of NLP
// lags, loads the RW_F + NLP - 1 positions they pair with ONCE from shared memory into registers
// (lane stride 21 words: odd, conflict free) -- 0.6 shared-memory reads per update instead of 3.
// The RW_F x NLP updates of a pass are fully unrolled register arithmetic (3 FADD + 3 FFMA each).
// The per-lane partial sums of a pass are folded into a per-warp shared-memory table
// [lag][lane] (conflict free), which is reduced over the lanes into fp64 once per atom.
// Chunks (plus the lag halo) are staged by 16-byte cp.async, double buffered.
constexpr int RW_F = 7;
constexpr int RW_CH = 32 * RW_F;
constexpr int RW_WARPS = 4;
constexpr int RW_ST = 2;             // cp.async ring depth (chunks in flight per warp; 3 measured slower: fewer CTAs fit)

__device__ __forceinline__ void rw_cp16(float* dst, const float* src, bool valid) {
  const unsigned d = smem_u32(dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void rw_cp4(float* dst, const float* src, bool valid) {
  const unsigned d = smem_u32(dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int NLP, bool VEC>
__global__ void __launch_bounds__(32 * RW_WARPS)
msd_rw_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
              long long t0, int W, int n_lags, int n_pass, int slab_len,
              double* __restrict__ msd_sum) {
  extern __shared__ __align__(16) float rw_smem[];
  // layout per warp: slab[RW_ST][3 * slab_len (padded to 4)] | sacc[n_pass * NLP][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slab_fl = (3 * slab_len + 3) & ~3;
  const int nl_pad = n_pass * NLP;
  float* __restrict__ wbase = rw_smem + (size_t)warp * (RW_ST * slab_fl + nl_pad * 32);
  float* __restrict__ sacc = wbase + RW_ST * slab_fl;
  const long long warps_total = (long long)gridDim.x * RW_WARPS;
  const long long n_el = ((long long)W + n_lags - 1) * 3;   // valid floats of a row from t0 on
  const int n_chunks = (W + RW_CH - 1) / RW_CH;

  for (int k = lane; k < nl_pad * 32; k += 32) sacc[k] = 0.f;
  double total[4] = {0.0, 0.0, 0.0, 0.0};   // lane l keeps lags l + 32 r
  // up to 12 lags per pass the register budget allows a second set of per-lane sums
  constexpr bool RW_REG_ACC = NLP <= 12;
  float racc[RW_REG_ACC ? NLP : 1];
#pragma unroll
  for (int k = 0; k < (RW_REG_ACC ? NLP : 1); ++k) racc[k] = 0.f;

  // a request copies the slab of chunk c: 16-byte copies up to the end of the row; the rest of
  // the slab (last chunks only) is zero-filled by 4-byte copies with source size 0
  auto stage = [&](const float* __restrict__ src, int c) {
    float* dst = wbase + (c % RW_ST) * slab_fl;
    const int e0 = c * (3 * RW_CH);
    const int left = (int)min(n_el - e0, (long long)slab_fl);   // valid floats of this slab
    const int n16 = VEC ? (max(left, 0) & ~3) : 0;
    const float* __restrict__ s0 = src + e0;
#pragma unroll 1
    for (int q = 4 * lane; q < n16; q += 128) rw_cp16(dst + q, s0 + q, true);
#pragma unroll 1
    for (int q = n16 + lane; q < slab_fl; q += 32)
      rw_cp4(dst + q, q < left ? s0 + q : src, q < left);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (long long a = a_lo + (long long)blockIdx.x * RW_WARPS + warp; a < a_hi; a += warps_total) {
    const float* __restrict__ src = traj + ((size_t)a * T + t0) * 3;
#pragma unroll
    for (int st = 0; st < RW_ST - 1; ++st) {
      if (st < n_chunks) stage(src, st);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = 0; c < n_chunks; ++c) {
      asm volatile("cp.async.wait_group %0;" ::"n"(RW_ST - 2) : "memory");
      __syncwarp();   // chunk c is visible to all lanes, and all lanes have left chunk c - 1
      if (c + RW_ST - 1 < n_chunks) stage(src, c + RW_ST - 1);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      const float* __restrict__ sl = wbase + (c % RW_ST) * slab_fl + 3 * RW_F * lane;
      // origins, negated: {x, y} as a register pair (FADD2 / FFMA2); the z components of two
      // CONSECUTIVE origins share a register pair as well, so that a pair of origins costs six
      // packed instructions per lag (3 FADD2 + 3 FFMA2) instead of eight
      float2 nxy[RW_F];
      float nz[RW_F];
#pragma unroll
      for (int f = 0; f < RW_F; ++f) {
        nxy[f] = make_float2(-sl[3 * f], -sl[3 * f + 1]);
        nz[f] = -sl[3 * f + 2];
      }
      const int n_valid = min(RW_F, max(0, W - (c * RW_CH + RW_F * lane)));
#pragma unroll 1
      for (int g = 0; g < n_pass; ++g) {
        const float* __restrict__ pwp = sl + 3 * g * NLP;
        constexpr int WN = RW_F + NLP - 1;      // window positions of a pass
        float2 pxy[WN];
        // z of the window as register pairs in both alignments: ze[i] = {z(2i), z(2i+1)},
        // zo[i] = {z(2i+1), z(2i+2)} (origin pair (f, f+1), f even, at lag k needs the pair
        // starting at f + k, whose parity is that of k)
        float2 ze[(WN + 1) / 2], zo[WN / 2];
#pragma unroll
        for (int j = 0; j < WN; ++j) pxy[j] = make_float2(pwp[3 * j], pwp[3 * j + 1]);
#pragma unroll
        for (int i = 0; i < (WN + 1) / 2; ++i)
          ze[i] = make_float2(pwp[6 * i + 2], 2 * i + 1 < WN ? pwp[6 * i + 5] : 0.f);
#pragma unroll
        for (int i = 0; i < WN / 2; ++i)
          zo[i] = make_float2(ze[i].y, 2 * i + 2 < WN ? ze[i + 1 < (WN + 1) / 2 ? i + 1 : i].x : 0.f);
        float2 axy[NLP];
#pragma unroll
        for (int k = 0; k < NLP; ++k) axy[k] = make_float2(0.f, 0.f);
        if (n_valid == RW_F) {
#pragma unroll
          for (int k = 0; k < NLP; ++k) {
#pragma unroll
            for (int f = 0; f + 1 < RW_F; f += 2) {
              const int j = f + k;
              const float2 zp = (j & 1) ? zo[j >> 1] : ze[j >> 1];
              const float2 d0 = __fadd2_rn(pxy[j], nxy[f]);
              const float2 d1 = __fadd2_rn(pxy[j + 1], nxy[f + 1]);
              const float2 dz = __fadd2_rn(zp, make_float2(nz[f], nz[f + 1]));
              axy[k] = __ffma2_rn(d0, d0, axy[k]);
              axy[k] = __ffma2_rn(d1, d1, axy[k]);
              axy[k] = __ffma2_rn(dz, dz, axy[k]);
            }
            if (RW_F & 1) {   // the last origin of an odd count: z scalar
              const int f = RW_F - 1, j = f + k;
              const float2 d = __fadd2_rn(pxy[j], nxy[f]);
              const float dz = ((j & 1) ? ze[j >> 1].y : ze[j >> 1].x) + nz[f];
              axy[k] = __ffma2_rn(d, d, axy[k]);
              axy[k].x = fmaf(dz, dz, axy[k].x);
            }
          }
        } else {
#pragma unroll
          for (int f = 0; f < RW_F; ++f)
            if (f < n_valid) {
#pragma unroll
              for (int k = 0; k < NLP; ++k) {
                const int j = f + k;
                const float2 d = __fadd2_rn(pxy[j], nxy[f]);
                const float dz = ((j & 1) ? ze[j >> 1].y : ze[j >> 1].x) + nz[f];
                axy[k] = __ffma2_rn(d, d, axy[k]);
                axy[k].x = fmaf(dz, dz, axy[k].x);
              }
            }
        }
        float acc[NLP];
#pragma unroll
        for (int k = 0; k < NLP; ++k) acc[k] = axy[k].x + axy[k].y;
        if (RW_REG_ACC && n_pass == 1) {
          // a single lag pass: the per-lane sums stay in registers for the whole atom
#pragma unroll
          for (int k = 0; k < NLP; ++k) racc[k] += acc[k];
        } else {
          float* __restrict__ sa = sacc + (size_t)g * NLP * 32 + lane;
#pragma unroll
          for (int k = 0; k < NLP; ++k) sa[k * 32] += acc[k];
        }
      }
    }
    if (RW_REG_ACC && n_pass == 1) {
#pragma unroll
      for (int k = 0; k < NLP; ++k) {
        sacc[k * 32 + lane] = racc[k];
        racc[k] = 0.f;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();     // the ring is idle before the next atom's first chunks are requested
    // fold: lane l sums the 32 lane-partials of lags l, l + 32, ... (rotated columns: conflict free)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lag = lane + 32 * r;
      if (lag < nl_pad) {
        float v = 0.f;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int col = (j + lane) & 31;
          v += sacc[lag * 32 + col];
          sacc[lag * 32 + col] = 0.f;
        }
        total[r] += (double)v;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int lag = lane + 32 * r;
    if (lag < n_lags) atomicAdd(msd_sum + lag, total[r]);
  }
}

// ---- Green-Kubo lag products ------------------------------------------------------------
// P[t][m] += sum_a sum_d v[a,t,d] v[a,t+m,d].  grid.x: origin chunk of ACF_TC frames,
// grid.y: atom group.  Thread k owns lags k + r*NT and keeps ACF_TC x RL fp64 sums.
constexpr int ACF_TC = 8;
constexpr int ACF_G = 8;  // atoms summed in fp32 before folding into fp64

__global__ void __launch_bounds__(DYN_NT)
acf_lagprod_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                   int atoms_per_cta, long long t0, int B, int N, double* __restrict__ P) {
  extern __shared__ float s_tile[];  // 3 * (ACF_TC + N - 1)
  const int tid = threadIdx.x;
  const int tb = blockIdx.x * ACF_TC;  // first origin (relative to t0)
  if (tb >= B) return;
  const int len = min(ACF_TC + N - 1, B - tb);
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);

  for (int kb = 0; kb < N; kb += DYN_NT * DYN_RL) {
    double acc64[ACF_TC][DYN_RL];
    float acc32[ACF_TC][DYN_RL];
#pragma unroll
    for (int t = 0; t < ACF_TC; ++t)
#pragma unroll
      for (int r = 0; r < DYN_RL; ++r) {
        acc64[t][r] = 0.0;
        acc32[t][r] = 0.f;
      }
    int in_group = 0;
    for (long long a = a0; a < a1; ++a) {
      const float* __restrict__ src = traj + ((size_t)a * T + t0 + tb) * 3;
      __syncthreads();
      for (int e = tid; e < 3 * len; e += DYN_NT) s_tile[e] = __ldg(src + e);
      __syncthreads();
#pragma unroll
      for (int t = 0; t < ACF_TC; ++t) {
        if (t < len) {
          const float x0 = s_tile[3 * t], y0 = s_tile[3 * t + 1], z0 = s_tile[3 * t + 2];
#pragma unroll
          for (int r = 0; r < DYN_RL; ++r) {
            const int m = kb + r * DYN_NT + tid;
            if (m < N && t + m < len) {
              const float* __restrict__ q = s_tile + 3 * (t + m);
              acc32[t][r] = fmaf(x0, q[0], acc32[t][r]);
              acc32[t][r] = fmaf(y0, q[1], acc32[t][r]);
              acc32[t][r] = fmaf(z0, q[2], acc32[t][r]);
            }
          }
        }
      }
      if (++in_group == ACF_G) {
        in_group = 0;
#pragma unroll
        for (int t = 0; t < ACF_TC; ++t)
#pragma unroll
          for (int r = 0; r < DYN_RL; ++r) {
            acc64[t][r] += (double)acc32[t][r];
            acc32[t][r] = 0.f;
          }
      }
    }
#pragma unroll
    for (int t = 0; t < ACF_TC; ++t)
#pragma unroll
      for (int r = 0; r < DYN_RL; ++r) {
        const int m = kb + r * DYN_NT + tid;
        if (t < len && m < N && t + m < len) {
          const double v = acc64[t][r] + (double)acc32[t][r];
          atomicAdd(P + (size_t)(tb + t) * N + m, v);
        }
      }
  }
}

// ---- Green-Kubo lag products, register-tiled band Gram kernel -------------------------------
// P[t][t'-t] += sum_a sum_d v[a][t][d] * v[a][t'][d] for t <= t' < t + N: the band of the Gram
// matrix of the (3A x B) velocity matrix.  A CTA owns a 64 x 64 tile of (t, t') and a slice of
// the atoms; a thread owns 8 x 8 outputs (64 fp32 accumulators as 32 f32x2 pairs) so that one
// atom costs 96 FFMA2 against 12 LDS.128.  Operands are staged per atom as SoA rows
// [side][dim][64] by 4-byte cp.async (global layout is [A][T][3], so a (atom, 64 frames) slab is
// 768 contiguous bytes) in a 3-stage ring.  fp32 accumulation runs over at most GB_FOLD atoms
// (3 * GB_FOLD products), then folds into the global fp64 P with atomicAdd(double).
// A 12 x 8 tile per thread (96 x 64 per CTA; 4.8 instead of 4 FMAs per operand float, which
// would leave the FMA pipe as the only limiter) was built and measured in round 2: 4.58e12
// against 5.46e12 updates/s -- it needs 254 registers (8 warps per SM), capped at 168 it keeps
// accumulators in local memory, and the wider tiles waste more of the band's edges.  Removed.
constexpr int GB_T = 64;      // tile edge
constexpr int GB_NT = 64;     // threads per CTA (2 warps)
constexpr int GB_KA = 4;      // atoms per pipeline stage (22 KB per CTA with 3 stages)
constexpr int GB_ST = 3;      // pipeline stages
constexpr int GB_FOLD = 256;  // atoms per fp32 accumulation run
constexpr int GB_ROW = GB_T + 12;        // padded SoA row: dims land 12 banks apart, so the
                                        // 4-byte cp.async scatter (lane -> frame q/3, dim q%3) is
                                        // (almost) conflict free; multiple of 4 for LDS.128
constexpr int GB_SIDE = 3 * GB_ROW;      // floats per (atom, side)
constexpr int GB_SLAB = 2 * GB_SIDE;     // floats per atom per stage (a side + b side)

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
  const unsigned d = smem_u32(dst_smem);
  const int sz = valid ? 4 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(GB_NT, 8)
acf_band_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                int atoms_per_cta, long long t0, int B, int N, int nbj, double* __restrict__ P) {
  extern __shared__ __align__(16) float gb_smem[];  // GB_ST * GB_KA * GB_SLAB floats
  const int tid = threadIdx.x;
  const int bi = blockIdx.x / nbj;
  const int bj = bi + blockIdx.x % nbj;
  const int ta0 = bi * GB_T, tb0 = bj * GB_T;  // tile origins (relative to t0)
  if (tb0 >= B) return;
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);
  if (a0 >= a1) return;

  const int warp = tid >> 5, lane = tid & 31;
  const int li = lane >> 3, lj = lane & 7;
  const int ra = 32 * warp + 4 * li;  // local a rows: ra + (i&3) + 16*(i>>2)
  const int cb = 4 * lj;              // local b cols: cb + (j&3) + 32*(j>>2)

  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

  // stage loader.  A stage holds GB_KA atoms x 2 sides x 192 floats; for one (atom, side) the
  // 192 source floats are contiguous in global memory.  Thread tid copies the three elements
  // q = tid, tid + 64, tid + 128 of every (atom, side) slab: their dim/frame split and smem
  // offsets are per-thread constants.
  int q_src[3], q_dst[3], q_t[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int q = c * GB_NT + tid;
    q_t[c] = q / 3;
    q_src[c] = q;
    q_dst[c] = (q - 3 * q_t[c]) * GB_ROW + q_t[c];
  }
  bool ok_a[3], ok_b[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ok_a[c] = ta0 + q_t[c] < B;
    ok_b[c] = tb0 + q_t[c] < B;
  }
  const float* const src_a0 = traj + ((size_t)a0 * T + t0 + ta0) * 3;
  const float* const src_b0 = traj + ((size_t)a0 * T + t0 + tb0) * 3;
  const size_t row_stride = (size_t)T * 3;
  auto load_stage = [&](long long a_first, int stage) {
    float* dst = gb_smem + (size_t)stage * GB_KA * GB_SLAB;
    const float* sa = src_a0 + (size_t)(a_first - a0) * row_stride;
    const float* sb = src_b0 + (size_t)(a_first - a0) * row_stride;
#pragma unroll 1
    for (int ka = 0; ka < GB_KA; ++ka) {
      const bool live = a_first + ka < a1;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const bool va = live && ok_a[c], vb = live && ok_b[c];
        cp_async4(dst + q_dst[c], va ? sa + q_src[c] : src_a0, va);
        cp_async4(dst + GB_SIDE + q_dst[c], vb ? sb + q_src[c] : src_a0, vb);
      }
      dst += GB_SLAB;
      sa += row_stride;
      sb += row_stride;
    }
  };

  auto flush = [&]() {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ta = ta0 + ra + (i & 3) + 16 * (i >> 2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int tb = tb0 + cb + (j & 3) + 32 * (j >> 2);
        const int m = tb - ta;
        const float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
        if (ta < B && tb < B && m >= 0 && m < N) atomicAdd(P + (size_t)ta * N + m, (double)v);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  };

  const int n_stage_total = (int)((a1 - a0 + GB_KA - 1) / GB_KA);
  // prologue
#pragma unroll
  for (int s = 0; s < GB_ST - 1; ++s) {
    if (s < n_stage_total) load_stage(a0 + (long long)s * GB_KA, s);
    cp_async_commit();
  }
  int since_fold = 0;
  for (int st = 0; st < n_stage_total; ++st) {
    cp_async_wait<GB_ST - 2>();
    __syncthreads();
    {
      const int nxt = st + GB_ST - 1;
      if (nxt < n_stage_total) load_stage(a0 + (long long)nxt * GB_KA, nxt % GB_ST);
      cp_async_commit();
    }
    const float* sbase = gb_smem + (size_t)(st % GB_ST) * GB_KA * GB_SLAB;
#pragma unroll 2
    for (int ka = 0; ka < GB_KA; ++ka) {
      const float* sa = sbase + ka * GB_SLAB;
      const float* sb = sa + GB_SIDE;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float4 a_lo4 = *reinterpret_cast<const float4*>(sa + d * GB_ROW + ra);
        const float4 a_hi4 = *reinterpret_cast<const float4*>(sa + d * GB_ROW + ra + 16);
        const float4 b_lo4 = *reinterpret_cast<const float4*>(sb + d * GB_ROW + cb);
        const float4 b_hi4 = *reinterpret_cast<const float4*>(sb + d * GB_ROW + cb + 32);
        const float av[8] = {a_lo4.x, a_lo4.y, a_lo4.z, a_lo4.w, a_hi4.x, a_hi4.y, a_hi4.z, a_hi4.w};
        const float2 bv[4] = {make_float2(b_lo4.x, b_lo4.y), make_float2(b_lo4.z, b_lo4.w),
                              make_float2(b_hi4.x, b_hi4.y), make_float2(b_hi4.z, b_hi4.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a2 = make_float2(av[i], av[i]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(a2, bv[j], acc[i][j]);
        }
      }
    }
    since_fold += GB_KA;
    if (since_fold >= GB_FOLD) {
      flush();
      since_fold = 0;
    }
  }
  cp_async_wait<0>();
  flush();
}

// ---- Green-Kubo lag products, short lag ranges (N <= 16): HBM-streaming kernel ----------------
// P[t][m] += sum_a sum_d v[a][t][d] v[a][t+m][d] needs one accumulator per (t, m), summed over the
// atoms: a warp owns a chunk of 128 origins and loops over the atoms of its group, keeping
// 4 x NL fp32 sums per lane (lane l <-> origins 32 s + l) that are folded into the global fp64 P
// every ACS_FOLD atoms.  Per atom it streams 128 + 32 frames (the chunk and the lag partners of
// its last origins) through a per-warp shared-memory slab; the next atom is prefetched in
// registers while the current one is processed.
constexpr int ACS_WARPS = 4;
constexpr int ACS_CH = 128;
constexpr int ACS_HALO = 32;
constexpr int ACS_FOLD = 256;

template <int NL, bool VEC>
__global__ void __launch_bounds__(32 * ACS_WARPS)
acf_stream_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
                  int atoms_per_cta, long long t0, int B, int N, double* __restrict__ P) {
  __shared__ __align__(16) float s_slab[ACS_WARPS][3 * (ACS_CH + ACS_HALO)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* __restrict__ slab = s_slab[warp];
  const int tc = blockIdx.x * ACS_CH;  // first origin of the chunk (relative to t0)
  if (tc >= B) return;
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);
  const long long n_el = (long long)(B - tc) * 3;  // elements of this atom row from the chunk on

  float acc[ACS_CH / 32][NL];
#pragma unroll
  for (int sr = 0; sr < ACS_CH / 32; ++sr)
#pragma unroll
    for (int k = 0; k < NL; ++k) acc[sr][k] = 0.f;

  auto load_atom = [&](long long a, float4 (&r)[3], float (&h)[3]) {
    const float* __restrict__ src = traj + ((size_t)a * T + t0 + tc) * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const long long e = j * 128 + 4 * lane;
      if (VEC && e + 3 < n_el) {
        r[j] = __ldg(reinterpret_cast<const float4*>(src + e));
      } else {
        r[j].x = e + 0 < n_el ? __ldg(src + e + 0) : 0.f;
        r[j].y = e + 1 < n_el ? __ldg(src + e + 1) : 0.f;
        r[j].z = e + 2 < n_el ? __ldg(src + e + 2) : 0.f;
        r[j].w = e + 3 < n_el ? __ldg(src + e + 3) : 0.f;
      }
      const long long eh = 3 * ACS_CH + j * 32 + lane;
      h[j] = eh < n_el ? __ldg(src + eh) : 0.f;
    }
  };
  auto flush = [&]() {
#pragma unroll
    for (int sr = 0; sr < ACS_CH / 32; ++sr) {
      const int t = tc + 32 * sr + lane;
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        if (k < N && t + k < B) atomicAdd(P + (size_t)t * N + k, (double)acc[sr][k]);
        acc[sr][k] = 0.f;
      }
    }
  };

  float4 pre[3];
  float preh[3];
  long long a = a0 + warp;
  if (a < a1) load_atom(a, pre, preh);
  int since_fold = 0;
  for (; a < a1; a += ACS_WARPS) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      *reinterpret_cast<float4*>(slab + j * 128 + 4 * lane) = pre[j];
      slab[3 * ACS_CH + j * 32 + lane] = preh[j];
    }
    if (a + ACS_WARPS < a1) load_atom(a + ACS_WARPS, pre, preh);  // in flight during the math
    __syncwarp();
#pragma unroll
    for (int sr = 0; sr < ACS_CH / 32; ++sr) {
      const int o = 3 * (32 * sr + lane);
      const float ox = slab[o], oy = slab[o + 1], oz = slab[o + 2];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        if (k < N) {
          const int p = o + 3 * k;
          acc[sr][k] = fmaf(oz, slab[p + 2], fmaf(oy, slab[p + 1], fmaf(ox, slab[p], acc[sr][k])));
        }
      }
    }
    if (++since_fold == ACS_FOLD) {
      flush();
      since_fold = 0;
    }
  }
  flush();
}

// ---- Green-Kubo lag products, medium lag ranges (5 .. 128 lags): register-window kernel --------
// The band-Gram tiles are 64 x 64 in (t, t'): for N < ~200 most of a tile lies outside the band
// 0 <= t' - t < N (19 % useful at N = 24, 52 % at N = 100).  In (t, m) coordinates the band is a
// rectangle, so a CTA here owns ARW_CH = 32 x ARW_F consecutive origins and an atom slice; warp g
// owns the lag block [16 g, 16 g + 16) and lane l the ARW_F consecutive origins 5 l .. 5 l + 4:
// per atom the lane loads its 5 origin and 5 + 15 window velocities from the slab the CTA staged
// (16-byte cp.async, double buffered over atoms; lane stride 15 words: conflict free) and
// updates its 5 x 16 fp32 sums in registers (3 FFMA each) -- one shared-memory read per update,
// no work outside the band.  The sums are folded into the global fp64 P every ARW_FOLD atoms.
constexpr int ARW_F = 5;
constexpr int ARW_CH = 32 * ARW_F;   // 160 origins per CTA
constexpr int ARW_NLP = 16;
constexpr int ARW_FOLD = 256;
constexpr int ARW_ST = 4;            // cp.async ring depth (atoms in flight per CTA)

template <bool VEC>
__global__ void __launch_bounds__(256)
acf_rw_kernel(const float* __restrict__ traj, long long T, long long a_lo, long long a_hi,
              int atoms_per_cta, long long t0, int B, int N, int slab_len,
              double* __restrict__ P) {
  extern __shared__ __align__(16) float arw_smem[];   // ARW_ST slabs of 3 * slab_len floats (padded)
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int nthreads = blockDim.x;
  const int slab_fl = (3 * slab_len + 3) & ~3;
  const int tc = blockIdx.x * ARW_CH;   // first origin of the chunk (relative to t0)
  if (tc >= B) return;
  const long long a0 = a_lo + (long long)blockIdx.y * atoms_per_cta;
  const long long a1 = min(a_hi, a0 + atoms_per_cta);
  if (a0 >= a1) return;
  const long long n_el = (long long)(B - tc) * 3;   // floats of an atom row from the chunk on
  const int m0 = g * ARW_NLP;

  // 16-byte copies cover the slab up to the end of the row; what lies beyond (last chunk only)
  // is zero-filled by 4-byte copies with source size 0
  const int n16 = VEC ? (int)(min((long long)slab_fl, n_el) & ~3ll) : 0;
  const size_t row_stride = (size_t)T * 3;
  // this thread's slice of the copy: source pointer of the next atom to request (advanced by one
  // row per request) and fixed offsets, so that a request is a handful of instructions
  const float* __restrict__ next_src = traj + ((size_t)a0 * T + t0 + tc) * 3;
  auto stage = [&](int buf) {
    float* dst = arw_smem + buf * slab_fl;
#pragma unroll 1
    for (int q = 4 * tid; q < n16; q += 4 * nthreads) rw_cp16(dst + q, next_src + q, true);
#pragma unroll 1
    for (int q = n16 + tid; q < slab_fl; q += nthreads)
      rw_cp4(dst + q, next_src + (q < n_el ? q : 0), q < n_el);
    asm volatile("cp.async.commit_group;" ::: "memory");
    next_src += row_stride;
  };

  float acc[ARW_F][ARW_NLP];
#pragma unroll
  for (int f = 0; f < ARW_F; ++f)
#pragma unroll
    for (int k = 0; k < ARW_NLP; ++k) acc[f][k] = 0.f;

  auto flush = [&]() {
#pragma unroll
    for (int f = 0; f < ARW_F; ++f) {
      const int t = tc + ARW_F * lane + f;
#pragma unroll
      for (int k = 0; k < ARW_NLP; ++k) {
        const int m = m0 + k;
        if (t < B && m < N && t + m < B) atomicAdd(P + (size_t)t * N + m, (double)acc[f][k]);
        acc[f][k] = 0.f;
      }
    }
  };

  // ring of ARW_ST slabs: atom a + ARW_ST - 1 is requested while atom a is processed.  One block
  // barrier per atom does double duty: the slab of atom a is visible to every warp, and every
  // warp has left atom a - 1, whose slab the new request overwrites.
#pragma unroll
  for (int st = 0; st < ARW_ST - 1; ++st) {
    if (a0 + st < a1) stage(st);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int since_fold = 0, buf = 0;
  for (long long a = a0; a < a1; ++a) {
    asm volatile("cp.async.wait_group %0;" ::"n"(ARW_ST - 2) : "memory");
    __syncthreads();
    if (a + ARW_ST - 1 < a1) stage(buf == 0 ? ARW_ST - 1 : buf - 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* __restrict__ sl = arw_smem + buf * slab_fl + 3 * ARW_F * lane;
    float po[ARW_F][3];
#pragma unroll
    for (int f = 0; f < ARW_F; ++f)
#pragma unroll
      for (int d = 0; d < 3; ++d) po[f][d] = sl[3 * f + d];
    const float* __restrict__ pwp = sl + 3 * m0;
    float pw[ARW_F + ARW_NLP - 1][3];
#pragma unroll
    for (int j = 0; j < ARW_F + ARW_NLP - 1; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) pw[j][d] = pwp[3 * j + d];
#pragma unroll
    for (int f = 0; f < ARW_F; ++f)
#pragma unroll
      for (int k = 0; k < ARW_NLP; ++k)
        acc[f][k] = fmaf(po[f][2], pw[f + k][2],
                         fmaf(po[f][1], pw[f + k][1], fmaf(po[f][0], pw[f + k][0], acc[f][k])));
    buf = buf + 1 == ARW_ST ? 0 : buf + 1;
    if (++since_fold == ARW_FOLD) {
      flush();
      since_fold = 0;
    }
  }
  flush();
}

// ---- prefix sum of P along t (in place, inclusive), 8 lags x 128 time chunks per CTA -----
constexpr int SCAN_M = 8;
constexpr int SCAN_C = 128;

__global__ void __launch_bounds__(SCAN_M* SCAN_C)
acf_prefix_kernel(double* __restrict__ P, int B, int N) {
  __shared__ double s_tot[SCAN_C][SCAN_M];
  const int lm = threadIdx.x % SCAN_M;
  const int c = threadIdx.x / SCAN_M;
  const int m = blockIdx.x * SCAN_M + lm;
  const int per = (B + SCAN_C - 1) / SCAN_C;
  const int t_lo = c * per, t_hi = min(B, t_lo + per);
  double run = 0.0;
  if (m < N)
    for (int t = t_lo; t < t_hi; ++t) run += P[(size_t)t * N + m];
  s_tot[c][lm] = run;
  __syncthreads();
  double off = 0.0;
  for (int cc = 0; cc < c; ++cc) off += s_tot[cc][lm];
  if (m < N) {
    run = off;
    for (int t = t_lo; t < t_hi; ++t) {
      run += P[(size_t)t * N + m];
      P[(size_t)t * N + m] = run;
    }
  }
}

// ---- window sums from the prefix array --------------------------------------------------
// grid.x: lag block, grid.y: window chunk of ACFW_WC windows
constexpr int ACFW_WC = 32;

__global__ void __launch_bounds__(128)
acf_windows_kernel(const double* __restrict__ C, int B, int N, int W, int ct,
                   double* __restrict__ acf_sum, double* __restrict__ acf_win) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= N) return;
  const int w0 = blockIdx.y * ACFW_WC, w1 = min(W, w0 + ACFW_WC);
  const double inv = 1.0 / (double)(N - m);
  double tot = 0.0;
  for (int w = w0; w < w1; ++w) {
    const long long s = (long long)w * ct;
    const long long e = s + N - 1 - m;
    double v = C[(size_t)e * N + m];
    if (s > 0) v -= C[(size_t)(s - 1) * N + m];
    v *= inv;
    if (acf_win) acf_win[(size_t)w * N + m] = v;
    tot += v;
  }
  atomicAdd(acf_sum + m, tot);
}

static int pick_atoms_per_cta(long long n_atoms, long long chunks, int ctas_per_sm = 8) {
  // aim for ~ctas_per_sm CTAs per SM overall while keeping >= 1 atom per CTA
  const long long target = (long long)sm_count() * ctas_per_sm;
  long long groups = (target + chunks - 1) / chunks;
  if (groups < 1) groups = 1;
  if (groups > n_atoms) groups = n_atoms;
  long long apc = (n_atoms + groups - 1) / groups;
  if (apc < 1) apc = 1;
  return (int)apc;
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_msd_windowed(const float* traj, long long A, long long T, long long a_lo,
                                long long a_hi, long long t0, int W, int ct, const int* tau,
                                int n_tau, int span, double* msd_sum, mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && tau && msd_sum, "msd_windowed: null pointer");
  MDK_CHECK_ARG(0 <= a_lo && a_lo <= a_hi && a_hi <= A, "msd_windowed: bad atom range");
  MDK_CHECK_ARG(W >= 0 && ct >= 1 && n_tau >= 1 && span >= 1, "msd_windowed: bad window spec");
  if (W == 0 || a_lo == a_hi) return MDK_OK;
  MDK_CHECK_ARG(t0 >= 0 && t0 + (long long)(W - 1) * ct + span <= T,
                "msd_windowed: windows [t0=%lld, W=%d, ct=%d, span=%d] exceed T=%lld", t0, W, ct,
                span, T);
  // window chunk: bounded by shared memory (<= ~96 KB tile)
  int Wc = 512;
  const long long max_len = (96 * 1024) / 12;
  while (Wc > 1 && (long long)(Wc - 1) * ct + span > max_len) Wc /= 2;
  const long long len = (long long)(Wc - 1) * ct + span;
  if (len * 12 > 200 * 1024) {
    set_error("msd_windowed: data_range span %d does not fit in shared memory", span);
    return MDK_EUNSUPPORTED;
  }
  const int chunks = (W + Wc - 1) / Wc;
  const int apc = pick_atoms_per_cta(a_hi - a_lo, chunks);
  const long long groups = (a_hi - a_lo + apc - 1) / apc;
  MDK_CHECK_ARG(groups <= 65535, "msd_windowed: too many atom groups");
  const size_t smem = (size_t)len * 12;
  MDK_CUDA(cudaFuncSetAttribute(msd_windowed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid(chunks, (unsigned)groups);
  msd_windowed_kernel<<<grid, DYN_NT, smem, as_stream(stream)>>>(
      traj, T, a_lo, a_hi, apc, t0, W, ct, tau, n_tau, span, Wc, msd_sum);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_msd_dense(const float* traj, long long A, long long T, long long a_lo,
                             long long a_hi, long long t0, int W, int n_lags, double* msd_sum,
                             mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && msd_sum, "msd_dense: null pointer");
  MDK_CHECK_ARG(0 <= a_lo && a_lo <= a_hi && a_hi <= A, "msd_dense: bad atom range");
  MDK_CHECK_ARG(W >= 0 && n_lags >= 1, "msd_dense: bad window spec");
  if (W == 0 || a_lo == a_hi) return MDK_OK;
  MDK_CHECK_ARG(t0 >= 0 && t0 + (long long)(W - 1) + n_lags <= T,
                "msd_dense: windows [t0=%lld, W=%d, n_lags=%d] exceed T=%lld", t0, W, n_lags, T);
  // medium lag ranges: register-window kernel (MDK_MSD_RW_MIN / _MAX move the hand-over points)
  int rw_min = 5, rw_max = 128;
  if (const char* e = getenv("MDK_MSD_RW_MIN")) rw_min = atoi(e);
  if (const char* e = getenv("MDK_MSD_RW_MAX")) rw_max = atoi(e);
  if (rw_max > 128) rw_max = 128;
  if (n_lags >= rw_min && n_lags <= rw_max) {
    const bool vec = (T % 4 == 0) && (t0 % 4 == 0) && (reinterpret_cast<uintptr_t>(traj) % 16 == 0);
    // lags per pass: the block size that pads the lag range least; on ties in the measured order
    // of preference 12, 16, 8, 20 (B200, 125k atoms x 2000 frames: 0.41-0.44 of the FP32 peak
    // at 24..64 lags with a fitting block, 0.26-0.33 with a padded one)
    int NLP = 16;
    if (const char* e = getenv("MDK_MSD_RW_NLP")) {
      NLP = atoi(e);
    } else {
      int best_pad = 1 << 30;
      for (int cand : {12, 16, 8, 20}) {
        const int pad = (n_lags + cand - 1) / cand * cand;
        if (pad < best_pad) {
          best_pad = pad;
          NLP = cand;
        }
      }
    }
    MDK_CHECK_ARG(NLP == 8 || NLP == 12 || NLP == 16 || NLP == 20, "msd_dense: bad MDK_MSD_RW_NLP");
    const int n_pass = (n_lags + NLP - 1) / NLP;
    const int slab_len = RW_CH + n_pass * NLP + RW_F;   // chunk + lag halo (+ slack of the last lane)
    const size_t per_warp = (size_t)(RW_ST * ((3 * slab_len + 3) & ~3) + n_pass * NLP * 32) * sizeof(float);
    const size_t smem = per_warp * RW_WARPS;
    const long long warps = a_hi - a_lo;
    long long blocks = (warps + RW_WARPS - 1) / RW_WARPS;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 12) per_sm = 12;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = as_stream(stream);
#define MDK_RW_LAUNCH(NLPV, V)                                                                  \
  do {                                                                                          \
    MDK_CUDA(cudaFuncSetAttribute(msd_rw_kernel<NLPV, V>,                                       \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    msd_rw_kernel<NLPV, V><<<(unsigned)blocks, 32 * RW_WARPS, smem, st>>>(                      \
        traj, T, a_lo, a_hi, t0, W, n_lags, n_pass, slab_len, msd_sum);                         \
  } while (0)
    if (NLP == 8) { if (vec) MDK_RW_LAUNCH(8, true); else MDK_RW_LAUNCH(8, false); }
    else if (NLP == 12) { if (vec) MDK_RW_LAUNCH(12, true); else MDK_RW_LAUNCH(12, false); }
    else if (NLP == 16) { if (vec) MDK_RW_LAUNCH(16, true); else MDK_RW_LAUNCH(16, false); }
    else { if (vec) MDK_RW_LAUNCH(20, true); else MDK_RW_LAUNCH(20, false); }
#undef MDK_RW_LAUNCH
    MDK_LAUNCH_CHECK();
    return MDK_OK;
  }
  if (n_lags <= 16 && !getenv("MDK_MSD_NO_STREAM")) {
    // short lag ranges: HBM-streaming kernel, one warp per atom
    const bool vec = (T % 4 == 0) && (t0 % 4 == 0) && (reinterpret_cast<uintptr_t>(traj) % 16 == 0);
    const long long warps = a_hi - a_lo;
    long long blocks = (warps + MS_WARPS - 1) / MS_WARPS;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const unsigned nb = (unsigned)blocks, nt = 32 * MS_WARPS;
    cudaStream_t st = as_stream(stream);
#define MDK_MS_LAUNCH(NL)                                                                      \
  do {                                                                                         \
    if (vec)                                                                                   \
      msd_stream_kernel<NL, true><<<nb, nt, 0, st>>>(traj, T, a_lo, a_hi, t0, W, n_lags, msd_sum); \
    else                                                                                       \
      msd_stream_kernel<NL, false><<<nb, nt, 0, st>>>(traj, T, a_lo, a_hi, t0, W, n_lags, msd_sum); \
  } while (0)
    if (n_lags <= 4) MDK_MS_LAUNCH(4);
    else if (n_lags <= 8) MDK_MS_LAUNCH(8);
    else MDK_MS_LAUNCH(16);
#undef MDK_MS_LAUNCH
    MDK_LAUNCH_CHECK();
    return MDK_OK;
  }
  // window groups (one-atom kernel, R = 9) only pay off when the lag range leaves at least half
  // of the CTA idle; otherwise the two-atom kernel with the smallest odd R that covers the range
  const bool grouped = (n_lags + MD_R - 1) / MD_R <= MD_NT / 2;
  // grouped: the odd R in {5, 7, 9} that keeps most threads busy (G lag-threads x NG groups;
  // ties go to the larger R, which reuses a loaded position more often).  The two-atom kernel
  // was measured slower there (register pressure of the group bookkeeping).
  int R = n_lags <= MD_NT * 5 ? 5 : (n_lags <= MD_NT * 7 ? 7 : 9);
  if (grouped) {
    int best = -1;
    for (int r : {9, 7, 5}) {
      const int G = (n_lags + r - 1) / r;
      if (G > MD_NT) continue;
      const int busy = G * (MD_NT / G);
      if (busy > best) {
        best = busy;
        R = r;
      }
    }
  }
  const int lag_span = MD_NT * R;
  const int lag_blocks = (n_lags + lag_span - 1) / lag_span;
  // every thread may read up to one ring refill past its last lag: the slab covers the window
  // chunk plus the lag span of the block (plus R), so those (discarded) reads stay inside it.
  // Short lag ranges take long window chunks (up to ~4096 frames, 48 KB): with few lags per
  // atom the staging, not the arithmetic, is the cost, and it amortises over more origins.
  const int lag_alloc = grouped ? ((n_lags + R - 1) / R) * R + R : lag_span + R;
  const int wc_max = grouped ? 4096 - lag_alloc : 512;
  const int Wc = W < wc_max ? W : wc_max;
  const int len_alloc = (Wc + lag_alloc + 3) & ~3;  // multiple of 4: float2 views stay aligned
  // the two-atom kernel stages two atoms per sweep
  const size_t smem = ((size_t)(grouped ? 3 : 6) * len_alloc +
                       (lag_blocks > 1 ? (size_t)6 * Wc : 0)) * sizeof(float);
  const int chunks = (W + Wc - 1) / Wc;
  // 64-thread CTAs: aim for several waves of ~11 resident CTAs per SM
  const int apc = pick_atoms_per_cta(a_hi - a_lo, (long long)chunks * lag_blocks, 48);
  const long long groups = (a_hi - a_lo + apc - 1) / apc;
  MDK_CHECK_ARG(groups <= 65535 && lag_blocks <= 65535, "msd_dense: grid too large");
  dim3 grid(chunks, (unsigned)groups, lag_blocks);
  if (grouped) {
#define MDK_MD1_LAUNCH(RR)                                                                    \
  do {                                                                                        \
    MDK_CUDA(cudaFuncSetAttribute(msd_dense_kernel<true, RR>,                                 \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    msd_dense_kernel<true, RR><<<grid, MD_NT, smem, as_stream(stream)>>>(                     \
        traj, T, a_lo, a_hi, apc, t0, W, n_lags, Wc, len_alloc, msd_sum);                     \
  } while (0)
    if (R == 5) MDK_MD1_LAUNCH(5);
    else if (R == 7) MDK_MD1_LAUNCH(7);
    else MDK_MD1_LAUNCH(9);
#undef MDK_MD1_LAUNCH
  } else {
#define MDK_MD2_LAUNCH(RR)                                                                    \
  do {                                                                                        \
    MDK_CUDA(cudaFuncSetAttribute(msd_dense2_kernel<false, RR>,                               \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    msd_dense2_kernel<false, RR><<<grid, MD_NT, smem, as_stream(stream)>>>(                   \
        traj, T, a_lo, a_hi, apc, t0, W, n_lags, Wc, len_alloc, msd_sum);                     \
  } while (0)
    if (R == 5) MDK_MD2_LAUNCH(5);
    else if (R == 7) MDK_MD2_LAUNCH(7);
    else MDK_MD2_LAUNCH(9);
#undef MDK_MD2_LAUNCH
  }
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_acf_lagprod(const float* traj, long long A, long long T, long long a_lo,
                               long long a_hi, long long t0, int B, int N, double* P,
                               mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && P, "acf_lagprod: null pointer");
  MDK_CHECK_ARG(0 <= a_lo && a_lo <= a_hi && a_hi <= A, "acf_lagprod: bad atom range");
  MDK_CHECK_ARG(B >= 1 && N >= 1 && t0 >= 0 && t0 + B <= T, "acf_lagprod: bad frame range");
  if (a_lo == a_hi) return MDK_OK;
  const long long n_atoms = a_hi - a_lo;
  const char* legacy = getenv("MDK_ACF_LEGACY");
  if (legacy && legacy[0] == '1') {
    const size_t smem = (size_t)(ACF_TC + N - 1) * 12;
    if (smem > 200 * 1024) {
      set_error("acf_lagprod: data_range %d does not fit in shared memory", N);
      return MDK_EUNSUPPORTED;
    }
    const int chunks = (B + ACF_TC - 1) / ACF_TC;
    const int apc = pick_atoms_per_cta(n_atoms, chunks);
    const long long groups = (n_atoms + apc - 1) / apc;
    MDK_CHECK_ARG(groups <= 65535, "acf_lagprod: too many atom groups");
    MDK_CUDA(cudaFuncSetAttribute(acf_lagprod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    dim3 grid(chunks, (unsigned)groups);
    acf_lagprod_kernel<<<grid, DYN_NT, smem, as_stream(stream)>>>(traj, T, a_lo, a_hi, apc, t0, B,
                                                                  N, P);
    MDK_LAUNCH_CHECK();
    return MDK_OK;
  }
  // medium lag ranges: register-window kernel (MDK_ACF_RW_MIN / _MAX move the hand-over points)
  int arw_min = 9, arw_max = 72;   // measured: streaming kernel up to 8 lags, band Gram from ~80
  if (const char* e = getenv("MDK_ACF_RW_MIN")) arw_min = atoi(e);
  if (const char* e = getenv("MDK_ACF_RW_MAX")) arw_max = atoi(e);
  if (arw_max > 128) arw_max = 128;
  if (N >= arw_min && N <= arw_max) {
    const bool vec = (T % 4 == 0) && (t0 % 4 == 0) && (reinterpret_cast<uintptr_t>(traj) % 16 == 0);
    const int n_blk = (N + ARW_NLP - 1) / ARW_NLP;           // warps per CTA (lag blocks)
    const int slab_len = ARW_CH + n_blk * ARW_NLP + ARW_F;   // chunk + lag halo (+ last-lane slack)
    const size_t smem = (size_t)ARW_ST * ((3 * slab_len + 3) & ~3) * sizeof(float);
    const int chunks = (B + ARW_CH - 1) / ARW_CH;
    // one CTA of n_blk warps holds 127 registers per thread: size the atom split for a few
    // waves of the CTAs that fit
    int per_sm = 65536 / (128 * 32 * n_blk);
    if (per_sm < 1) per_sm = 1;
    long long want = ((long long)sm_count() * per_sm * 3 + chunks - 1) / chunks;
    if (want < 1) want = 1;
    long long apc = (n_atoms + want - 1) / want;
    if (apc < 32) apc = n_atoms < 32 ? n_atoms : 32;
    const long long groups = (n_atoms + apc - 1) / apc;
    MDK_CHECK_ARG(groups <= 65535, "acf_lagprod: too many atom groups");
    dim3 grid((unsigned)chunks, (unsigned)groups);
    cudaStream_t st = as_stream(stream);
    if (vec)
      acf_rw_kernel<true><<<grid, 32 * n_blk, smem, st>>>(traj, T, a_lo, a_hi, (int)apc, t0, B, N,
                                                          slab_len, P);
    else
      acf_rw_kernel<false><<<grid, 32 * n_blk, smem, st>>>(traj, T, a_lo, a_hi, (int)apc, t0, B, N,
                                                           slab_len, P);
    MDK_LAUNCH_CHECK();
    return MDK_OK;
  }
  if (N <= 16 && !getenv("MDK_ACF_NO_STREAM")) {
    // short lag ranges: HBM-streaming kernel (a warp per 128-origin chunk and atom slice)
    const bool vec = (T % 4 == 0) && (t0 % 4 == 0) && (reinterpret_cast<uintptr_t>(traj) % 16 == 0);
    const int chunks = (B + ACS_CH - 1) / ACS_CH;
    long long want = ((long long)sm_count() * 32 + chunks - 1) / chunks;  // CTAs over atoms
    if (want < 1) want = 1;
    long long apc = (n_atoms + want - 1) / want;
    if (apc < ACS_WARPS) apc = ACS_WARPS;
    const long long groups = (n_atoms + apc - 1) / apc;
    MDK_CHECK_ARG(groups <= 65535, "acf_lagprod: too many atom groups");
    dim3 grid((unsigned)chunks, (unsigned)groups);
    cudaStream_t st = as_stream(stream);
#define MDK_ACS_LAUNCH(NL)                                                                    \
  do {                                                                                        \
    if (vec)                                                                                  \
      acf_stream_kernel<NL, true><<<grid, 32 * ACS_WARPS, 0, st>>>(traj, T, a_lo, a_hi, (int)apc, \
                                                                   t0, B, N, P);              \
    else                                                                                      \
      acf_stream_kernel<NL, false><<<grid, 32 * ACS_WARPS, 0, st>>>(traj, T, a_lo, a_hi,      \
                                                                    (int)apc, t0, B, N, P);   \
  } while (0)
    if (N <= 4) MDK_ACS_LAUNCH(4);
    else if (N <= 8) MDK_ACS_LAUNCH(8);
    else MDK_ACS_LAUNCH(16);
#undef MDK_ACS_LAUNCH
    MDK_LAUNCH_CHECK();
    return MDK_OK;
  }
  // band Gram kernel: tiles (bi, bi + dj), dj < nbj
  const int n_bi = (B + GB_T - 1) / GB_T;
  const int nbj = (GB_T - 1 + N - 1) / GB_T + 1;
  const long long tiles = (long long)n_bi * nbj;
  // atom split: enough CTAs to fill the machine several times over, at least GB_KA atoms each
  long long want = ((long long)sm_count() * 96 + tiles - 1) / tiles;  // ~12 waves of 8 CTAs/SM
  if (want < 1) want = 1;
  long long apc = (n_atoms + want - 1) / want;
  if (apc < 64) apc = n_atoms < 64 ? n_atoms : 64;
  apc = ((apc + GB_KA - 1) / GB_KA) * GB_KA;
  const long long groups = (n_atoms + apc - 1) / apc;
  MDK_CHECK_ARG(groups <= 65535 && tiles < (1ll << 31), "acf_lagprod: grid too large");
  const size_t smem = (size_t)GB_ST * GB_KA * GB_SLAB * sizeof(float);
  MDK_CUDA(cudaFuncSetAttribute(acf_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid((unsigned)tiles, (unsigned)groups);
  acf_band_kernel<<<grid, GB_NT, smem, as_stream(stream)>>>(traj, T, a_lo, a_hi, (int)apc, t0, B, N,
                                                            nbj, P);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_acf_windows(double* P, int B, int N, int W, int ct, double* acf_sum,
                               double* acf_win, mdk_stream_t stream) {
  MDK_CHECK_ARG(P && acf_sum, "acf_windows: null pointer");
  MDK_CHECK_ARG(B >= 1 && N >= 1 && W >= 0 && ct >= 1, "acf_windows: bad argument");
  if (W == 0) return MDK_OK;
  MDK_CHECK_ARG((long long)(W - 1) * ct + N <= B, "acf_windows: windows exceed the batch");
  cudaStream_t s = as_stream(stream);
  acf_prefix_kernel<<<(N + SCAN_M - 1) / SCAN_M, SCAN_M * SCAN_C, 0, s>>>(P, B, N);
  MDK_LAUNCH_CHECK();
  dim3 grid((N + 127) / 128, (W + ACFW_WC - 1) / ACFW_WC);
  acf_windows_kernel<<<grid, 128, 0, s>>>(P, B, N, W, ct, acf_sum, acf_win);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}
