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

static int pick_atoms_per_cta(long long n_atoms, long long chunks) {
  // aim for ~8 CTAs per SM overall while keeping >= 1 atom per CTA
  const long long target = (long long)sm_count() * 8;
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

extern "C" int mdk_acf_lagprod(const float* traj, long long A, long long T, long long a_lo,
                               long long a_hi, long long t0, int B, int N, double* P,
                               mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && P, "acf_lagprod: null pointer");
  MDK_CHECK_ARG(0 <= a_lo && a_lo <= a_hi && a_hi <= A, "acf_lagprod: bad atom range");
  MDK_CHECK_ARG(B >= 1 && N >= 1 && t0 >= 0 && t0 + B <= T, "acf_lagprod: bad frame range");
  if (a_lo == a_hi) return MDK_OK;
  const size_t smem = (size_t)(ACF_TC + N - 1) * 12;
  if (smem > 200 * 1024) {
    set_error("acf_lagprod: data_range %d does not fit in shared memory", N);
    return MDK_EUNSUPPORTED;
  }
  const int chunks = (B + ACF_TC - 1) / ACF_TC;
  const int apc = pick_atoms_per_cta(a_hi - a_lo, chunks);
  const long long groups = (a_hi - a_lo + apc - 1) / apc;
  MDK_CHECK_ARG(groups <= 65535, "acf_lagprod: too many atom groups");
  MDK_CUDA(cudaFuncSetAttribute(acf_lagprod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  dim3 grid(chunks, (unsigned)groups);
  acf_lagprod_kernel<<<grid, DYN_NT, smem, as_stream(stream)>>>(traj, T, a_lo, a_hi, apc, t0, B, N,
                                                                P);
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
