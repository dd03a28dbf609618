// Atom reductions behind the flux transformations: system observables
// J[t][k] = sum over atoms of a per-atom quantity, accumulated in fp64 and coalesced along
// (t, k) exactly like the ionic current (transform.cu).  HBM bound: every input element is
// read once.
//
// Replaces the transform_batch bodies of
//   transformations/momentum_flux.py:45-55           (stress components 3..5 summed over atoms)
//   transformations/integrated_heat_current.py:49-60 (sum_a r_a (KE_a + PE_a))
//   transformations/thermal_flux.py:51-92            (sum_a (KE_a + PE_a) v_a - S_a v_a)
#include "mdk_common.cuh"

namespace mdk {

// J[t][k] += sum_a w(a, t) * x[a][t][comp0 + k],  k < 3,  w = 1 or w1[a][t] (+ w2[a][t])
template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
flux_sum_kernel(const float* __restrict__ x, long long A, long long T, int ncomp, int comp0,
                const float* __restrict__ w1, const float* __restrict__ w2, int atoms_per_slice,
                double* __restrict__ J) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= 3 * T) return;
  const long long t = e / 3;
  const int k = (int)(e - 3 * t);
  const long long a0 = (long long)blockIdx.y * atoms_per_slice;
  const long long a1 = min(A, a0 + atoms_per_slice);
  double acc = 0.0;
  for (long long a = a0; a < a1; ++a) {
    const size_t at = (size_t)a * T + t;
    const double v = (double)__ldg(x + at * ncomp + comp0 + k);
    if (WEIGHTED) {
      double w = (double)__ldg(w1 + at);
      if (w2) w += (double)__ldg(w2 + at);
      acc += v * w;
    } else {
      acc += v;
    }
  }
  atomicAdd(J + e, acc);
}

// J[t][k] += sum_a ( (KE + PE) v_k - (S v)_k ),  S symmetric from the 6 LAMMPS stress components
// (xx, yy, zz, xy, xz, yz)
__global__ void __launch_bounds__(256)
thermal_flux_kernel(const float* __restrict__ stress, const float* __restrict__ vel,
                    const float* __restrict__ ke, const float* __restrict__ pe, long long A,
                    long long T, int atoms_per_slice, double* __restrict__ J) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= 3 * T) return;
  const long long t = e / 3;
  const int k = (int)(e - 3 * t);
  // row k of the symmetric stress matrix in the 6-component layout
  const int c0 = k == 0 ? 0 : (k == 1 ? 3 : 4);
  const int c1 = k == 0 ? 3 : (k == 1 ? 1 : 5);
  const int c2 = k == 0 ? 4 : (k == 1 ? 5 : 2);
  const long long a0 = (long long)blockIdx.y * atoms_per_slice;
  const long long a1 = min(A, a0 + atoms_per_slice);
  double ev = 0.0, phi = 0.0;
  for (long long a = a0; a < a1; ++a) {
    const size_t at = (size_t)a * T + t;
    const float* __restrict__ s = stress + at * 6;
    const float* __restrict__ v = vel + at * 3;
    const double v0 = (double)__ldg(v), v1 = (double)__ldg(v + 1), v2 = (double)__ldg(v + 2);
    const double vk = k == 0 ? v0 : (k == 1 ? v1 : v2);
    phi += (double)__ldg(s + c0) * v0 + (double)__ldg(s + c1) * v1 + (double)__ldg(s + c2) * v2;
    ev += ((double)__ldg(ke + at) + (double)__ldg(pe + at)) * vk;
  }
  atomicAdd(J + e, ev - phi);
}

static inline int slice_atoms(long long A, long long T, dim3* grid) {
  const long long xblocks = (3 * T + 255) / 256;
  long long slices = ((long long)sm_count() * 16 + xblocks - 1) / xblocks;
  if (slices < 1) slices = 1;
  if (slices > A) slices = A;
  if (slices > 65535) slices = 65535;
  const int aps = (int)((A + slices - 1) / slices);
  slices = (A + aps - 1) / aps;
  *grid = dim3((unsigned)xblocks, (unsigned)slices);
  return aps;
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_flux_sum(const float* x, long long A, long long T, int ncomp, int comp0,
                            const float* w1, const float* w2, double* J, mdk_stream_t stream) {
  MDK_CHECK_ARG(x && J, "flux_sum: null pointer");
  MDK_CHECK_ARG(A >= 0 && T >= 1 && ncomp >= 3 && comp0 >= 0 && comp0 + 3 <= ncomp,
                "flux_sum: bad shape (A=%lld, T=%lld, ncomp=%d, comp0=%d)", A, T, ncomp, comp0);
  MDK_CHECK_ARG(w1 || !w2, "flux_sum: w2 given without w1");
  if (A == 0) return MDK_OK;
  dim3 grid;
  const int aps = slice_atoms(A, T, &grid);
  if (w1)
    flux_sum_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(x, A, T, ncomp, comp0, w1, w2, aps, J);
  else
    flux_sum_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(x, A, T, ncomp, comp0, nullptr,
                                                                nullptr, aps, J);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_thermal_flux(const float* stress, const float* vel, const float* ke,
                                const float* pe, long long A, long long T, double* J,
                                mdk_stream_t stream) {
  MDK_CHECK_ARG(stress && vel && ke && pe && J, "thermal_flux: null pointer");
  MDK_CHECK_ARG(A >= 0 && T >= 1, "thermal_flux: bad shape");
  if (A == 0) return MDK_OK;
  dim3 grid;
  const int aps = slice_atoms(A, T, &grid);
  thermal_flux_kernel<<<grid, 256, 0, as_stream(stream)>>>(stress, vel, ke, pe, A, T, aps, J);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}
