// Transformations on the hot path: coordinate unwrap (segmented scan along time) and the
// ionic current (charge-weighted reduction over atoms).
//
// Replaces
//   transformations/unwrap_coordinates.py:51-81   diff -> round -> cumsum -> pos + img * L (fp64)
//   transformations/unwrap_via_indices.py:49-57   pos + img * L
//   transformations/ionic_current.py:48-58        sum_a q * v, summed over species
// The persisted result of every transformation is float32 (simulation_database.py:491-497),
// so the kernels form the fp64 value the reference forms and round once to fp32.
#include "mdk_common.cuh"

namespace mdk {

// One warp per atom; each lane owns one frame of a 32-frame chunk.  The jump count is an
// inclusive warp scan per dimension, the running image is carried across chunks in fp64.
// jump = rint((p_t - p_{t-1}) / L) is decided in fp32 when the quotient is clearly away
// from a half-integer and recomputed exactly as the reference does (fp64 division,
// half-to-even) otherwise.
__device__ __forceinline__ int jump_of(float p, float prev, float inv_l32, double l64) {
  const float q = (p - prev) * inv_l32;
  const float n = rintf(q);
  if (fabsf(fabsf(q - n) - 0.5f) < 1e-3f || !(fabsf(q) < 1000.f)) {
    const double qq = ((double)p - (double)prev) / l64;
    return (int)rint(qq);
  }
  return (int)n;
}

__global__ void __launch_bounds__(256)
unwrap_kernel(const float* __restrict__ pos, long long A, long long T, double lx, double ly,
              double lz, float* __restrict__ carry_pos, int have_carry,
              double* __restrict__ carry_img, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long a = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (a >= A) return;
  const float* __restrict__ src = pos + (size_t)a * T * 3;
  float* __restrict__ dst = out + (size_t)a * T * 3;
  const float ilx = (float)(1.0 / lx), ily = (float)(1.0 / ly), ilz = (float)(1.0 / lz);

  // carry: previous position and image (first batch: p_{-1} = p_0, img = 0)
  float px, py, pz;
  if (have_carry) {
    px = carry_pos[a * 3 + 0];
    py = carry_pos[a * 3 + 1];
    pz = carry_pos[a * 3 + 2];
  } else {
    px = src[0];
    py = src[1];
    pz = src[2];
  }
  double ix = carry_img[a * 3 + 0], iy = carry_img[a * 3 + 1], iz = carry_img[a * 3 + 2];

  float nx = 0.f, ny = 0.f, nz = 0.f;  // prefetched chunk
  bool nvalid = lane < T;
  if (nvalid) {
    nx = __ldg(src + 3 * lane);
    ny = __ldg(src + 3 * lane + 1);
    nz = __ldg(src + 3 * lane + 2);
  }
  for (long long tb = 0; tb < T; tb += 32) {
    const long long t = tb + lane;
    const bool valid = nvalid;
    const float x = nx, y = ny, z = nz;
    const long long tn = t + 32;
    nvalid = tn < T;
    if (nvalid) {
      nx = __ldg(src + 3 * tn);
      ny = __ldg(src + 3 * tn + 1);
      nz = __ldg(src + 3 * tn + 2);
    }
    float qx = __shfl_up_sync(0xffffffffu, x, 1);
    float qy = __shfl_up_sync(0xffffffffu, y, 1);
    float qz = __shfl_up_sync(0xffffffffu, z, 1);
    if (lane == 0) {
      qx = px;
      qy = py;
      qz = pz;
    }
    int jx = 0, jy = 0, jz = 0;
    if (valid) {
      jx = jump_of(x, qx, ilx, lx);
      jy = jump_of(y, qy, ily, ly);
      jz = jump_of(z, qz, ilz, lz);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ux = __shfl_up_sync(0xffffffffu, jx, o);
      const int uy = __shfl_up_sync(0xffffffffu, jy, o);
      const int uz = __shfl_up_sync(0xffffffffu, jz, o);
      if (lane >= o) {
        jx += ux;
        jy += uy;
        jz += uz;
      }
    }
    if (valid) {
      const double mx = ix - (double)jx, my = iy - (double)jy, mz = iz - (double)jz;
      dst[3 * t + 0] = __double2float_rn(__dadd_rn((double)x, __dmul_rn(mx, lx)));
      dst[3 * t + 1] = __double2float_rn(__dadd_rn((double)y, __dmul_rn(my, ly)));
      dst[3 * t + 2] = __double2float_rn(__dadd_rn((double)z, __dmul_rn(mz, lz)));
    }
    // chunk carry: last valid lane of this chunk
    const int last = (int)min((long long)31, T - 1 - tb);
    ix -= (double)__shfl_sync(0xffffffffu, jx, last);
    iy -= (double)__shfl_sync(0xffffffffu, jy, last);
    iz -= (double)__shfl_sync(0xffffffffu, jz, last);
    px = __shfl_sync(0xffffffffu, x, last);
    py = __shfl_sync(0xffffffffu, y, last);
    pz = __shfl_sync(0xffffffffu, z, last);
  }
  if (lane == 0) {
    carry_img[a * 3 + 0] = ix;
    carry_img[a * 3 + 1] = iy;
    carry_img[a * 3 + 2] = iz;
    if (carry_pos) {
      carry_pos[a * 3 + 0] = px;
      carry_pos[a * 3 + 1] = py;
      carry_pos[a * 3 + 2] = pz;
    }
  }
}

__global__ void __launch_bounds__(256)
unwrap_indices_kernel(const float* __restrict__ pos, const float* __restrict__ img,
                      long long n3, double lx, double ly, double lz, float* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n3;
       e += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(e % 3);
    const double l = d == 0 ? lx : (d == 1 ? ly : lz);
    out[e] = __double2float_rn(__dadd_rn((double)pos[e], __dmul_rn((double)img[e], l)));
  }
}

// J[e] += sum_{a in slice} q_a * v[a][e], e = 3*t + d (contiguous, coalesced).
// grid.x: element blocks, grid.y: atom slices.
template <int QMODE>
__global__ void __launch_bounds__(256)
ionic_current_kernel(const float* __restrict__ vel, long long A, long long T3, double q_scalar,
                     const float* __restrict__ q, int atoms_per_slice, double* __restrict__ J) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= T3) return;
  const long long a0 = (long long)blockIdx.y * atoms_per_slice;
  const long long a1 = min(A, a0 + atoms_per_slice);
  double acc = 0.0;
  const float* __restrict__ p = vel + (size_t)a0 * T3 + e;
  long long a = a0;
  for (; a + 4 <= a1; a += 4) {
    const float v0 = __ldg(p), v1 = __ldg(p + T3), v2 = __ldg(p + 2 * T3), v3 = __ldg(p + 3 * T3);
    double c0, c1, c2, c3;
    if (QMODE == 0) {
      c0 = c1 = c2 = c3 = q_scalar;
    } else if (QMODE == 1) {
      c0 = q[a]; c1 = q[a + 1]; c2 = q[a + 2]; c3 = q[a + 3];
    } else {
      const long long T = T3 / 3, t = e / 3;
      c0 = q[a * T + t]; c1 = q[(a + 1) * T + t]; c2 = q[(a + 2) * T + t]; c3 = q[(a + 3) * T + t];
    }
    acc += c0 * (double)v0;
    acc += c1 * (double)v1;
    acc += c2 * (double)v2;
    acc += c3 * (double)v3;
    p += 4 * T3;
  }
  for (; a < a1; ++a) {
    double c;
    if (QMODE == 0) c = q_scalar;
    else if (QMODE == 1) c = q[a];
    else c = q[a * (T3 / 3) + e / 3];
    acc += c * (double)__ldg(p);
    p += T3;
  }
  atomicAdd(J + e, acc);
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_unwrap(const float* pos, long long A, long long T, const double* box,
                          float* carry_pos, int have_carry, double* carry_img, float* out,
                          mdk_stream_t stream) {
  MDK_CHECK_ARG(pos && box && carry_img && out, "unwrap: null pointer");
  MDK_CHECK_ARG(A >= 0 && T >= 1, "unwrap: bad shape");
  MDK_CHECK_ARG(!have_carry || carry_pos, "unwrap: have_carry set but carry_pos is NULL");
  MDK_CHECK_ARG(box[0] > 0 && box[1] > 0 && box[2] > 0, "unwrap: box must be positive");
  if (A == 0) return MDK_OK;
  const long long threads = A * 32;
  const long long blocks = (threads + 255) / 256;
  MDK_CHECK_ARG(blocks < (1ll << 31), "unwrap: too many atoms for one launch");
  unwrap_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      pos, A, T, box[0], box[1], box[2], carry_pos, have_carry, carry_img, out);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_unwrap_indices(const float* pos, const float* img, long long n_atom_frames,
                                  const double* box, float* out, mdk_stream_t stream) {
  MDK_CHECK_ARG(pos && img && box && out && n_atom_frames >= 0, "unwrap_indices: bad argument");
  if (n_atom_frames == 0) return MDK_OK;
  const long long n3 = n_atom_frames * 3;
  long long blocks = (n3 + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  unwrap_indices_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pos, img, n3, box[0],
                                                                         box[1], box[2], out);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_ionic_current(const float* vel, long long A, long long T, const void* q,
                                 int q_mode, double* J, mdk_stream_t stream) {
  MDK_CHECK_ARG(vel && q && J, "ionic_current: null pointer");
  MDK_CHECK_ARG(A >= 0 && T >= 1 && q_mode >= 0 && q_mode <= 2, "ionic_current: bad argument");
  if (A == 0) return MDK_OK;
  const long long T3 = T * 3;
  const long long xblocks = (T3 + 255) / 256;
  long long slices = ((long long)sm_count() * 16 + xblocks - 1) / xblocks;
  if (slices < 1) slices = 1;
  if (slices > A) slices = A;
  if (slices > 65535) slices = 65535;
  const int aps = (int)((A + slices - 1) / slices);
  slices = (A + aps - 1) / aps;
  dim3 grid((unsigned)xblocks, (unsigned)slices);
  cudaStream_t s = as_stream(stream);
  if (q_mode == 0)
    ionic_current_kernel<0><<<grid, 256, 0, s>>>(vel, A, T3, *static_cast<const double*>(q),
                                                 nullptr, aps, J);
  else if (q_mode == 1)
    ionic_current_kernel<1><<<grid, 256, 0, s>>>(vel, A, T3, 0.0, static_cast<const float*>(q), aps, J);
  else
    ionic_current_kernel<2><<<grid, 256, 0, s>>>(vel, A, T3, 0.0, static_cast<const float*>(q), aps, J);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}
