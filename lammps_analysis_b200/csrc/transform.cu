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

#include <type_traits>

namespace mdk {

// One warp per atom; each lane owns one frame of a 32-frame chunk.  The jump count is an
// inclusive warp scan per dimension, the running image is carried across chunks and batches.
// jump = rint((p_t - p_{t-1}) / L) is decided in fp32 (magic-number rounding, no conversion
// instructions) when the quotient is clearly away from a half-integer and recomputed exactly as
// the reference does (fp64 division, half-to-even) otherwise.  Jumps and images are small
// integers and are kept as floats (exact below 2^24), so the scan needs no int<->float
// conversions; the conversion (XU) pipe was the limiter of the first version of this kernel.
// Output: fl32(double(p) + img * L) -- with a single fp32 FMA when L is exactly representable in
// fp32 (then fmaf rounds the exact sum once, which is the same value), and skipped entirely
// while the image is zero.
__device__ __forceinline__ float jump_of(float p, float prev, float inv_l32, double l64) {
  const float q = (p - prev) * inv_l32;
  const float n = __fadd_rn(__fadd_rn(q, 12582912.0f), -12582912.0f);  // rint for |q| < 2^22
  if (fabsf(fabsf(q - n) - 0.5f) < 1e-3f || !(fabsf(q) < 1000.f)) {
    const double qq = ((double)p - (double)prev) / l64;
    return (float)rint(qq);
  }
  return n;
}

// The same decision for two (position, previous position) pairs on packed instructions.  The
// fast result stands when |q - rint(q)| is clearly below 1/2 (q = rint(q) + r with |r| <= 1/2,
// so "within 1e-3 of a half-integer" is |r| > 0.499) and |q| < 1000; otherwise bit 0 / bit 1 of
// the return value asks for the reference's fp64 division (jump_exact) for the first / second
// pair.
__device__ __forceinline__ unsigned jump_pair(float2 cur, float2 prev, float inv_l32, float& j0,
                                              float& j1) {
  const float2 q = __fmul2_rn(__fadd2_rn(cur, make_float2(-prev.x, -prev.y)),
                              make_float2(inv_l32, inv_l32));
  const float2 n = __fadd2_rn(__fadd2_rn(q, make_float2(12582912.0f, 12582912.0f)),
                              make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __fadd2_rn(q, make_float2(-n.x, -n.y));
  j0 = n.x;
  j1 = n.y;
  const unsigned s0 = (fabsf(r.x) > 0.499f || !(fabsf(q.x) < 1000.f)) ? 1u : 0u;
  const unsigned s1 = (fabsf(r.y) > 0.499f || !(fabsf(q.y) < 1000.f)) ? 2u : 0u;
  return s0 | s1;
}
__device__ __noinline__ float jump_exact(float p, float prev, double l64) {
  return (float)rint(((double)p - (double)prev) / l64);
}

template <bool L32>
__device__ __forceinline__ float shifted(float x, float m, double l64, float l32) {
  if (m == 0.f) return x;
  if (L32) return fmaf(m, l32, x);
  return __double2float_rn(__dadd_rn((double)x, __dmul_rn((double)m, l64)));
}

constexpr int UNW_F = 4;    // frames per lane and chunk: a warp moves 128 frames = 1536 contiguous bytes
constexpr int UNW_PF = 2;   // chunks kept in flight in registers
constexpr int UNW_CH = 32 * UNW_F;        // frames per chunk
constexpr int UNW_SLAB = 3 * UNW_CH;      // floats per chunk
constexpr int UNW_WARPS = 4;

// One warp per atom.  The warp streams the atom's row in chunks of 128 frames: the 384 floats are
// read (and later written) as three fully coalesced 512-byte LDG.128/STG.128 rows and
// re-distributed through a per-warp shared-memory slab so that lane l owns the 4 consecutive
// frames 4l .. 4l+3.  Jumps are scanned inside the lane, then across the warp.
template <bool L32, bool VEC, int MINB>
__global__ void __launch_bounds__(32 * UNW_WARPS, MINB)
unwrap_kernel(const float* __restrict__ pos, long long A, long long T, double lx, double ly,
              double lz, float* __restrict__ carry_pos, int have_carry,
              double* __restrict__ carry_img, float* __restrict__ out) {
  __shared__ __align__(16) float s_slab[UNW_WARPS][UNW_SLAB];
  const int lane = threadIdx.x & 31;
  const long long a = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (a >= A) return;
  float* __restrict__ wsm = s_slab[threadIdx.x >> 5];
  const float* __restrict__ src = pos + (size_t)a * T * 3;
  float* __restrict__ dst = out + (size_t)a * T * 3;
  const long long n_el = T * 3;
  const float il[3] = {(float)(1.0 / lx), (float)(1.0 / ly), (float)(1.0 / lz)};
  const double l64[3] = {lx, ly, lz};
  const float l32[3] = {(float)lx, (float)ly, (float)lz};

  // carry: previous position and image (first batch: p_{-1} = p_0, img = 0)
  float prevp[3], img[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    prevp[d] = have_carry ? carry_pos[a * 3 + d] : src[d];
    // images are integers, exact in fp32 below 2^24 box crossings
    img[d] = (float)carry_img[a * 3 + d];
  }

  // element e of a chunk is loaded by lane (e/4)%32 in row (e/128): 16-byte vectors when the
  // row start is 16-byte aligned (VEC: T % 4 == 0 and an aligned base), scalars otherwise.
  // FULL chunks (this chunk and the one being prefetched lie wholly inside the row, VEC) run
  // without any bounds arithmetic: the 64-bit index compares and selects of the general path were
  // 40 % of the instructions of this kernel, which is bound by instruction issue, not by HBM.
  auto load_chunk = [&](long long tb, float4 (&r)[3], auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const long long e = tb * 3 + j * 128 + 4 * lane;
      if (FULL || (VEC && e + 3 < n_el)) {
        r[j] = __ldg(reinterpret_cast<const float4*>(src + e));
      } else {
        r[j].x = e + 0 < n_el ? __ldg(src + e + 0) : 0.f;
        r[j].y = e + 1 < n_el ? __ldg(src + e + 1) : 0.f;
        r[j].z = e + 2 < n_el ? __ldg(src + e + 2) : 0.f;
        r[j].w = e + 3 < n_el ? __ldg(src + e + 3) : 0.f;
      }
    }
  };
  float4 pre[UNW_PF][3];
#pragma unroll
  for (int k = 0; k < UNW_PF; ++k) load_chunk((long long)k * UNW_CH, pre[k], std::false_type{});

  auto chunk = [&](long long tb, auto full_tag) {
    constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
    for (int j = 0; j < 3; ++j) *reinterpret_cast<float4*>(wsm + j * 128 + 4 * lane) = pre[0][j];
#pragma unroll
    for (int k = 0; k + 1 < UNW_PF; ++k)
#pragma unroll
      for (int j = 0; j < 3; ++j) pre[k][j] = pre[k + 1][j];
    load_chunk(tb + (long long)UNW_PF * UNW_CH, pre[UNW_PF - 1], full_tag);
    __syncwarp();
    // lane l owns frames tb + 4l .. tb + 4l + 3: floats [12 l, 12 l + 12) of the slab
    float p[UNW_F][3];
    {
      const float4 v0 = *reinterpret_cast<const float4*>(wsm + 12 * lane);
      const float4 v1 = *reinterpret_cast<const float4*>(wsm + 12 * lane + 4);
      const float4 v2 = *reinterpret_cast<const float4*>(wsm + 12 * lane + 8);
      p[0][0] = v0.x; p[0][1] = v0.y; p[0][2] = v0.z; p[1][0] = v0.w;
      p[1][1] = v1.x; p[1][2] = v1.y; p[2][0] = v1.z; p[2][1] = v1.w;
      p[2][2] = v2.x; p[3][0] = v2.y; p[3][1] = v2.z; p[3][2] = v2.w;
    }
    __syncwarp();
    const long long t_first = tb + UNW_F * lane;
    const int n_here = FULL ? UNW_F : (int)min((long long)UNW_F, max(0ll, T - t_first));
    const int last_lane = FULL ? 31 : (int)((min((long long)UNW_CH, T - tb) - 1) / UNW_F);
    const int last_f = FULL ? UNW_F - 1 : (int)((min((long long)UNW_CH, T - tb) - 1) % UNW_F);
    float o[UNW_F][3];
    float tot[3] = {0.f, 0.f, 0.f};  // jumps of the whole chunk
    static_assert(UNW_F == 4, "the packed jump test below pairs frames (0, 1) and (2, 3)");
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float q = __shfl_up_sync(0xffffffffu, p[UNW_F - 1][d], 1);
      if (lane == 0) q = prevp[d];
      // two frames per packed instruction: (p0, p1) - (q, p0) and (p2, p3) - (p1, p2); the rare
      // exact re-evaluations of the four frames share one branch
      float j4[UNW_F];
      const unsigned s01 = jump_pair(make_float2(p[0][d], p[1][d]), make_float2(q, p[0][d]),
                                     il[d], j4[0], j4[1]);
      const unsigned s23 = jump_pair(make_float2(p[2][d], p[3][d]),
                                     make_float2(p[1][d], p[2][d]), il[d], j4[2], j4[3]);
      if (s01 | s23) {
        if (s01 & 1u) j4[0] = jump_exact(p[0][d], q, l64[d]);
        if (s01 & 2u) j4[1] = jump_exact(p[1][d], p[0][d], l64[d]);
        if (s23 & 1u) j4[2] = jump_exact(p[2][d], p[1][d], l64[d]);
        if (s23 & 2u) j4[3] = jump_exact(p[3][d], p[2][d], l64[d]);
      }
      float jl[UNW_F];       // inclusive prefix of the jumps inside the lane
      float run = 0.f;
      bool jumped = false;   // any jump among this lane's frames (two may cancel in `run`)
#pragma unroll
      for (int f = 0; f < UNW_F; ++f) {
        const float jv = (FULL || f < n_here) ? j4[f] : 0.f;
        jumped |= jv != 0.f;
        run += jv;
        jl[f] = run;
      }
      // most chunks hold no jump in a given dimension: then the image is the same for all of
      // its frames, and the scan and the per-frame image arithmetic are skipped
      if (__any_sync(0xffffffffu, jumped)) {
        float inc = run;
#pragma unroll
        for (int s2 = 1; s2 < 32; s2 <<= 1) {
          const float u = __shfl_up_sync(0xffffffffu, inc, s2);
          if (lane >= s2) inc += u;
        }
        const float off = inc - run;   // exclusive scan of the lane totals
        tot[d] = __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
        for (int f = 0; f < UNW_F; ++f)
          o[f][d] = shifted<L32>(p[f][d], img[d] - (off + jl[f]), l64[d], l32[d]);
      } else if (img[d] != 0.f) {
#pragma unroll
        for (int f = 0; f < UNW_F; ++f) o[f][d] = shifted<L32>(p[f][d], img[d], l64[d], l32[d]);
      } else {
#pragma unroll
        for (int f = 0; f < UNW_F; ++f) o[f][d] = p[f][d];
      }
    }
    *reinterpret_cast<float4*>(wsm + 12 * lane) = make_float4(o[0][0], o[0][1], o[0][2], o[1][0]);
    *reinterpret_cast<float4*>(wsm + 12 * lane + 4) = make_float4(o[1][1], o[1][2], o[2][0], o[2][1]);
    *reinterpret_cast<float4*>(wsm + 12 * lane + 8) = make_float4(o[2][2], o[3][0], o[3][1], o[3][2]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const long long e = tb * 3 + j * 128 + 4 * lane;
      const float4 v = *reinterpret_cast<const float4*>(wsm + j * 128 + 4 * lane);
      if (FULL || (VEC && e + 3 < n_el)) {
        *reinterpret_cast<float4*>(dst + e) = v;
      } else {
        if (e + 0 < n_el) dst[e + 0] = v.x;
        if (e + 1 < n_el) dst[e + 1] = v.y;
        if (e + 2 < n_el) dst[e + 2] = v.z;
        if (e + 3 < n_el) dst[e + 3] = v.w;
      }
    }
    __syncwarp();
    // chunk carry: image after the chunk, last valid position
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      img[d] -= tot[d];
      float lp = p[0][d];
#pragma unroll
      for (int f = 1; f < UNW_F; ++f) lp = (f == last_f) ? p[f][d] : lp;
      if (last_f == 0) lp = p[0][d];
      prevp[d] = __shfl_sync(0xffffffffu, lp, last_lane);
    }
  };

  for (long long tb = 0; tb < T; tb += UNW_CH) {
    if (VEC && tb + (long long)(UNW_PF + 1) * UNW_CH <= T)
      chunk(tb, std::true_type{});
    else
      chunk(tb, std::false_type{});
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      carry_img[a * 3 + d] = (double)img[d];
      if (carry_pos) carry_pos[a * 3 + d] = prevp[d];
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

// v(t) = (x(t + 1) - x(t)) / dt in fp32, the last frame repeats the one before it
// (velocity_from_positions.py:62-77: roll, subtract, divide, drop and re-append the last value).
__global__ void velocity_from_positions_kernel(const float* __restrict__ pos, long long A,
                                               long long T, float dt, float* __restrict__ out) {
  const long long total = A * T * 3;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long t = (e / 3) % T;
    const long long src = (t == T - 1 && T > 1) ? e - 3 : e;   // last frame: copy of frame T - 2
    const float v = T > 1 ? __fdiv_rn(__fsub_rn(__ldg(pos + src + 3), __ldg(pos + src)), dt) : 0.f;
    out[e] = v;
  }
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_velocity_from_positions(const float* pos, long long A, long long T, float dt,
                                           float* out, mdk_stream_t stream) {
  MDK_CHECK_ARG(A >= 0 && T >= 0, "velocity_from_positions: negative size");
  if (A == 0 || T == 0) return MDK_OK;
  MDK_CHECK_ARG(pos && out, "velocity_from_positions: null pointer");
  MDK_CHECK_ARG(dt != 0.f, "velocity_from_positions: dt must not be zero");
  long long blocks = (A * T * 3 + 255) / 256;
  const long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  velocity_from_positions_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pos, A, T, dt,
                                                                                   out);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_unwrap(const float* pos, long long A, long long T, const double* box,
                          float* carry_pos, int have_carry, double* carry_img, float* out,
                          mdk_stream_t stream) {
  MDK_CHECK_ARG(pos && box && carry_img && out, "unwrap: null pointer");
  MDK_CHECK_ARG(A >= 0 && T >= 1, "unwrap: bad shape");
  MDK_CHECK_ARG(!have_carry || carry_pos, "unwrap: have_carry set but carry_pos is NULL");
  MDK_CHECK_ARG(box[0] > 0 && box[1] > 0 && box[2] > 0, "unwrap: box must be positive");
  if (A == 0) return MDK_OK;
  const long long blocks = (A + UNW_WARPS - 1) / UNW_WARPS;
  MDK_CHECK_ARG(blocks < (1ll << 31), "unwrap: too many atoms for one launch");
  const bool l32 = box[0] == (double)(float)box[0] && box[1] == (double)(float)box[1] &&
                   box[2] == (double)(float)box[2];
  // 16-byte vector access needs every atom row (T * 12 bytes apart) to start 16-byte aligned
  const bool vec = (T % 4 == 0) && (reinterpret_cast<uintptr_t>(pos) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const unsigned nb = (unsigned)blocks, nt = 32 * UNW_WARPS;
  cudaStream_t st = as_stream(stream);
  int minb = 0;  // tuning: resident CTAs per SM the register allocation is capped for
  if (const char* e = getenv("MDK_UNWRAP_MINB")) minb = atoi(e);
#define MDK_UNWRAP_LAUNCH(L, V)                                                               \
  do {                                                                                        \
    if (minb == 6)                                                                            \
      unwrap_kernel<L, V, 6><<<nb, nt, 0, st>>>(pos, A, T, box[0], box[1], box[2], carry_pos, \
                                                have_carry, carry_img, out);                  \
    else if (minb == 8)                                                                       \
      unwrap_kernel<L, V, 8><<<nb, nt, 0, st>>>(pos, A, T, box[0], box[1], box[2], carry_pos, \
                                                have_carry, carry_img, out);                  \
    else                                                                                      \
      unwrap_kernel<L, V, 0><<<nb, nt, 0, st>>>(pos, A, T, box[0], box[1], box[2], carry_pos, \
                                                have_carry, carry_img, out);                  \
  } while (0)
  if (l32 && vec) MDK_UNWRAP_LAUNCH(true, true);
  else if (l32) MDK_UNWRAP_LAUNCH(true, false);
  else if (vec) MDK_UNWRAP_LAUNCH(false, true);
  else MDK_UNWRAP_LAUNCH(false, false);
#undef MDK_UNWRAP_LAUNCH
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
