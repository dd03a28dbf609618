// Angular distribution function: neighbour search + triplet-angle histograms.
//
// Replaces the TF op chain of the reference
//   utils/neighbour_list.py:53-112   get_neighbour_list  (all n(n-1)/2 minimum-image vectors)
//   calculators/angular_distribution_function.py:302-328  (dense n x n x 3 r_ij matrix)
//   utils/neighbour_list.py:116-177  get_triplets (n x n x n roll-and-compare in float16)
//   utils/linalg.py:30-81            get_angles (unit vectors, dot, acos, |r_ij||r_ik|)
//   calculators/angular_distribution_function.py:365-403  species masks + np.histogram
// which is O(n^2) memory and O(n^3) work, with a cell-list pass that is O(n * neighbours^2):
//   1. atoms of every frame are binned into a periodic cell grid (cell edge >= cutoff):
//      adf_cell_count -> exclusive scan (CUB) -> adf_cell_fill (positions + species, cell order);
//   2. adf_triplet_kernel: one WARP per centre atom gathers its neighbours from the 27 (or
//      fewer, for small boxes) stencil cells into shared memory -- minimum image
//      r - rint(r / L) * L with the reference's fp32 rounding sequence, |r| by a correctly
//      rounded sqrt, and the reference's FLOAT16 cutoff test half(|r|) < half(r_cut) with
//      zero distances excluded -- then walks all ordered neighbour pairs (j, k), j != k:
//      cos = u_ij . u_ik (products and sums rounded separately), clipped, angle = acos in fp64
//      rounded to fp32, weight 1 / (|r_ij| |r_ik|)^p, bin by numpy's uniform-bin rule on fp32
//      edges.  Only species triples with s_i <= s_j <= s_k are histogrammed (the reference's
//      combinations_with_replacement over (centre, j, k), :380).
//   3. counts (u32) and weights (fp32) go to CTA-private shared-memory histograms, flushed to
//      global u64 / fp64 tables when the CTA retires.
// No tensor cores: nothing here is a contraction.
#include "mdk_common.cuh"

#include <cuda_fp16.h>

#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstring>

namespace mdk {

constexpr int ADF_WARPS = 8;            // centre atoms in flight per CTA
constexpr int ADF_MAX_CELLS_DIM = 128;

struct AdfParams {
  const float* pos;              // [F][N][3]
  long long n_atoms;
  int n_frames;
  int n_species;
  int sp_hi[MDK_MAX_SPECIES];    // exclusive end of each species block in the concatenated order
  float box[3];
  int nc[3];                     // cells per dimension
  long long cells_per_frame;
  float rc16;                    // float(half(r_cut)): the cutoff as the reference compares it
  int nbins;
  float range_hi;                // last bin edge as fp32
  double step;                   // (hi - lo) / nbins in fp64: edge[k] = float(k * step)
  float inv_width;               // nbins / range_hi
  double norm_power;
  int capacity;                  // neighbours per centre held in shared memory
  int combo_of[MDK_MAX_SPECIES * MDK_MAX_SPECIES * MDK_MAX_SPECIES];  // -1: not histogrammed
  int n_combos;
  int smem_hist;                 // 1: CTA-private histograms in shared memory
  // workspace
  unsigned int* cell_count;      // [F * cells + 1] counts, then exclusive offsets (in place)
  unsigned int* cell_fill;       // [F * cells]
  float4* sorted;                // [F][N] (x, y, z, species as int bits) in cell order
  // outputs
  double* hist_w;                // [n_combos][nbins] +=
  unsigned long long* hist_c;    // [n_combos][nbins] +=
  int* overflow;                 // set to the largest neighbour count that exceeded capacity
};

__device__ __forceinline__ int cell_coord(float x, float L, int n) {
  // wrapped fractional coordinate in [0, 1): coordinates may lie outside the box
  float u = x / L;
  u -= floorf(u);
  int c = static_cast<int>(u * static_cast<float>(n));
  return c >= n ? n - 1 : (c < 0 ? 0 : c);
}

__device__ __forceinline__ long long cell_of(const AdfParams& P, float x, float y, float z) {
  const int cx = cell_coord(x, P.box[0], P.nc[0]);
  const int cy = cell_coord(y, P.box[1], P.nc[1]);
  const int cz = cell_coord(z, P.box[2], P.nc[2]);
  return (static_cast<long long>(cz) * P.nc[1] + cy) * P.nc[0] + cx;
}

__global__ void adf_cell_count_kernel(const __grid_constant__ AdfParams P) {
  const long long total = P.n_atoms * P.n_frames;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long f = e / P.n_atoms;
    const float* p = P.pos + 3 * e;
    const long long c = f * P.cells_per_frame + cell_of(P, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    atomicAdd(P.cell_count + c, 1u);
  }
}

__global__ void adf_cell_fill_kernel(const __grid_constant__ AdfParams P) {
  const long long total = P.n_atoms * P.n_frames;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long f = e / P.n_atoms;
    const int a = static_cast<int>(e - f * P.n_atoms);
    const float* p = P.pos + 3 * e;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const long long c = f * P.cells_per_frame + cell_of(P, x, y, z);
    const unsigned slot = P.cell_count[c] + atomicAdd(P.cell_fill + c, 1u);  // offsets after scan
    int s = 0;
    while (s + 1 < P.n_species && a >= P.sp_hi[s]) ++s;
    P.sorted[slot] = make_float4(x, y, z, __int_as_float(s));
  }
}

// numpy.histogram's uniform-bin rule (lib/_histograms_impl.py) on fp32 edges: the bin k with
// edge[k] <= a < edge[k + 1], the last bin closed on the right.
__device__ __forceinline__ int adf_bin(float a, const float* __restrict__ edge, int nbins,
                                       float inv_width) {
  int k = static_cast<int>(a * inv_width);
  k = k < 0 ? 0 : (k > nbins - 1 ? nbins - 1 : k);
  while (k > 0 && a < edge[k]) --k;
  while (k < nbins - 1 && a >= edge[k + 1]) ++k;
  return k;
}

__global__ void __launch_bounds__(ADF_WARPS * 32)
adf_triplet_kernel(const __grid_constant__ AdfParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: edges [nbins + 1] | neighbour slabs [ADF_WARPS][capacity] float4 (ux, uy, uz, d) |
  // species [ADF_WARPS][capacity] (u8, padded) | hist_w [n_combos][nbins] f32 | hist_c u32
  float* s_edge = reinterpret_cast<float*>(smem_raw);
  const int edge_len = (P.nbins + 1 + 3) & ~3;
  float4* s_nb = reinterpret_cast<float4*>(s_edge + edge_len);
  unsigned char* s_sp = reinterpret_cast<unsigned char*>(s_nb + (size_t)ADF_WARPS * P.capacity);
  const int sp_len = (ADF_WARPS * P.capacity + 15) & ~15;
  float* s_hw = reinterpret_cast<float*>(s_sp + sp_len);
  unsigned int* s_hc = reinterpret_cast<unsigned int*>(s_hw + (size_t)P.n_combos * P.nbins);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k <= P.nbins; k += blockDim.x)
    s_edge[k] = k == P.nbins ? P.range_hi : static_cast<float>(static_cast<double>(k) * P.step);
  if (P.smem_hist)
    for (int k = tid; k < P.n_combos * P.nbins; k += blockDim.x) {
      s_hw[k] = 0.f;
      s_hc[k] = 0u;
    }
  __syncthreads();

  float4* nb = s_nb + (size_t)warp * P.capacity;
  unsigned char* nsp = s_sp + (size_t)warp * P.capacity;
  const long long total = P.n_atoms * P.n_frames;
  const bool int_power = P.norm_power == 4.0;

  for (long long e = (long long)blockIdx.x * ADF_WARPS + warp; e < total;
       e += (long long)gridDim.x * ADF_WARPS) {
    const long long f = e / P.n_atoms;
    const float4 ci = P.sorted[e];                 // centres in cell order
    const int si = __float_as_int(ci.w);
    const unsigned int* cstart = P.cell_count + f * P.cells_per_frame;
    const float4* fsorted = P.sorted;              // offsets in cell_count are global
    const int cx = cell_coord(ci.x, P.box[0], P.nc[0]);
    const int cy = cell_coord(ci.y, P.box[1], P.nc[1]);
    const int cz = cell_coord(ci.z, P.box[2], P.nc[2]);
    // stencil: the three periodic neighbours per dimension, or every cell when fewer than three
    const int nx = P.nc[0] >= 3 ? 3 : P.nc[0], ny = P.nc[1] >= 3 ? 3 : P.nc[1],
              nz = P.nc[2] >= 3 ? 3 : P.nc[2];
    int n_found = 0;
    for (int s = 0; s < nx * ny * nz; ++s) {
      const int ox = s % nx, oy = (s / nx) % ny, oz = s / (nx * ny);
      const int jx = P.nc[0] >= 3 ? (cx + ox - 1 + P.nc[0]) % P.nc[0] : ox;
      const int jy = P.nc[1] >= 3 ? (cy + oy - 1 + P.nc[1]) % P.nc[1] : oy;
      const int jz = P.nc[2] >= 3 ? (cz + oz - 1 + P.nc[2]) % P.nc[2] : oz;
      const long long c = (static_cast<long long>(jz) * P.nc[1] + jy) * P.nc[0] + jx;
      const unsigned lo = cstart[c], hi = cstart[c + 1];
      for (unsigned base = lo; base < hi; base += 32) {
        const unsigned j = base + lane;
        bool near = false;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        int sj = 0;
        if (j < hi) {
          const float4 pj = fsorted[j];
          sj = __float_as_int(pj.w);
          // r = p_i - p_j; r -= rint(r / L) * L, every operation rounded (neighbour_list.py:88-94)
          float rx = __fsub_rn(ci.x, pj.x), ry = __fsub_rn(ci.y, pj.y), rz = __fsub_rn(ci.z, pj.z);
          rx = __fsub_rn(rx, __fmul_rn(rintf(__fdiv_rn(rx, P.box[0])), P.box[0]));
          ry = __fsub_rn(ry, __fmul_rn(rintf(__fdiv_rn(ry, P.box[1])), P.box[1]));
          rz = __fsub_rn(rz, __fmul_rn(rintf(__fdiv_rn(rz, P.box[2])), P.box[2]));
          const float d2 =
              __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
          const float d = __fsqrt_rn(d2);
          // float16 comparison of the reference (neighbour_list.py:150-152, :166)
          const float h = __half2float(__float2half_rn(d));
          near = (h != 0.f) && (h < P.rc16);
          if (near)
            out = make_float4(__fdiv_rn(rx, d), __fdiv_rn(ry, d), __fdiv_rn(rz, d), d);
        }
        const unsigned m = __ballot_sync(0xffffffffu, near);
        if (near) {
          const int slot = n_found + __popc(m & ((1u << lane) - 1u));
          if (slot < P.capacity) {
            nb[slot] = out;
            nsp[slot] = static_cast<unsigned char>(sj);
          }
        }
        n_found += __popc(m);
      }
    }
    __syncwarp();
    if (n_found > P.capacity) {
      if (lane == 0) atomicMax(P.overflow, n_found);
      continue;  // the host repeats the batch with a larger capacity
    }
    // ---- all ordered neighbour pairs (j, k), j != k ------------------------------------
    const int n = n_found;
    const int n_pairs = n * (n - 1);
    for (int q = lane; q < n_pairs; q += 32) {
      const int j = q / (n - 1);
      int k = q - j * (n - 1);
      k += (k >= j);
      const int sj = nsp[j], sk = nsp[k];
      const int combo = P.combo_of[(si * MDK_MAX_SPECIES + sj) * MDK_MAX_SPECIES + sk];
      if (combo < 0) continue;
      const float4 a = nb[j], b = nb[k];
      float c = __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
      c = fminf(fmaxf(c, -1.0f), 1.0f);
      const float ang = static_cast<float>(acos(static_cast<double>(c)));
      if (!(ang >= 0.f && ang <= P.range_hi)) continue;  // outside the histogram range (or NaN)
      const float pre = __fmul_rn(a.w, b.w);
      float pw;
      if (int_power) {
        const double p2 = static_cast<double>(pre) * static_cast<double>(pre);  // exact
        pw = static_cast<float>(p2 * p2);
      } else {
        pw = static_cast<float>(pow(static_cast<double>(pre), P.norm_power));
      }
      const float w = __fdiv_rn(1.0f, pw);
      const int bin = adf_bin(ang, s_edge, P.nbins, P.inv_width);
      const int idx = combo * P.nbins + bin;
      if (P.smem_hist) {
        atomicAdd(s_hw + idx, w);
        atomicAdd(s_hc + idx, 1u);
      } else {
        atomicAdd(P.hist_w + idx, static_cast<double>(w));
        atomicAdd(P.hist_c + idx, 1ull);
      }
    }
    __syncwarp();
  }
  if (P.smem_hist) {
    __syncthreads();
    for (int k = tid; k < P.n_combos * P.nbins; k += blockDim.x) {
      const unsigned cnt = s_hc[k];
      if (cnt) {
        atomicAdd(P.hist_c + k, static_cast<unsigned long long>(cnt));
        atomicAdd(P.hist_w + k, static_cast<double>(s_hw[k]));
      }
    }
  }
}

static void adf_grid(const float* box, float cutoff, int nc[3]) {
  // cell edge >= cutoff with a margin that covers the fp32 rounding of the wrapped coordinate
  const double edge = static_cast<double>(cutoff) * 1.001 + 1e-4;
  for (int d = 0; d < 3; ++d) {
    int n = static_cast<int>(std::floor(static_cast<double>(box[d]) / edge));
    if (n < 1) n = 1;
    if (n > ADF_MAX_CELLS_DIM) n = ADF_MAX_CELLS_DIM;
    nc[d] = n;
  }
}

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace mdk

using namespace mdk;

extern "C" long long mdk_adf_workspace(long long n_atoms, int n_frames, const float* box,
                                       float cutoff) {
  if (!box || n_atoms < 0 || n_frames < 0 || !(cutoff > 0.f)) return -1;
  int nc[3];
  adf_grid(box, cutoff, nc);
  const size_t cells = (size_t)nc[0] * nc[1] * nc[2] * (size_t)n_frames;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned int*)nullptr,
                                (unsigned int*)nullptr, (int)(cells + 1));
  return (long long)(align256((cells + 1) * 4) + align256(cells * 4) +
                     align256((size_t)n_atoms * n_frames * 16) + align256(scan_bytes) + 256);
}

extern "C" int mdk_adf_hist(const float* pos, int n_frames, long long n_atoms, const int* sp_hi,
                            int n_species, const float* box, float cutoff, int nbins,
                            double range_hi, double norm_power, int capacity, double* hist_w,
                            unsigned long long* hist_c, int* overflow, void* workspace,
                            long long workspace_bytes, mdk_stream_t stream) {
  MDK_CHECK_ARG(pos && sp_hi && box && hist_w && hist_c && overflow && workspace,
                "adf_hist: null pointer");
  MDK_CHECK_ARG(n_species >= 1 && n_species <= MDK_MAX_SPECIES,
                "adf_hist: n_species %d outside [1, %d]", n_species, MDK_MAX_SPECIES);
  MDK_CHECK_ARG(nbins >= 1 && cutoff > 0.f && range_hi > 0.0 && n_frames >= 0 && n_atoms >= 0,
                "adf_hist: bad sizes");
  MDK_CHECK_ARG(capacity >= 1 && capacity <= 8192, "adf_hist: capacity %d outside [1, 8192]",
                capacity);
  MDK_CHECK_ARG(n_atoms * (long long)n_frames < (1ll << 31), "adf_hist: batch too large");
  for (int s = 0; s < n_species; ++s)
    MDK_CHECK_ARG(sp_hi[s] >= (s ? sp_hi[s - 1] : 0) && sp_hi[s] <= n_atoms &&
                      box[s % 3] > 0.f, "adf_hist: bad species blocks or box");
  MDK_CHECK_ARG(sp_hi[n_species - 1] == n_atoms, "adf_hist: species blocks must cover all atoms");
  if (n_frames == 0 || n_atoms == 0) return MDK_OK;
  MDK_CHECK_ARG(workspace_bytes >= mdk_adf_workspace(n_atoms, n_frames, box, cutoff),
                "adf_hist: workspace too small");
  cudaStream_t s = as_stream(stream);

  AdfParams P;
  memset(&P, 0, sizeof(P));
  P.pos = pos;
  P.n_atoms = n_atoms;
  P.n_frames = n_frames;
  P.n_species = n_species;
  for (int q = 0; q < n_species; ++q) P.sp_hi[q] = sp_hi[q];
  for (int d = 0; d < 3; ++d) P.box[d] = box[d];
  adf_grid(box, cutoff, P.nc);
  P.cells_per_frame = (long long)P.nc[0] * P.nc[1] * P.nc[2];
  P.rc16 = __half2float(__float2half_rn(cutoff));
  P.nbins = nbins;
  // numpy.linspace(0, range_hi, nbins + 1) in fp64, cast to fp32 (the dtype numpy.histogram
  // picks for float32 samples): edge[k] = float(k * step), edge[nbins] = float(range_hi)
  P.range_hi = static_cast<float>(range_hi);
  P.step = range_hi / static_cast<double>(nbins);
  P.inv_width = static_cast<float>(static_cast<double>(nbins) / range_hi);
  P.norm_power = norm_power;
  P.capacity = capacity;
  // combinations_with_replacement(species, 3) in order: (centre, j, k) with a <= b <= c
  int n_combos = 0;
  for (int q = 0; q < MDK_MAX_SPECIES * MDK_MAX_SPECIES * MDK_MAX_SPECIES; ++q) P.combo_of[q] = -1;
  for (int a = 0; a < n_species; ++a)
    for (int b = a; b < n_species; ++b)
      for (int c = b; c < n_species; ++c)
        P.combo_of[(a * MDK_MAX_SPECIES + b) * MDK_MAX_SPECIES + c] = n_combos++;
  P.n_combos = n_combos;
  P.hist_w = hist_w;
  P.hist_c = hist_c;
  P.overflow = overflow;

  const size_t cells = (size_t)P.cells_per_frame * n_frames;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  P.cell_count = reinterpret_cast<unsigned int*>(w);
  w += align256((cells + 1) * 4);
  P.cell_fill = reinterpret_cast<unsigned int*>(w);
  w += align256(cells * 4);
  P.sorted = reinterpret_cast<float4*>(w);
  w += align256((size_t)n_atoms * n_frames * 16);
  void* scan_tmp = w;
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, P.cell_count, P.cell_count, (int)(cells + 1));

  MDK_CUDA(cudaMemsetAsync(P.cell_count, 0, (cells + 1) * 4, s));
  MDK_CUDA(cudaMemsetAsync(P.cell_fill, 0, cells * 4, s));
  const long long total = n_atoms * n_frames;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  adf_cell_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(P);
  MDK_LAUNCH_CHECK();
  MDK_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, P.cell_count, P.cell_count,
                                         (int)(cells + 1), s));
  adf_cell_fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(P);
  MDK_LAUNCH_CHECK();

  const size_t hist_bytes = (size_t)n_combos * nbins * 8;
  const size_t base_bytes = (size_t)((nbins + 1 + 3) & ~3) * 4 +
                            (size_t)ADF_WARPS * capacity * 16 +
                            (size_t)((ADF_WARPS * capacity + 15) & ~15);
  P.smem_hist = base_bytes + hist_bytes <= 200 * 1024 ? 1 : 0;
  const size_t smem = base_bytes + (P.smem_hist ? hist_bytes : 0);
  MDK_CHECK_ARG(smem <= 227 * 1024, "adf_hist: capacity %d with %d bins does not fit shared "
                "memory", capacity, nbins);
  MDK_CUDA(cudaFuncSetAttribute(adf_triplet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sm_count() * per_sm;
  const long long need = (total + ADF_WARPS - 1) / ADF_WARPS;
  if (grid > need) grid = need;
  adf_triplet_kernel<<<(unsigned)grid, ADF_WARPS * 32, smem, s>>>(P);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}
