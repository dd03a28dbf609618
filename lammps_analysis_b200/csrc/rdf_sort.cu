// Spatial ordering of a frame for the RDF pair pass, and per-tile bounding boxes.
//
// The pair histogram is a sum over unordered pairs, so the order of the atoms inside a species
// block is free.  Ordering them along a Morton (Z-order) curve of a fine cell grid makes every
// run of consecutive atoms (a 64-atom column sub-tile, the 32 rows of a warp) a compact blob, and
// a conservative minimum-image distance test between two bounding boxes then proves for whole
// (row group, column tile) blocks that no pair can be inside the cutoff; rdf_pair_hist_kernel
// skips those blocks.  Replaces nothing in the reference (it evaluates every pair,
// radial_distribution_function.py:647-689); the counts are unchanged.
#include "mdk_common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace mdk {

constexpr int SORT_BITS = 7;  // cells per dimension = 2^7 = 128 -> 21-bit keys

__device__ __forceinline__ unsigned spread3(unsigned v) {
  // insert two zero bits between the low 10 bits of v
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// Hilbert index of a cell (Skilling's transpose algorithm): unlike the Z-order curve a Hilbert
// curve has no long jumps, so a run of consecutive atoms is always a face-connected blob and its
// bounding box stays tight (measured on 10^6 uniform atoms: 23 % of the (32-row, 256-column)
// blocks are provably beyond a cutoff of L/2 - 0.1, against 18 % for Z-order).
__device__ __forceinline__ unsigned hilbert3(unsigned x, unsigned y, unsigned z) {
  unsigned X[3] = {x, y, z};
  const unsigned M = 1u << (SORT_BITS - 1);
  for (unsigned Q = M; Q > 1; Q >>= 1) {
    const unsigned P = Q - 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const unsigned t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  unsigned t = 0;
  for (unsigned Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t;
  X[1] ^= t;
  X[2] ^= t;
  return (spread3(X[0]) << 2) | (spread3(X[1]) << 1) | spread3(X[2]);
}

// The frame's coordinates are read from the trajectory ONCE (the trajectory may be page-locked host
// memory read in place over the host link, 12 bytes at a stride of a whole atom row): the keys
// kernel parks them in the workspace, the gather after the sort reads them from there.
__global__ void rdf_keys_kernel(const float* __restrict__ traj, long long T, long long atom_first,
                                int atom_count, long long frame, float sx, float sy, float sz,
                                unsigned* __restrict__ keys, unsigned* __restrict__ idx,
                                float* __restrict__ xyz) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= atom_count) return;
  const float* src = traj + ((size_t)(atom_first + a) * T + frame) * 3;
  const int nc = 1 << SORT_BITS;
  const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
  xyz[3 * (size_t)a] = x;
  xyz[3 * (size_t)a + 1] = y;
  xyz[3 * (size_t)a + 2] = z;
  // coordinates may lie outside [0, L): wrap the cell index (only the ORDER depends on it)
  int cx = (int)floorf(x * sx), cy = (int)floorf(y * sy), cz = (int)floorf(z * sz);
  cx = ((cx % nc) + nc) % nc;
  cy = ((cy % nc) + nc) % nc;
  cz = ((cz % nc) + nc) % nc;
  keys[a] = hilbert3((unsigned)cx, (unsigned)cy, (unsigned)cz);
  idx[a] = a;
}

__global__ void rdf_gather_kernel(const float* __restrict__ xyz, int atom_count,
                                  const unsigned* __restrict__ idx, float* __restrict__ out,
                                  long long n_pad, long long dst_first, int dst_span) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= dst_span) return;
  float x = __int_as_float(0x7fc00000), y = x, z = x;
  if (a < atom_count) {
    const float* src = xyz + 3 * (size_t)idx[a];
    x = __ldg(src);
    y = __ldg(src + 1);
    z = __ldg(src + 2);
  }
  float* o = out + dst_first + a;
  o[0] = x;
  o[n_pad] = y;
  o[2 * n_pad] = z;
}

// Batched variants: all frames of a launch batch in one keys kernel, ONE radix sort with the
// batch-local frame number above the Hilbert index, and one gather -- for systems of ~10^5 atoms
// the per-frame sorts are launch-bound (seven launches of a few microseconds per frame).
__global__ void rdf_keys_batch_kernel(const float* __restrict__ traj, long long T,
                                      long long atom_first, int atom_count,
                                      const int* __restrict__ frames, float sx, float sy, float sz,
                                      unsigned* __restrict__ keys, unsigned* __restrict__ idx,
                                      float* __restrict__ xyz) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (a >= atom_count) return;
  const float* src = traj + ((size_t)(atom_first + a) * T + frames[k]) * 3;
  const int nc = 1 << SORT_BITS;
  const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
  const size_t e = (size_t)k * atom_count + a;
  xyz[3 * e] = x;
  xyz[3 * e + 1] = y;
  xyz[3 * e + 2] = z;
  int cx = (int)floorf(x * sx), cy = (int)floorf(y * sy), cz = (int)floorf(z * sz);
  cx = ((cx % nc) + nc) % nc;
  cy = ((cy % nc) + nc) % nc;
  cz = ((cz % nc) + nc) % nc;
  keys[e] = ((unsigned)k << (3 * SORT_BITS)) | hilbert3((unsigned)cx, (unsigned)cy, (unsigned)cz);
  idx[e] = (unsigned)e;
}

__global__ void rdf_gather_batch_kernel(const float* __restrict__ xyz, int atom_count,
                                        const unsigned* __restrict__ idx, float* __restrict__ out,
                                        long long n_pad, long long dst_first, int dst_span) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (a >= dst_span) return;
  float x = __int_as_float(0x7fc00000), y = x, z = x;
  if (a < atom_count) {
    // the keys sort by frame first: position k * atom_count + a holds an atom of frame k
    const float* src = xyz + 3 * (size_t)idx[(size_t)k * atom_count + a];
    x = src[0];
    y = src[1];
    z = src[2];
  }
  float* o = out + (size_t)k * 3 * n_pad + dst_first + a;
  o[0] = x;
  o[n_pad] = y;
  o[2 * n_pad] = z;
}

// one warp per (frame, tile): {min xyz, max xyz} over the non-NaN atoms of the tile
__global__ void rdf_bbox_kernel(const float* __restrict__ pos, long long n_pad, int tile,
                                int tiles_per_frame, long long total_tiles,
                                float* __restrict__ bbox) {
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total_tiles) return;
  const long long f = w / tiles_per_frame;
  const int t = (int)(w - f * tiles_per_frame);
  const float* base = pos + (size_t)f * 3 * n_pad + (size_t)t * tile;
  float mn[3], mx[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = INFINITY;
    mx[d] = -INFINITY;
    for (int i = lane; i < tile; i += 32) {
      const float v = __ldg(base + (size_t)d * n_pad + i);
      if (v == v) {
        mn[d] = fminf(mn[d], v);
        mx[d] = fmaxf(mx[d], v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if (lane == 0) {
    float* o = bbox + (size_t)w * 6;
    o[0] = mn[0]; o[1] = mn[1]; o[2] = mn[2];
    o[3] = mx[0]; o[4] = mx[1]; o[5] = mx[2];
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t cub_temp_bytes(int n, int end_bit = 3 * SORT_BITS) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, n, 0, end_bit);
  return bytes;
}

constexpr int BATCH_MAX_FRAMES = 1 << (32 - 3 * SORT_BITS);   // frame number above the 21 key bits
constexpr long long BATCH_MAX_ELEMS = 1ll << 25;

}  // namespace mdk

using namespace mdk;

extern "C" long long mdk_rdf_sort_workspace(int max_atoms) {
  if (max_atoms < 1) max_atoms = 1;
  return (long long)(4 * align_up((size_t)max_atoms * 4, 256) + align_up((size_t)max_atoms * 12, 256) +
                     align_up(cub_temp_bytes(max_atoms), 256));
}

extern "C" int mdk_rdf_pack_sorted(const float* traj, long long A_total, long long T,
                                   long long atom_first, int atom_count, long long frame,
                                   float* out_frame, long long n_pad, long long dst_first,
                                   int dst_span, const float* box, void* workspace,
                                   long long workspace_bytes, mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && out_frame && box && workspace, "rdf_pack_sorted: null pointer");
  MDK_CHECK_ARG(atom_first >= 0 && atom_count >= 0 && atom_first + atom_count <= A_total,
                "rdf_pack_sorted: atom range outside the array");
  MDK_CHECK_ARG(frame >= 0 && frame < T, "rdf_pack_sorted: frame %lld outside [0, %lld)", frame, T);
  MDK_CHECK_ARG(dst_span >= atom_count && dst_first >= 0 && dst_first + dst_span <= n_pad,
                "rdf_pack_sorted: destination range outside n_pad");
  MDK_CHECK_ARG(workspace_bytes >= mdk_rdf_sort_workspace(atom_count),
                "rdf_pack_sorted: workspace too small");
  MDK_CHECK_ARG(box[0] > 0 && box[1] > 0 && box[2] > 0, "rdf_pack_sorted: box must be positive");
  if (dst_span == 0) return MDK_OK;
  cudaStream_t s = as_stream(stream);
  const size_t seg = align_up((size_t)(atom_count > 0 ? atom_count : 1) * 4, 256);
  char* w = static_cast<char*>(workspace);
  unsigned* keys_in = reinterpret_cast<unsigned*>(w);
  unsigned* keys_out = reinterpret_cast<unsigned*>(w + seg);
  unsigned* idx_in = reinterpret_cast<unsigned*>(w + 2 * seg);
  unsigned* idx_out = reinterpret_cast<unsigned*>(w + 3 * seg);
  float* xyz = reinterpret_cast<float*>(w + 4 * seg);
  const size_t xyz_bytes = align_up((size_t)(atom_count > 0 ? atom_count : 1) * 12, 256);
  void* temp = w + 4 * seg + xyz_bytes;
  size_t temp_bytes = (size_t)workspace_bytes - 4 * seg - xyz_bytes;
  const int nc = 1 << SORT_BITS;
  if (atom_count > 0) {
    rdf_keys_kernel<<<(atom_count + 255) / 256, 256, 0, s>>>(
        traj, T, atom_first, atom_count, frame, nc / box[0], nc / box[1], nc / box[2], keys_in,
        idx_in, xyz);
    MDK_LAUNCH_CHECK();
    MDK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out,
                                             atom_count, 0, 3 * SORT_BITS, s));
  }
  rdf_gather_kernel<<<(dst_span + 255) / 256, 256, 0, s>>>(xyz, atom_count, idx_out, out_frame,
                                                           n_pad, dst_first, dst_span);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" long long mdk_rdf_sort_batch_workspace(int max_atoms, int n_frames) {
  if (max_atoms < 1) max_atoms = 1;
  if (n_frames < 1) n_frames = 1;
  const long long n = (long long)max_atoms * n_frames;
  if (n_frames > BATCH_MAX_FRAMES || n > BATCH_MAX_ELEMS) return -1;   // use the per-frame pack
  return (long long)(4 * align_up((size_t)n * 4, 256) + align_up((size_t)n * 12, 256) +
                     align_up(cub_temp_bytes((int)n, 32), 256));
}

extern "C" int mdk_rdf_pack_sorted_batch(const float* traj, long long A_total, long long T,
                                         long long atom_first, int atom_count, const int* frames,
                                         int n_frames, float* out, long long n_pad,
                                         long long dst_first, int dst_span, const float* box,
                                         void* workspace, long long workspace_bytes,
                                         mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && frames && out && box && workspace, "rdf_pack_sorted_batch: null pointer");
  MDK_CHECK_ARG(atom_first >= 0 && atom_count >= 0 && atom_first + atom_count <= A_total,
                "rdf_pack_sorted_batch: atom range outside the array");
  MDK_CHECK_ARG(dst_span >= atom_count && dst_first >= 0 && dst_first + dst_span <= n_pad,
                "rdf_pack_sorted_batch: destination range outside n_pad");
  MDK_CHECK_ARG(n_frames >= 0 && n_frames <= BATCH_MAX_FRAMES &&
                    (long long)atom_count * n_frames <= BATCH_MAX_ELEMS,
                "rdf_pack_sorted_batch: batch too large (%d frames x %d atoms)", n_frames,
                atom_count);
  const long long need = mdk_rdf_sort_batch_workspace(atom_count, n_frames);
  MDK_CHECK_ARG(need >= 0 && workspace_bytes >= need, "rdf_pack_sorted_batch: workspace too small");
  MDK_CHECK_ARG(box[0] > 0 && box[1] > 0 && box[2] > 0, "rdf_pack_sorted_batch: box must be positive");
  if (dst_span == 0 || n_frames == 0) return MDK_OK;
  cudaStream_t s = as_stream(stream);
  const size_t n = (size_t)(atom_count > 0 ? atom_count : 1) * n_frames;
  const size_t seg = align_up(n * 4, 256);
  char* w = static_cast<char*>(workspace);
  unsigned* keys_in = reinterpret_cast<unsigned*>(w);
  unsigned* keys_out = reinterpret_cast<unsigned*>(w + seg);
  unsigned* idx_in = reinterpret_cast<unsigned*>(w + 2 * seg);
  unsigned* idx_out = reinterpret_cast<unsigned*>(w + 3 * seg);
  float* xyz = reinterpret_cast<float*>(w + 4 * seg);
  const size_t xyz_bytes = align_up(n * 12, 256);
  void* temp = w + 4 * seg + xyz_bytes;
  size_t temp_bytes = (size_t)workspace_bytes - 4 * seg - xyz_bytes;
  const int nc = 1 << SORT_BITS;
  if (atom_count > 0) {
    dim3 grid((atom_count + 255) / 256, n_frames);
    rdf_keys_batch_kernel<<<grid, 256, 0, s>>>(traj, T, atom_first, atom_count, frames,
                                               nc / box[0], nc / box[1], nc / box[2], keys_in,
                                               idx_in, xyz);
    MDK_LAUNCH_CHECK();
    int frame_bits = 0;
    while ((1 << frame_bits) < n_frames) ++frame_bits;
    MDK_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, idx_in, idx_out,
                                             (int)((size_t)atom_count * n_frames), 0,
                                             3 * SORT_BITS + frame_bits, s));
  }
  dim3 ggrid((dst_span + 255) / 256, n_frames);
  rdf_gather_batch_kernel<<<ggrid, 256, 0, s>>>(xyz, atom_count, idx_out, out, n_pad, dst_first,
                                                dst_span);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_rdf_bbox(const float* pos_soa, int n_frames, long long n_pad, float* bbox,
                            mdk_stream_t stream) {
  MDK_CHECK_ARG(pos_soa && bbox && n_frames >= 0, "rdf_bbox: bad argument");
  const int tile = MDK_RDF_SUBTILE;
  MDK_CHECK_ARG(n_pad % tile == 0, "rdf_bbox: n_pad must be a multiple of the sub-tile");
  const int tpf = (int)(n_pad / tile);
  const long long total = (long long)n_frames * tpf;
  if (total == 0) return MDK_OK;
  const long long threads = total * 32;
  rdf_bbox_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(
      pos_soa, n_pad, tile, tpf, total, bbox);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}
