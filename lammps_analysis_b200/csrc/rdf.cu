// rdf_pair_hist: minimum-image all-pairs distance histogram, one launch for every
// species pair and every frame of a batch.
//
// Replaces the TF op chain of the reference
//   utils/linalg.py:102-122   get_partial_triu_indices  (never materialised here)
//   utils/linalg.py:84-99     apply_minimum_image
//   radial_distribution_function.py:647-689  get_dij   (gather, sub, min-image, norm)
//   radial_distribution_function.py:616-645  bin_minibatch (species mask, cutoff, histogram)
// with a tiled pair pass:
//   * persistent CTAs pull work items (species pair, frame, row tile, column chunk) from an
//     atomic counter;
//   * each thread keeps R row atoms in registers (negated, duplicated into f32x2 pairs);
//   * column tiles (x[], y[], z[] of TJ atoms) are staged into shared memory by the TMA
//     engine (cp.async.bulk + mbarrier, double buffered);
//   * the geometry runs on packed FFMA2/FADD2/FMUL2 (two column atoms per instruction) and
//     reproduces the reference's fp32 rounding sequence exactly:
//         r = p_j - p_i ; r -= rint(r/L)*L ; d2 = (x*x + y*y) + z*z
//   * binning is exact: a fast fp32 guess g = rint(sqrt.approx(d2)/step) is corrected
//     against a table of fp32 thresholds on d2 (mdk_rdf_thresholds) held in shared memory;
//   * counts go to a CTA-private u32 histogram in shared memory and are flushed to the
//     global u64 histogram when the species pair changes / the CTA retires;
//   * on Hilbert-sorted frames with bounding boxes (mdk_rdf_pack_sorted, mdk_rdf_bbox) whole
//     blocks of pairs beyond the cutoff are skipped, and blocks whose boxes prove one common
//     periodic image replace the per-pair rint by a warp-uniform shift (AM 7, sub_tile_uni).
//
// Variants (template parameter AM, selected in mdk_rdf_hist; all produce the same integers):
//   2          table compare for every in-cutoff lane, unconditional ATOMS.POPC.INC with one
//              dump word for the lanes outside the cutoff (general coordinates, small tiles)
//   4          tables and histogram in global memory (bin counts beyond shared memory)
//   5          as 2 with the wrapped-coordinate minimum image min(|d|, L - |d|)
//   7          DEFAULT on sorted frames: uniform-image blocks + clamped gated compare (6 fraction
//              bits in the bin guess: one lane in 64 reads its threshold)
//   8          DEFAULT on unsorted wrapped frames: wrapped minimum image + clamped gated compare
//              (2 fraction bits)
// Same-species tiles on the diagonal run on the same fast paths: blocks above the diagonal are
// counted, blocks below it skipped, the blocks it crosses go through sub_tile_tri.
// (round 1 also carried a predicated-red, a per-lane-dump, a 7-fraction-bit and a quarter-bit
// gated variant -- AM 0, 1, 3, 6; all measured slower, removed in round 2)
#include "mdk_common.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace mdk {

constexpr int TJ = 256;  // column tile (atoms) == padding granule of species blocks
constexpr int SUB = MDK_RDF_SUBTILE;  // atoms per bounding box (culling granule)
constexpr int NSUB = TJ / SUB;
constexpr int MAX_PAIRS = MDK_MAX_SPECIES * (MDK_MAX_SPECIES + 1) / 2;
constexpr float RINT_MAGIC = 12582912.0f;         // 1.5 * 2^23: ulp == 1, even
constexpr unsigned RINT_MAGIC_BITS = 0x4B400000u; // __float_as_uint(RINT_MAGIC)

struct RdfParams {
  const float* pos;            // [F][3][n_pad]
  long long n_pad;
  int n_frames;
  int n_species;
  int n_pairs;
  int sp_lo[MDK_MAX_SPECIES];
  int sp_hi[MDK_MAX_SPECIES];
  // per species pair p = (a, b)
  int pair_a[MAX_PAIRS];
  int pair_b[MAX_PAIRS];
  int row_tiles[MAX_PAIRS];      // ceil(len_a / TI)
  int col_tiles[MAX_PAIRS];      // len_b / TJ
  int chunks_per_row[MAX_PAIRS]; // ceil(col_tiles / CJ)
  unsigned long long item_start[MAX_PAIRS + 1];
  unsigned long long total_items;
  int CJ;                        // column tiles per work item
  float box[3];
  float inv_box[3];
  float cut2;
  float inv_step;
  int nbins;
  const float* thr;              // [nbins + 1]
  unsigned long long* hist;      // [n_pairs][nbins]
  unsigned long long* counter;   // dynamic work counter
  unsigned int flush_tiles;      // flush after this many column tiles (u32 overflow guard)
  unsigned int one;              // == 1, kept opaque to the compiler (see bin_two)
  float onef;                    // == 1.0f, opaque as well (packed exact adds, see sub_tile)
  int dump_shared;               // 1: all out-of-cutoff lanes hit ONE dump word (tuning)
  // culling (optional): per (frame, 256-atom tile) bounding boxes {min xyz, max xyz}
  const float* bbox;             // [F][boxes_per_frame][6] (one per SUB atoms) or nullptr
  int boxes_per_frame;
  float cull2;                   // squared distance beyond which a block cannot hold a pair
  float cull_eps[3];             // absolute slack per dimension for the box arithmetic
  float clamp2;                  // ((nbins + 1/2) * step)^2: see bin_two_c
};

__device__ __forceinline__ void box_union(float (&acc)[6], const float* __restrict__ b) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    acc[d] = fminf(acc[d], __ldg(b + d));
    acc[d + 3] = fmaxf(acc[d + 3], __ldg(b + d + 3));
  }
}

// Conservative test: can any pair (p in box a, q in box b) have a minimum-image distance below
// sqrt(cull2)?  Along one dimension |p - q| lies in [lo, hi] = [dc - h, dc + h] (dc = distance
// of the centres, h = sum of the half widths), and the minimum image of a value v in [0, L) is
// min(v, L - v) >= min(lo, L - hi); when hi >= L the bound is <= 0 and nothing is culled, so the
// test is safe for any coordinates.  Returns true when the block can be skipped.
__device__ __forceinline__ bool boxes_far(const float (&a)[6], const float* __restrict__ b,
                                          const float (&L)[3], const float (&eps)[3],
                                          float cull2) {
  if (!(b[0] <= b[3])) return true;  // empty tile (padding only): nothing to count
  float g2 = 0.f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float dc = fabsf(0.5f * (a[d] + a[d + 3]) - 0.5f * (b[d] + b[d + 3]));
    const float h = 0.5f * (a[d + 3] - a[d]) + 0.5f * (b[d + 3] - b[d]) + eps[d];
    const float g = fmaxf(fminf(dc - h, L[d] - (dc + h)), 0.f);
    g2 = fmaf(g, g, g2);
  }
  return g2 > cull2;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact bin of an in-cutoff squared distance: k in {g-1, g}, g = rint(sqrt(d2)/step)
// (|sqrt.approx error| * nbins << 0.5), decided against the threshold table thr[g].
// Scalar C++ form, used on the slow paths (diagonal tiles, exact-division mode).
__device__ __forceinline__ void bin_one(float d2, bool valid, const float* __restrict__ s_thr,
                                        unsigned int* __restrict__ s_cnt, float inv_step) {
  if (valid) {
    const float d = sqrt_approx(d2);
    const float t = fmaf(d, inv_step, RINT_MAGIC);
    const unsigned g = __float_as_uint(t) - RINT_MAGIC_BITS;
    const float tg = s_thr[g];
    const unsigned k = g - (d2 >= tg ? 0u : 1u);
    atomicAdd(&s_cnt[k], 1u);
  }
}

// Same decision with the threshold table and the u64 histogram in GLOBAL memory: the fallback
// for bin counts whose tables do not fit in shared memory (AM == 4).
__device__ __forceinline__ void bin_one_global(float d2, bool valid, const float* __restrict__ thr,
                                               unsigned long long* __restrict__ hist,
                                               float inv_step) {
  if (valid) {
    const float d = sqrt_approx(d2);
    const float t = fmaf(d, inv_step, RINT_MAGIC);
    const unsigned g = __float_as_uint(t) - RINT_MAGIC_BITS;
    const float tg = __ldg(thr + g);
    const unsigned k = g - (d2 >= tg ? 0u : 1u);
    atomicAdd(hist + k, 1ull);
  }
}

// Same decision for two squared distances without divergent control flow:
//   p  = d2 < cut2                       in cutoff
//   g  = bits(fma(sqrt(d2), 1/step, 1.5*2^23)) - bits(1.5*2^23)
//   tg = thr[g]                           (shared, predicated on p)
//   k  = g - (d2 < tg) ;  p -> ++cnt[k]
// thr_c = smem address of thr[0] minus 4*bits(1.5*2^23); delta = &cnt[0] - &thr[0] (bytes).
// ptxas never predicates ATOMS (it wraps it in BSSY/BRA/BSYNC), so the increment is issued
// unconditionally with a literal 1 (ATOMS.POPC.INC): lanes outside the cutoff all hit one dump
// word, which POPC.INC merges into a single increment.
__device__ __forceinline__ void bin_two(float2 d2, float cut2, float2 inv_step2, uint32_t thr_c,
                                        uint32_t delta, uint32_t one, uint32_t dump) {
#define MDK_BIN_HEAD                                   \
  "{\n"                                                \
  ".reg .pred p0, q0, p1, q1;\n"                       \
  ".reg .f32 e0, e1, t0, t1;\n"                        \
  ".reg .b64 ee, tt;\n"                                \
  ".reg .u32 a0, a1, b0, b1;\n"                        \
  "setp.lt.f32 p0, %0, %2;\n"                          \
  "setp.lt.f32 p1, %1, %2;\n"                          \
  "sqrt.approx.ftz.f32 e0, %0;\n"                      \
  "sqrt.approx.ftz.f32 e1, %1;\n"                      \
  "mov.b64 ee, {e0, e1};\n"                            \
  "fma.rn.f32x2 tt, ee, %3, %4;\n"                     \
  "mov.b64 {t0, t1}, tt;\n"                            \
  "mov.b32 b0, t0;\n"                                  \
  "mov.b32 b1, t1;\n"                                  \
  "mad.lo.u32 a0, b0, 4, %5;\n"                        \
  "mad.lo.u32 a1, b1, 4, %5;\n"                        \
  "@p0 ld.shared.f32 e0, [a0];\n"                      \
  "@p1 ld.shared.f32 e1, [a1];\n"                      \
  "setp.lt.f32 q0, %0, e0;\n"                          \
  "setp.lt.f32 q1, %1, e1;\n"                          \
  "@q0 add.u32 a0, a0, -4;\n"                          \
  "@q1 add.u32 a1, a1, -4;\n"
#define MDK_BIN_ARGS                                                                        \
  "f"(d2.x), "f"(d2.y), "f"(cut2), "l"(*reinterpret_cast<unsigned long long*>(&inv_step2)), \
      "l"(0x4B4000004B400000ull), "r"(thr_c), "r"(delta), "r"(one), "r"(dump)
  asm volatile(MDK_BIN_HEAD
               "selp.u32 a0, a0, %8, p0;\n"
               "selp.u32 a1, a1, %8, p1;\n"
               "add.u32 a0, a0, %6;\n"
               "add.u32 a1, a1, %6;\n"
               "red.shared.add.u32 [a0], 1;\n"
               "red.shared.add.u32 [a1], 1;\n"
               "}\n" ::MDK_BIN_ARGS
               : "memory");
#undef MDK_BIN_HEAD
#undef MDK_BIN_ARGS
}

// Clamped, quarter-bit gated variant (AM == 7 / 8).  The guess carries two fraction bits,
//   tm = fma(sqrt.approx(d2), 4/step, 1.5 * 2^23)  ->  mantissa = round(4 t),  t = d / step,
// so (mantissa & ~3) is already the byte offset of bin floor(t) and (mantissa & 3) == 0 says
// that t lies within 1/8 of an integer: the exact bin differs from t by < 0.01, hence only those
// lanes (one in four) need the threshold-table compare.  There is no in-cutoff predicate and
// no dump-slot select: d2 is first clamped to clamp2 = ((nbins + 1/2) *
// step)^2 (min.f32 also maps the NaN of padding atoms to it), so every lane well outside the
// cutoff computes M = 4 * nbins + 2, needs no table compare and increments the unused word
// cnt[nbins] (one address for all such lanes: ATOMS.POPC.INC merges them).  Lanes whose guess
// rounds to exactly 4 * nbins are "ambiguous" and read thr[nbins] -- which the kernel sets to
// cut2 in its shared copy of the table: inside the cutoff they pass d2 < cut2 and land in bin
// nbins - 1, as they must, outside they stay on cnt[nbins].  Per pair: FMNMX, MUFU, LOP3 x2,
// FSETP, predicated IADD on the ALU pipe -- two fewer than bin_two -- and three in four of the
// random threshold gathers are predicated off.
// The shared-memory address of the threshold table is folded into the magic constant
// (magic_thr = 1.5 * 2^23 + &thr[0], an integer below 2^24 and a multiple of 4), so the masked
// mantissa IS the address of thr[floor-or-carry(t)]; cnt_delta = &cnt[0] - &thr[0] rides in the
// ATOMS address as a uniform register.
// BIN_FRAC fraction bits in the guess, M = round(2^BIN_FRAC t): a lane needs the threshold
// compare only when the BIN_FRAC low bits of M are all zero, i.e. t lies within 2^-(BIN_FRAC+1)
// of an integer.  The guess t is off by at most 0.003 bins at 13.5k bins (sqrt.approx 2^-23
// relative, the fp32 scale 2^-24, the table's own rounding), so up to BIN_FRAC = 6 (zone half
// width 1/128 = 0.0078) the decision stays exact; beyond that the margin is gone.
//   2: M = round(4 t) is the byte offset itself, one lane in four gathers its threshold, and
//      every predicated LDS carries ~8 random addresses (1.3 wavefronts)
//   6: one lane in 64 -- six LDS in ten find no active lane at all and cost no shared-memory
//      wavefront (0.4 on average) -- for one extra shift per pair.  The shared-memory pipe is
//      the limiter of this kernel (90 % busy in the round-1 capture), the ALU pipe is not.
// The magic constant carries the table address scaled by 2^(BIN_FRAC - 2) (bin_magic()).
#ifndef MDK_RDF_BIN_FRAC
#define MDK_RDF_BIN_FRAC 6
#endif
// Measured on B200 (10^6 atoms sorted / 10^5 sorted / 10^5 unsorted): 6 bits +2.4 % / +0.7 % /
// -4.5 % against 2 bits, so the sorted kernel (AM 7) takes 6 and the unsorted one (AM 8), which
// is ALU bound, keeps 2.
__host__ __device__ constexpr int bin_frac_of(int am) { return am == 7 ? MDK_RDF_BIN_FRAC : 2; }
template <int BIN_FRAC>
__device__ __forceinline__ float bin_magic(uint32_t thr_s) {
  return RINT_MAGIC + static_cast<float>(thr_s << (BIN_FRAC - 2));
}
template <int BIN_FRAC>
__device__ __forceinline__ float bin_scale(float inv_step) {
  return static_cast<float>(1 << BIN_FRAC) * inv_step;
}
template <int BIN_FRAC>
__device__ __forceinline__ void bin_two_c(float2 d2, float clamp2, float2 inv_step4,
                                          float2 magic_thr, uint32_t cnt_delta) {
  static_assert(BIN_FRAC >= 2 && BIN_FRAC <= 6, "see the error budget above");
  if (BIN_FRAC == 2) {
    asm volatile(
        "{\n"
        ".reg .pred m0, m1, q0, q1;\n"
        ".reg .f32 c0, c1, e0, e1, t0, t1;\n"
        ".reg .b64 ee, tt;\n"
        ".reg .u32 a0, a1, b0, b1, f0, f1;\n"
        "min.f32 c0, %0, %2;\n"
        "min.f32 c1, %1, %2;\n"
        "sqrt.approx.ftz.f32 e0, c0;\n"
        "sqrt.approx.ftz.f32 e1, c1;\n"
        "mov.b64 ee, {e0, e1};\n"
        "fma.rn.f32x2 tt, ee, %3, %4;\n"
        "mov.b64 {t0, t1}, tt;\n"
        "mov.b32 b0, t0;\n"
        "mov.b32 b1, t1;\n"
        "and.b32 a0, b0, 0x003ffffc;\n"
        "and.b32 a1, b1, 0x003ffffc;\n"
        "and.b32 f0, b0, 3;\n"
        "and.b32 f1, b1, 3;\n"
        "setp.eq.u32 m0, f0, 0;\n"
        "setp.eq.u32 m1, f1, 0;\n"
        "@m0 ld.shared.f32 e0, [a0];\n"  // into the (dead) sqrt register: a fresh destination
        "@m1 ld.shared.f32 e1, [a1];\n"  // of a predicated load would be loop-carried
        "setp.lt.and.f32 q0, c0, e0, m0;\n"
        "setp.lt.and.f32 q1, c1, e1, m1;\n"
        "@q0 add.u32 a0, a0, -4;\n"
        "@q1 add.u32 a1, a1, -4;\n"
        "add.u32 a0, a0, %5;\n"
        "add.u32 a1, a1, %5;\n"
        "red.shared.add.u32 [a0], 1;\n"
        "red.shared.add.u32 [a1], 1;\n"
        "}\n" ::"f"(d2.x),
        "f"(d2.y), "f"(clamp2), "l"(*reinterpret_cast<unsigned long long*>(&inv_step4)),
        "l"(*reinterpret_cast<unsigned long long*>(&magic_thr)), "r"(cnt_delta)
        : "memory");
  } else {
    // mantissa = (table address << (BIN_FRAC - 2)) + M: shift the fraction bits out, clear the
    // two address bits they leave behind (and the exponent bits of the magic constant)
    asm volatile(
        "{\n"
        ".reg .pred m0, m1, q0, q1;\n"
        ".reg .f32 c0, c1, e0, e1, t0, t1;\n"
        ".reg .b64 ee, tt;\n"
        ".reg .u32 a0, a1, b0, b1, f0, f1;\n"
        "min.f32 c0, %0, %2;\n"
        "min.f32 c1, %1, %2;\n"
        "sqrt.approx.ftz.f32 e0, c0;\n"
        "sqrt.approx.ftz.f32 e1, c1;\n"
        "mov.b64 ee, {e0, e1};\n"
        "fma.rn.f32x2 tt, ee, %3, %4;\n"
        "mov.b64 {t0, t1}, tt;\n"
        "mov.b32 b0, t0;\n"
        "mov.b32 b1, t1;\n"
        "shr.u32 a0, b0, %6;\n"
        "shr.u32 a1, b1, %6;\n"
        "and.b32 a0, a0, %7;\n"
        "and.b32 a1, a1, %7;\n"
        "and.b32 f0, b0, %8;\n"
        "and.b32 f1, b1, %8;\n"
        "setp.eq.u32 m0, f0, 0;\n"
        "setp.eq.u32 m1, f1, 0;\n"
        "@m0 ld.shared.f32 e0, [a0];\n"
        "@m1 ld.shared.f32 e1, [a1];\n"
        "setp.lt.and.f32 q0, c0, e0, m0;\n"
        "setp.lt.and.f32 q1, c1, e1, m1;\n"
        "@q0 add.u32 a0, a0, -4;\n"
        "@q1 add.u32 a1, a1, -4;\n"
        "add.u32 a0, a0, %5;\n"
        "add.u32 a1, a1, %5;\n"
        "red.shared.add.u32 [a0], 1;\n"
        "red.shared.add.u32 [a1], 1;\n"
        "}\n" ::"f"(d2.x),
        "f"(d2.y), "f"(clamp2), "l"(*reinterpret_cast<unsigned long long*>(&inv_step4)),
        "l"(*reinterpret_cast<unsigned long long*>(&magic_thr)), "r"(cnt_delta),
        "n"(BIN_FRAC - 2), "n"((0x003fffffu >> (BIN_FRAC - 2)) & ~3u), "n"((1 << BIN_FRAC) - 1)
        : "memory");
  }
}

// Flush the CTA-private histogram to the global one and clear it.
template <int NT>
__device__ __forceinline__ void flush_hist(unsigned int* s_cnt, int nbins,
                                           unsigned long long* __restrict__ ghist) {
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += NT) {
    const unsigned v = s_cnt[b];
    if (v) {
      atomicAdd(&ghist[b], static_cast<unsigned long long>(v));
      s_cnt[b] = 0u;
    }
  }
  __syncthreads();
}

// Loop-invariant operands of the pair arithmetic.
struct GeoConst {
  float invLx, invLy, invLz, nLx, nLy, nLz, inv_step;
  float cut2, onef, Lx, Ly, Lz, clamp2;
  uint32_t thr_c, cnt_delta, one, dump, thr_s, cnt_s, dump_off;
};
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

// One SUB-atom column sub-tile against the row groups selected by the compile-time mask M (bit r
// = row group r of this warp is within reach).  The arithmetic is the reference's rounding
// sequence; see the kernel header.
template <bool MASKED, int R, int AM, int UNROLL>
__device__ __forceinline__ void sub_tile(unsigned m, const float* __restrict__ sx,
                                         const float* __restrict__ sy,
                                         const float* __restrict__ sz, int jj0,
                                         const float2 (&nxi)[R], const float2 (&nyi)[R],
                                         const float2 (&nzi)[R], const GeoConst& c) {
  constexpr bool WRAP = (AM == 5 || AM == 8);  // wrapped-coordinate minimum image
  constexpr int BF = bin_frac_of(AM);
  const float2 magic_thr = dup2(bin_magic<BF>(c.thr_s));  // see bin_two_c
  const float2 magic2 = dup2(RINT_MAGIC), nmagic2 = dup2(-RINT_MAGIC);
  const float2 inv_step2 = dup2(c.inv_step);
  const float2 one2 = dup2(c.onef);
#pragma unroll UNROLL
  for (int jj = jj0; jj < jj0 + SUB; jj += 2) {
    const float2 xj = *reinterpret_cast<const float2*>(sx + jj);
    const float2 yj = *reinterpret_cast<const float2*>(sy + jj);
    const float2 zj = *reinterpret_cast<const float2*>(sz + jj);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!MASKED || ((m >> r) & 1u)) {  // warp-uniform
        const float2 dx = __fadd2_rn(xj, nxi[r]);
        const float2 dy = __fadd2_rn(yj, nyi[r]);
        const float2 dz = __fadd2_rn(zj, nzi[r]);
        float2 rx, ry, rz;
        if (WRAP) {
          // all coordinates lie within one box length (checked on the host), so |d| < L and
          // rint(d / L) is -1, 0 or 1: the minimum image is min(|d|, L - |d|) -- the same fp32
          // value as d - rint(d / L) * L up to sign (the subtraction rounds symmetrically), and
          // where the two candidates tie (|d| = L/2) both exceed the cutoff.  One FADD and one
          // FMNMX (ALU pipe) replace FFMA2 + FADD2 + FFMA2 per component.
          const float2 bx = __fadd2_rn(dup2(c.Lx), make_float2(-fabsf(dx.x), -fabsf(dx.y)));
          const float2 by = __fadd2_rn(dup2(c.Ly), make_float2(-fabsf(dy.x), -fabsf(dy.y)));
          const float2 bz = __fadd2_rn(dup2(c.Lz), make_float2(-fabsf(dz.x), -fabsf(dz.y)));
          rx = make_float2(fminf(fabsf(dx.x), bx.x), fminf(fabsf(dx.y), bx.y));
          ry = make_float2(fminf(fabsf(dy.x), by.x), fminf(fabsf(dy.y), by.y));
          rz = make_float2(fminf(fabsf(dz.x), bz.x), fminf(fabsf(dz.y), bz.y));
        } else {
          const float2 tx = __ffma2_rn(dx, dup2(c.invLx), magic2);
          const float2 ty = __ffma2_rn(dy, dup2(c.invLy), magic2);
          const float2 tz = __ffma2_rn(dz, dup2(c.invLz), magic2);
          const float2 nx = __fadd2_rn(tx, nmagic2);
          const float2 ny = __fadd2_rn(ty, nmagic2);
          const float2 nz = __fadd2_rn(tz, nmagic2);
          rx = __ffma2_rn(nx, dup2(c.nLx), dx);
          ry = __ffma2_rn(ny, dup2(c.nLy), dy);
          rz = __ffma2_rn(nz, dup2(c.nLz), dz);
        }
        // (x*x + y*y) + z*z with every product and sum rounded separately, as the reference
        // does.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with explicit
        // .rn), which would fuse a product into a sum; the packed adds are therefore written as
        // fma(a, 1.0f, b) with a run-time 1.0f the compiler cannot see through: a * 1 is exact,
        // so the result is the correctly rounded a + b, on two pairs per instruction.
        const float2 xx = __fmul2_rn(rx, rx);
        const float2 yy = __fmul2_rn(ry, ry);
        const float2 zz = __fmul2_rn(rz, rz);
        const float2 d2 = __ffma2_rn(zz, one2, __ffma2_rn(xx, one2, yy));
        if (AM == 7 || AM == 8)
          bin_two_c<BF>(d2, c.clamp2, dup2(bin_scale<BF>(c.inv_step)), magic_thr, c.cnt_delta);
        else
          bin_two(d2, c.cut2, inv_step2, c.thr_c, c.cnt_delta, c.one, c.dump);
      }
    }
  }
}

// Uniform-image variant of sub_tile (AM == 7).  For a block of pairs (32 rows of a warp x one
// SUB-atom column sub-tile) whose bounding boxes show that n = rint((x_j - x_i) / L) is the same
// integer for every pair of the block (per dimension), the reference's
//     r = d - rint(d / L) * L          (utils/linalg.py:84-99; here fma(n, -L, d))
// is d + sh with the warp-uniform sh = -n * L (exact for |n| <= 2): one packed FADD2 per
// component and pair couple instead of FFMA2 + FADD2 + FFMA2 (or FADD2 + 2 FMNMX for wrapped
// coordinates).  The block classification is done once per column tile by 16 lanes in parallel
// (rdf_pair_hist_kernel); blocks that straddle a half-box boundary take the general path.
template <int R, bool SHIFT>
__device__ __forceinline__ void sub_tile_uni(const float* __restrict__ sx,
                                             const float* __restrict__ sy,
                                             const float* __restrict__ sz, int jj0,
                                             const float2 (&nxi)[R], const float2 (&nyi)[R],
                                             const float2 (&nzi)[R], const float (&shx)[R],
                                             const float (&shy)[R], const float (&shz)[R],
                                             const GeoConst& c) {
  constexpr int BF = bin_frac_of(7);
  const float2 inv_step4 = dup2(bin_scale<BF>(c.inv_step));
  const float2 magic_thr = dup2(bin_magic<BF>(c.thr_s));  // see bin_two_c
  const float2 one2 = dup2(c.onef);
#pragma unroll 1
  for (int jj = jj0; jj < jj0 + SUB; jj += 4) {
    // four columns per trip: one 16-byte broadcast load per coordinate array
    const float4 xj4 = *reinterpret_cast<const float4*>(sx + jj);
    const float4 yj4 = *reinterpret_cast<const float4*>(sy + jj);
    const float4 zj4 = *reinterpret_cast<const float4*>(sz + jj);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 xj = h ? make_float2(xj4.z, xj4.w) : make_float2(xj4.x, xj4.y);
      const float2 yj = h ? make_float2(yj4.z, yj4.w) : make_float2(yj4.x, yj4.y);
      const float2 zj = h ? make_float2(zj4.z, zj4.w) : make_float2(zj4.x, zj4.y);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // SHIFT == false: no pair of the block crosses a periodic boundary (all shifts are 0)
        float2 rx = __fadd2_rn(xj, nxi[r]);
        float2 ry = __fadd2_rn(yj, nyi[r]);
        float2 rz = __fadd2_rn(zj, nzi[r]);
        if (SHIFT) {
          rx = __fadd2_rn(rx, dup2(shx[r]));
          ry = __fadd2_rn(ry, dup2(shy[r]));
          rz = __fadd2_rn(rz, dup2(shz[r]));
        }
        const float2 xx = __fmul2_rn(rx, rx);
        const float2 yy = __fmul2_rn(ry, ry);
        const float2 zz = __fmul2_rn(rz, rz);
        const float2 d2 = __ffma2_rn(zz, one2, __ffma2_rn(xx, one2, yy));
        bin_two_c<BF>(d2, c.clamp2, inv_step4, magic_thr, c.cnt_delta);
      }
    }
  }
}

// One row group of the warp against a SUB-atom column sub-tile (partially live blocks): eight
// columns per trip keep four independent pair couples in flight; no per-row-group branches in
// the column loop.
template <bool SHIFT>
__device__ __forceinline__ void sub_tile_uni_row(const float* __restrict__ sx,
                                                 const float* __restrict__ sy,
                                                 const float* __restrict__ sz, int jj0,
                                                 float2 nx, float2 ny, float2 nz, float shx,
                                                 float shy, float shz, const GeoConst& c) {
  constexpr int BF = bin_frac_of(7);
  const float2 inv_step4 = dup2(bin_scale<BF>(c.inv_step));
  const float2 magic_thr = dup2(bin_magic<BF>(c.thr_s));
  const float2 one2 = dup2(c.onef);
#pragma unroll 1
  for (int jj = jj0; jj < jj0 + SUB; jj += 8) {
    float4 xj4[2], yj4[2], zj4[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      xj4[v] = *reinterpret_cast<const float4*>(sx + jj + 4 * v);
      yj4[v] = *reinterpret_cast<const float4*>(sy + jj + 4 * v);
      zj4[v] = *reinterpret_cast<const float4*>(sz + jj + 4 * v);
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 xj = h ? make_float2(xj4[v].z, xj4[v].w) : make_float2(xj4[v].x, xj4[v].y);
        const float2 yj = h ? make_float2(yj4[v].z, yj4[v].w) : make_float2(yj4[v].x, yj4[v].y);
        const float2 zj = h ? make_float2(zj4[v].z, zj4[v].w) : make_float2(zj4[v].x, zj4[v].y);
        float2 rx = __fadd2_rn(xj, nx);
        float2 ry = __fadd2_rn(yj, ny);
        float2 rz = __fadd2_rn(zj, nz);
        if (SHIFT) {
          rx = __fadd2_rn(rx, dup2(shx));
          ry = __fadd2_rn(ry, dup2(shy));
          rz = __fadd2_rn(rz, dup2(shz));
        }
        const float2 xx = __fmul2_rn(rx, rx);
        const float2 yy = __fmul2_rn(ry, ry);
        const float2 zz = __fmul2_rn(rz, rz);
        const float2 d2 = __ffma2_rn(zz, one2, __ffma2_rn(xx, one2, yy));
        bin_two_c<BF>(d2, c.clamp2, inv_step4, magic_thr, c.cnt_delta);
      }
    }
  }
}

// General minimum image (r - rint(r / L) * L by magic-number rounding) for one row group of the
// warp: the blocks of AM 7 that straddle a half-box boundary.
__device__ __forceinline__ void sub_tile_gen_row(const float* __restrict__ sx,
                                                 const float* __restrict__ sy,
                                                 const float* __restrict__ sz, int jj0,
                                                 float2 nx, float2 ny, float2 nz,
                                                 const GeoConst& c) {
  constexpr int BF = bin_frac_of(7);
  const float2 inv_step4 = dup2(bin_scale<BF>(c.inv_step));
  const float2 magic_thr = dup2(bin_magic<BF>(c.thr_s));
  const float2 magic2 = dup2(RINT_MAGIC), nmagic2 = dup2(-RINT_MAGIC);
  const float2 one2 = dup2(c.onef);
#pragma unroll 1
  for (int jj = jj0; jj < jj0 + SUB; jj += 8) {
    float4 xj4[2], yj4[2], zj4[2];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      xj4[v] = *reinterpret_cast<const float4*>(sx + jj + 4 * v);
      yj4[v] = *reinterpret_cast<const float4*>(sy + jj + 4 * v);
      zj4[v] = *reinterpret_cast<const float4*>(sz + jj + 4 * v);
    }
#pragma unroll
    for (int v = 0; v < 2; ++v) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 xj = h ? make_float2(xj4[v].z, xj4[v].w) : make_float2(xj4[v].x, xj4[v].y);
        const float2 yj = h ? make_float2(yj4[v].z, yj4[v].w) : make_float2(yj4[v].x, yj4[v].y);
        const float2 zj = h ? make_float2(zj4[v].z, zj4[v].w) : make_float2(zj4[v].x, zj4[v].y);
        const float2 dx = __fadd2_rn(xj, nx);
        const float2 dy = __fadd2_rn(yj, ny);
        const float2 dz = __fadd2_rn(zj, nz);
        const float2 qx = __fadd2_rn(__ffma2_rn(dx, dup2(c.invLx), magic2), nmagic2);
        const float2 qy = __fadd2_rn(__ffma2_rn(dy, dup2(c.invLy), magic2), nmagic2);
        const float2 qz = __fadd2_rn(__ffma2_rn(dz, dup2(c.invLz), magic2), nmagic2);
        const float2 rx = __ffma2_rn(qx, dup2(c.nLx), dx);
        const float2 ry = __ffma2_rn(qy, dup2(c.nLy), dy);
        const float2 rz = __ffma2_rn(qz, dup2(c.nLz), dz);
        const float2 xx = __fmul2_rn(rx, rx);
        const float2 yy = __fmul2_rn(ry, ry);
        const float2 zz = __fmul2_rn(rz, rz);
        const float2 d2 = __ffma2_rn(zz, one2, __ffma2_rn(xx, one2, yy));
        bin_two_c<BF>(d2, c.clamp2, inv_step4, magic_thr, c.cnt_delta);
      }
    }
  }
}

// One row group of the warp against a SUB-atom column sub-tile that straddles the diagonal of a
// same-species tile: only pairs j > i count.  Scalar arithmetic, the reference's rounding
// sequence (general minimum image); a warp meets at most two such blocks per diagonal tile.
__device__ __forceinline__ void sub_tile_tri(const float* __restrict__ sx,
                                             const float* __restrict__ sy,
                                             const float* __restrict__ sz, int jj0, int j_first,
                                             int i, float nx, float ny, float nz,
                                             const GeoConst& c, const float* __restrict__ s_thr,
                                             unsigned int* __restrict__ s_cnt) {
#pragma unroll 2
  for (int jj = jj0; jj < jj0 + SUB; ++jj) {
    float rx = __fadd_rn(sx[jj], nx);
    float ry = __fadd_rn(sy[jj], ny);
    float rz = __fadd_rn(sz[jj], nz);
    const float qx = __fadd_rn(fmaf(rx, c.invLx, RINT_MAGIC), -RINT_MAGIC);
    const float qy = __fadd_rn(fmaf(ry, c.invLy, RINT_MAGIC), -RINT_MAGIC);
    const float qz = __fadd_rn(fmaf(rz, c.invLz, RINT_MAGIC), -RINT_MAGIC);
    rx = fmaf(qx, c.nLx, rx);
    ry = fmaf(qy, c.nLy, ry);
    rz = fmaf(qz, c.nLz, rz);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
    bin_one(d2, (d2 < c.cut2) && (j_first + jj > i), s_thr, s_cnt, c.inv_step);
  }
}

// register budget: 768 resident threads per SM without culling (85 registers), 512 with it
// (AM 7 needs 94; the 112 KB of tables per CTA at 13.5k bins allow two CTAs per SM anyway)
template <int NT, int R, bool EXACT, int AM, bool CULL>
__global__ void __launch_bounds__(NT, NT == 384 ? 2 : (CULL ? 512 : 768) / NT) rdf_pair_hist_kernel(const __grid_constant__ RdfParams P) {
  constexpr int TI = NT * R;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [stage0 xyz | stage1 xyz | mbar x4 | item, release counters | row-tile box |
  // thr (nbins+1) | cnt (nbins, padded to 32) | 32 dump slots]
  float* s_tile = reinterpret_cast<float*>(smem_raw);                 // 2 * 3 * TJ floats
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tile + 2 * 3 * TJ); // full[2], empty[2]
  uint64_t* s_ebar = s_bar + 2;
  unsigned long long* s_item = reinterpret_cast<unsigned long long*>(s_bar + 4);
  unsigned int* s_rel = reinterpret_cast<unsigned int*>(s_item + 1);  // warps done, per stage
  float* s_rbox = reinterpret_cast<float*>(s_item + 2);               // row-tile box (8 floats)
  float* s_thr = reinterpret_cast<float*>(s_item + 6);
  const int thr_len = (P.nbins + 1 + 3) & ~3;
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_thr + thr_len);

  const int tid = threadIdx.x;
  constexpr bool GLOBAL_HIST = (AM == 4);  // tables too large for shared memory
  if (!GLOBAL_HIST) {
    // AM 7 (bin_two_c) needs thr[nbins] == cut2; for the other variants any value > cut2 works
    for (int b = tid; b <= P.nbins; b += NT)
      s_thr[b] = ((AM == 7 || AM == 8) && b == P.nbins) ? P.cut2 : __ldg(P.thr + b);
    for (int b = tid; b < P.nbins; b += NT) s_cnt[b] = 0u;
  }
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_init(&s_ebar[0], NT / 32);  // one arrival per warp
    mbar_init(&s_ebar[1], NT / 32);
    s_rel[0] = 0u;
    s_rel[1] = 0u;
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t phase = 0u;   // bit s: parity to wait for on the full barrier of stage s
  // "stage is free" barriers: one phase per tile and stage (NT / 32 arrivals)
  uint32_t ephase = 0u;  // bit s: same for the empty barrier
  int cur_pair = -1;
  unsigned int tiles_since_flush = 0;

  const float inv_step = P.inv_step;
  const float2 inv_step2 = make_float2(inv_step, inv_step);
  const float cut2 = P.cut2;
  const uint32_t thr_c = smem_u32(s_thr) - 4u * RINT_MAGIC_BITS;
  const uint32_t cnt_delta = smem_u32(s_cnt) - smem_u32(s_thr);
  const uint32_t one = P.one;
  // dump slot of this lane, expressed relative to the thr table (bin_two adds cnt_delta)
  const uint32_t dump = smem_u32(s_thr) + 4u * ((P.nbins + 31) & ~31) + (P.dump_shared ? 0u : 4u * (tid & 31));
  const uint32_t thr_s = smem_u32(s_thr), cnt_s = smem_u32(s_cnt);
  const uint32_t dump_off = 4u * ((P.nbins + 31) & ~31) + (P.dump_shared ? 0u : 4u * (tid & 31));  // relative to cnt
  const float2 magic2 = make_float2(RINT_MAGIC, RINT_MAGIC);
  const float2 nmagic2 = make_float2(-RINT_MAGIC, -RINT_MAGIC);
  const float2 invLx = make_float2(P.inv_box[0], P.inv_box[0]);
  const float2 invLy = make_float2(P.inv_box[1], P.inv_box[1]);
  const float2 invLz = make_float2(P.inv_box[2], P.inv_box[2]);
  const float2 nLx = make_float2(-P.box[0], -P.box[0]);
  const float2 nLy = make_float2(-P.box[1], -P.box[1]);
  const float2 nLz = make_float2(-P.box[2], -P.box[2]);
  const GeoConst geo = {P.inv_box[0], P.inv_box[1], P.inv_box[2], -P.box[0], -P.box[1],
                        -P.box[2], inv_step, cut2, P.onef, P.box[0], P.box[1], P.box[2], P.clamp2, thr_c,
                        cnt_delta, one, dump, thr_s, cnt_s, dump_off};

  for (;;) {
    if (tid == 0) s_item[0] = atomicAdd(P.counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item[0];
    __syncthreads();
    if (item >= P.total_items) break;

    // ---- decode item -> (pair, frame, row tile, column chunk) -------------------------
    int p = 0;
    while (p + 1 < P.n_pairs && item >= P.item_start[p + 1]) ++p;
    const unsigned long long local = item - P.item_start[p];
    const int cpr = P.chunks_per_row[p];
    const unsigned long long ipf = static_cast<unsigned long long>(P.row_tiles[p]) * cpr;
    const int f = static_cast<int>(local / ipf);
    const int rem = static_cast<int>(local % ipf);
    const int I = rem / cpr;
    const int c = rem % cpr;
    const int a = P.pair_a[p], b = P.pair_b[p];
    const bool same = (a == b);
    int j_tile0 = c * P.CJ;
    int j_tile1 = min(j_tile0 + P.CJ, P.col_tiles[p]);
    if (same) j_tile0 = max(j_tile0, I * (TI / TJ));  // only columns j > i
    if (j_tile0 >= j_tile1) continue;                 // empty item (below the diagonal)

    if (p != cur_pair) {
      if (cur_pair >= 0 && !GLOBAL_HIST)
        flush_hist<NT>(s_cnt, P.nbins, P.hist + (size_t)cur_pair * P.nbins);
      cur_pair = p;
      tiles_since_flush = 0;
    }

    const float* __restrict__ fx = P.pos + (size_t)f * 3 * P.n_pad;

    // ---- row atoms -> registers (negated, duplicated) -----------------------------------
    const int row_base = P.sp_lo[a] + I * TI;
    float2 nxi[R], nyi[R], nzi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      // the R row groups of a warp come from R different 256-row regions of the tile: their
      // culling masks differ, which balances the live work across the warps of the CTA
      const int i = row_base + r * NT + tid;
      float x = __int_as_float(0x7fc00000), y = x, z = x;  // NaN rows never pass d2 < cut2
      if (i < P.sp_hi[a]) {
        x = __ldg(fx + i);
        y = __ldg(fx + P.n_pad + i);
        z = __ldg(fx + 2 * P.n_pad + i);
      }
      nxi[r] = make_float2(-x, -x);
      nyi[r] = make_float2(-y, -y);
      nzi[r] = make_float2(-z, -z);
    }

    // ---- bounding boxes of this row tile (CTA) and of each warp's 32-row groups -----------
    const float* __restrict__ fbox = P.bbox + (size_t)f * P.boxes_per_frame * 6;
    float rbox[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float wbox[R][6];
    float mybox[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // AM 7: box of row group (lane & 3)
    if constexpr (CULL) {
      const int r_last = min(row_base + TI, P.sp_hi[a]);  // exclusive row bound
      for (int t = row_base / SUB; t * SUB < r_last; ++t) box_union(rbox, fbox + (size_t)t * 6);
      // the row-tile box is the same for every thread: park it in shared memory (all threads
      // store identical values; every warp reads what it, or another warp, wrote) so that it
      // does not occupy six registers across the pair loops
#pragma unroll
      for (int d = 0; d < 6; ++d) s_rbox[d] = rbox[d];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float c[3] = {-nxi[r].x, -nyi[r].x, -nzi[r].x};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          float mn = c[d] == c[d] ? c[d] : INFINITY;
          float mx = c[d] == c[d] ? c[d] : -INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          if constexpr (AM == 7) {
            if ((tid & 3) == r) {
              mybox[d] = mn;
              mybox[d + 3] = mx;
            }
          } else {
            wbox[r][d] = mn;
            wbox[r][d + 3] = mx;
          }
        }
      }
    }
    const int col_box0 = P.sp_lo[b] / SUB;  // index of the first column box in the frame table
    auto tile_far = [&](int jt) {
      float cb6[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int q = 0; q < NSUB; ++q) box_union(cb6, fbox + (size_t)(col_box0 + jt * NSUB + q) * 6);
      float rb[6];
#pragma unroll
      for (int d = 0; d < 6; ++d) rb[d] = s_rbox[d];
      return boxes_far(rb, cb6, P.box, P.cull_eps, P.cull2);
    };
    auto next_live = [&](int jt) {
      if constexpr (CULL)
        while (jt < j_tile1 && tile_far(jt)) ++jt;
      return jt;
    };

    // ---- column tiles: TMA double buffer ------------------------------------------------
    const int col_base = P.sp_lo[b];
    auto issue = [&](int jt, int stage) {
      float* dst = s_tile + stage * 3 * TJ;
      const size_t off = (size_t)col_base + (size_t)jt * TJ;
      mbar_expect_tx(&s_bar[stage], 3u * TJ * sizeof(float));
      bulk_g2s(dst, fx + off, TJ * sizeof(float), &s_bar[stage]);
      bulk_g2s(dst + TJ, fx + P.n_pad + off, TJ * sizeof(float), &s_bar[stage]);
      bulk_g2s(dst + 2 * TJ, fx + 2 * P.n_pad + off, TJ * sizeof(float), &s_bar[stage]);
    };
    int jt = next_live(j_tile0);
    if (jt >= j_tile1) continue;  // every column tile of this item is out of range
    int jn = next_live(jt + 1);
    // Column tiles flow through a two-stage ring.  Warps are NOT block-synchronised per tile:
    // a warp signals "done with this stage" on the stage's empty barrier and moves on to the
    // next tile as soon as its data has landed.  The LAST warp to release a stage (elected by a
    // shared-memory counter) refills it, so no warp ever blocks on the others: with block
    // culling the work per (warp, tile) varies, a per-tile __syncthreads made every warp wait
    // for the slowest one, and a fixed producer warp stalled itself on the empty barrier (6 % of
    // the warp time in the round-1 profile).  Both stages are free here (block sync above).
    if (tid == 0) {
      issue(jt, 0);
      if (jn < j_tile1) issue(jn, 1);
    }

    for (int stage = 0; jt < j_tile1; stage ^= 1) {
      mbar_wait(&s_bar[stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
      const float* __restrict__ sx = s_tile + stage * 3 * TJ;
      const float* __restrict__ sy = sx + TJ;
      const float* __restrict__ sz = sy + TJ;
      const int j0 = col_base + jt * TJ;
      const bool diag = same && (j0 < row_base + TI);  // tile overlaps this row tile

      // which of this warp's row groups can reach which SUB-atom sub-tile: bit (q * R + r)
      unsigned rmask = 0xffffffffu;
      unsigned umask = 0u;             // AM 7: blocks with a uniform periodic image
      unsigned smask = 0u;             // AM 7: ... whose image shift is not zero
      unsigned tmask = 0u;             // diagonal tiles: blocks that straddle j == i
      float shx = 0.f, shy = 0.f, shz = 0.f;  // AM 7: image shift of block (lane % (NSUB * R))
      // A diagonal tile (same species, columns inside this row tile) splits into blocks whose
      // pairs all have j > i (counted on the fast path like any other block), blocks with
      // j <= i throughout (skipped) and the few blocks the diagonal runs through (sub_tile_tri).
      // Block (q, r) of this warp: rows row_base + r * NT + [32 w, 32 w + 32), columns
      // j0 + q * SUB + [0, SUB).
      auto diag_status = [&](int q, int r, bool& full, bool& tri) {
        const int r_lo = r * NT + (tid & ~31);
        const int c_lo = (j0 - row_base) + q * SUB;
        full = c_lo > r_lo + 31;
        tri = !full && (c_lo + SUB - 1 > r_lo);
      };
      if constexpr (AM == 7) {
        static_assert(AM != 7 || (R == 4 && (NSUB == 4 || NSUB == 8) && CULL),
                      "AM 7 needs 4 row groups x 4 or 8 sub-tiles per tile (one block per lane)");
        constexpr unsigned BLOCKS = NSUB * R == 32 ? 0xffffffffu : (1u << (NSUB * R)) - 1u;
        {
          // lane l classifies block (q = (l >> 2) % NSUB, r = l & 3): bit l of the masks
          const float* __restrict__ cbp =
              fbox + (size_t)(col_box0 + jt * NSUB + ((tid >> 2) & (NSUB - 1))) * 6;
          float cb[6];
#pragma unroll
          for (int d = 0; d < 6; ++d) cb[d] = __ldg(cbp + d);
          bool full = true, tri = false;
          if (diag) diag_status((tid >> 2) & (NSUB - 1), tid & 3, full, tri);
          const bool live = full && (mybox[0] <= mybox[3]) &&
                            !boxes_far(mybox, cb, P.box, P.cull_eps, P.cull2);
          bool uni = live;
          float sh[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            // d = x_j - x_i of every pair of the block lies in [dlo, dhi] (rounding is monotone)
            const float ulo = (cb[d] - mybox[d + 3]) * P.inv_box[d];
            const float uhi = (cb[d + 3] - mybox[d]) * P.inv_box[d];
            const float nlo = rintf(ulo), nhi = rintf(uhi);
            uni = uni && (nlo == nhi) && (fabsf(ulo - nlo) < 0.499f) &&
                  (fabsf(uhi - nhi) < 0.499f) && (fabsf(nlo) <= 2.0f);
            sh[d] = -nlo * P.box[d];
          }
          shx = sh[0];
          shy = sh[1];
          shz = sh[2];
          rmask = __ballot_sync(0xffffffffu, live) & BLOCKS;
          umask = __ballot_sync(0xffffffffu, uni) & BLOCKS;
          smask = __ballot_sync(0xffffffffu, shx != 0.f || shy != 0.f || shz != 0.f) & umask;
          tmask = __ballot_sync(0xffffffffu, tri) & BLOCKS;
        }
      } else {
        if (CULL) {
          rmask = 0u;
#pragma unroll
          for (int q = 0; q < NSUB; ++q) {
            const float* __restrict__ cb6 = fbox + (size_t)(col_box0 + jt * NSUB + q) * 6;
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (!boxes_far(wbox[r], cb6, P.box, P.cull_eps, P.cull2)) rmask |= 1u << (q * R + r);
          }
        }
        if (diag) {
          unsigned fmask = 0u;
#pragma unroll
          for (int q = 0; q < NSUB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
              bool full, tri;
              diag_status(q, r, full, tri);
              fmask |= (full ? 1u : 0u) << (q * R + r);
              tmask |= (tri ? 1u : 0u) << (q * R + r);
            }
          rmask &= fmask;
        }
      }

      if (!EXACT && !GLOBAL_HIST) {
        constexpr unsigned FULL = (1u << R) - 1u;
#pragma unroll 1
        for (int q = 0; q < NSUB; ++q) {
          const unsigned m = (rmask >> (q * R)) & FULL;
          if constexpr (AM == 7) {
            const unsigned mu = (umask >> (q * R)) & FULL;
            const unsigned mm = m & ~mu;
            // adding a zero shift is exact, so skipping it changes nothing: fully live blocks
            // that do not cross a periodic boundary (about half of them) save three FADD2 of the
            // 28 instructions per pair couple
            if (mu == FULL && ((smask >> (q * R)) & FULL) == 0u) {
              const float zero[R] = {0.f, 0.f, 0.f, 0.f};
              sub_tile_uni<R, false>(sx, sy, sz, q * SUB, nxi, nyi, nzi, zero, zero, zero, geo);
            } else if (mu != 0u) {
              float bx[R], by[R], bz[R];
#pragma unroll
              for (int r = 0; r < R; ++r) {
                bx[r] = __shfl_sync(0xffffffffu, shx, q * R + r);
                by[r] = __shfl_sync(0xffffffffu, shy, q * R + r);
                bz[r] = __shfl_sync(0xffffffffu, shz, q * R + r);
              }
              if (mu == FULL) {
                sub_tile_uni<R, true>(sx, sy, sz, q * SUB, nxi, nyi, nzi, bx, by, bz, geo);
              } else {
#pragma unroll
                for (int r = 0; r < R; ++r)
                  if ((mu >> r) & 1u)  // warp-uniform
                    sub_tile_uni_row<true>(sx, sy, sz, q * SUB, nxi[r], nyi[r], nzi[r], bx[r],
                                           by[r], bz[r], geo);
              }
            }
            if (mm != 0u) {
#pragma unroll
              for (int r = 0; r < R; ++r)
                if ((mm >> r) & 1u)  // warp-uniform
                  sub_tile_gen_row(sx, sy, sz, q * SUB, nxi[r], nyi[r], nzi[r], geo);
            }
          } else if (m == FULL)
            sub_tile<false, R, AM, (CULL ? 2 : 4)>(m, sx, sy, sz, q * SUB, nxi, nyi, nzi, geo);
          else if (m != 0u)
            sub_tile<true, R, AM, 2>(m, sx, sy, sz, q * SUB, nxi, nyi, nzi, geo);
        }
        if (tmask != 0u) {  // diagonal tiles only
#pragma unroll 1
          for (int q = 0; q < NSUB; ++q) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if ((tmask >> (q * R + r)) & 1u)  // warp-uniform
                sub_tile_tri(sx, sy, sz, q * SUB, j0, row_base + r * NT + tid, nxi[r].x, nyi[r].x,
                             nzi[r].x, geo, s_thr, s_cnt);
          }
        }
      } else {
        // exact-division fallback and global-memory tables: scalar path (j > i mask on diagonal
        // tiles)
        for (int jj = 0; jj < TJ; ++jj) {
          const float xj = sx[jj], yj = sy[jj], zj = sz[jj];
          const int j = j0 + jj;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float rx = __fsub_rn(xj, -nxi[r].x);
            float ry = __fsub_rn(yj, -nyi[r].x);
            float rz = __fsub_rn(zj, -nzi[r].x);
            if (EXACT) {
              rx = __fsub_rn(rx, __fmul_rn(rintf(__fdiv_rn(rx, P.box[0])), P.box[0]));
              ry = __fsub_rn(ry, __fmul_rn(rintf(__fdiv_rn(ry, P.box[1])), P.box[1]));
              rz = __fsub_rn(rz, __fmul_rn(rintf(__fdiv_rn(rz, P.box[2])), P.box[2]));
            } else {
              const float qx = __fadd_rn(fmaf(rx, P.inv_box[0], RINT_MAGIC), -RINT_MAGIC);
              const float qy = __fadd_rn(fmaf(ry, P.inv_box[1], RINT_MAGIC), -RINT_MAGIC);
              const float qz = __fadd_rn(fmaf(rz, P.inv_box[2], RINT_MAGIC), -RINT_MAGIC);
              rx = fmaf(qx, -P.box[0], rx);
              ry = fmaf(qy, -P.box[1], ry);
              rz = fmaf(qz, -P.box[2], rz);
            }
            const float d2 =
                __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
            const bool ok = (d2 < cut2) && (!same || !diag || j > row_base + r * NT + tid);
            if (GLOBAL_HIST)
              bin_one_global(d2, ok, P.thr, P.hist + (size_t)cur_pair * P.nbins, inv_step);
            else
              bin_one(d2, ok, s_thr, s_cnt, inv_step);
          }
        }
      }
      __syncwarp();
      const int jnn = jn < j_tile1 ? next_live(jn + 1) : j_tile1;
      if ((tid & 31) == 0) {
        mbar_arrive(&s_ebar[stage]);  // this warp is done reading the stage
        if (atomicAdd(&s_rel[stage], 1u) == NT / 32 - 1) {  // last warp out: refill the stage
          s_rel[stage] = 0u;
          if (jnn < j_tile1) {
            mbar_wait(&s_ebar[stage], (ephase >> stage) & 1u);  // completes at once (acquire)
            issue(jnn, stage);
          }
        }
      }
      ephase ^= 1u << stage;  // one phase of the empty barrier per tile and stage
      jt = jn;
      jn = jnn;
      if (!GLOBAL_HIST && ++tiles_since_flush >= P.flush_tiles) {
        flush_hist<NT>(s_cnt, P.nbins, P.hist + (size_t)cur_pair * P.nbins);
        tiles_since_flush = 0;
      }
    }
  }
  if (cur_pair >= 0 && !GLOBAL_HIST)
    flush_hist<NT>(s_cnt, P.nbins, P.hist + (size_t)cur_pair * P.nbins);
}

// ---------------------------------------------------------------------------------------
// pack / extent kernels
// ---------------------------------------------------------------------------------------
__global__ void rdf_pack_kernel(const float* __restrict__ traj, long long T, long long atom_first,
                                long long atom_count, const int* __restrict__ frames,
                                float* __restrict__ out, long long n_pad, long long dst_first,
                                long long dst_span) {
  const int k = blockIdx.y;
  const long long f = frames[k];
  float* ox = out + (size_t)k * 3 * n_pad + dst_first;
  const float nanv = __int_as_float(0x7fc00000);
  for (long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x; a < dst_span;
       a += (long long)gridDim.x * blockDim.x) {
    float x = nanv, y = nanv, z = nanv;
    if (a < atom_count) {
      const float* src = traj + ((size_t)(atom_first + a) * T + f) * 3;
      x = __ldg(src);
      y = __ldg(src + 1);
      z = __ldg(src + 2);
    }
    ox[a] = x;
    ox[n_pad + a] = y;
    ox[2 * n_pad + a] = z;
  }
}

// Sampled frames of an atom block, kept atom-major: out[a][k][:] = traj[a][frames[k]][:].
// traj may be page-locked host memory (read in place over the host link): this is the send
// buffer of the multi-rank frame exchange.
__global__ void gather_frames_kernel(const float* __restrict__ traj, long long T, long long A,
                                     const int* __restrict__ frames, int F,
                                     float* __restrict__ out) {
  const long long total = A * F;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long a = e / F;
    const int k = static_cast<int>(e - a * F);
    const float* src = traj + ((size_t)a * T + frames[k]) * 3;
    const float x = __ldg(src), y = __ldg(src + 1), z = __ldg(src + 2);
    out[3 * e] = x;
    out[3 * e + 1] = y;
    out[3 * e + 2] = z;
  }
}

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  // valid for any sign: compare as ordered ints
  if (v >= 0.f)
    atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void coord_extent_kernel(const float* __restrict__ pos, long long n_pad,
                                    long long total, float* __restrict__ minmax) {
  // pos viewed as [F*3][n_pad]; dim = row % 3
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(pos + e);
    const int d = static_cast<int>((e / n_pad) % 3);
    if (v == v) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q == d) {
          mn[q] = fminf(mn[q], v);
          mx[q] = fmaxf(mx[q], v);
        }
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[q] = fminf(mn[q], __shfl_xor_sync(0xffffffffu, mn[q], o));
      mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (mn[q] != INFINITY) atomic_min_f(minmax + q, mn[q]);
      if (mx[q] != -INFINITY) atomic_max_f(minmax + 3 + q, mx[q]);
    }
  }
}

// Bin-edge tie census (north_star: "bit-exact away from fp32 bin-edge ties, tie count
// reported").  For the pairs (i, j > i) of the first `n_rows` atoms of a packed frame against
// all atoms, the reference's bin -- double-step rule on the correctly rounded fp32 distance,
// i.e. the threshold table the pair kernel uses -- is compared with the bin a plain fp32
// histogram would take, floor(sqrt_rn(d2) * float(nbins / cutoff)).  out[0] += pairs inside the
// cutoff, out[1] += pairs whose two bins differ: those sit on a bin edge to within fp32
// rounding, and they are the only pairs an fp32-binning implementation could count differently.
__global__ void rdf_tie_count_kernel(const float* __restrict__ pos, long long n_pad,
                                     long long n_rows, const float* __restrict__ thr, int nbins,
                                     float cut2, float inv_step, float box0, float box1,
                                     float box2, int exact, unsigned long long* __restrict__ out) {
  unsigned long long in_cut = 0, ties = 0;
  const float L[3] = {box0, box1, box2};
  const long long total = n_rows * n_pad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / n_pad, j = e - i * n_pad;
    if (j <= i) continue;
    float d2 = 0.f, sq[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float r = __fsub_rn(__ldg(pos + d * n_pad + j), __ldg(pos + d * n_pad + i));
      if (exact)
        r = __fsub_rn(r, __fmul_rn(rintf(__fdiv_rn(r, L[d])), L[d]));
      else
        r = fmaf(__fadd_rn(fmaf(r, 1.0f / L[d], RINT_MAGIC), -RINT_MAGIC), -L[d], r);
      sq[d] = __fmul_rn(r, r);
    }
    d2 = __fadd_rn(__fadd_rn(sq[0], sq[1]), sq[2]);
    if (!(d2 < cut2)) continue;  // also drops NaN padding
    ++in_cut;
    const float t = fmaf(sqrt_approx(d2), inv_step, RINT_MAGIC);
    const unsigned g = min(__float_as_uint(t) - RINT_MAGIC_BITS, (unsigned)nbins);
    const int k_ref = (int)g - (d2 >= __ldg(thr + g) ? 0 : 1);
    const int k_f32 = min((int)floorf(__fmul_rn(__fsqrt_rn(d2), inv_step)), nbins - 1);
    ties += (k_ref != k_f32);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    in_cut += __shfl_xor_sync(0xffffffffu, in_cut, o);
    ties += __shfl_xor_sync(0xffffffffu, ties, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (in_cut) atomicAdd(out, in_cut);
    if (ties) atomicAdd(out + 1, ties);
  }
}

template <int NT, int R, bool EXACT, int AM, bool CULL>
int launch_rdf(const RdfParams& P, size_t smem, int grid, cudaStream_t s) {
  auto kern = rdf_pair_hist_kernel<NT, R, EXACT, AM, CULL>;
  MDK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, NT, smem, s>>>(P);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

template <int NT, int R>
int launch_rdf_cfg(const RdfParams& P, size_t smem, int grid, cudaStream_t s, bool exact, int am) {
  if (am == 4)
    return exact ? launch_rdf<NT, R, true, 4, false>(P, smem, grid, s)
                 : launch_rdf<NT, R, false, 4, false>(P, smem, grid, s);
  if (exact) return launch_rdf<NT, R, true, 2, false>(P, smem, grid, s);
  if constexpr (NT == 256 && R == 4) {
    if (P.bbox && am == 7) return launch_rdf<NT, R, false, 7, true>(P, smem, grid, s);
    if (am == 8)
      return P.bbox ? launch_rdf<NT, R, false, 8, true>(P, smem, grid, s)
                    : launch_rdf<NT, R, false, 8, false>(P, smem, grid, s);
  }
  if (am == 8) am = 5;
  if (am == 7) am = 2;  // uniform-image blocks need the culling boxes and the 256 x 4 tile
  if (P.bbox)           // culling variants: AM 2 (table), AM 5 (wrapped)
    return am == 5 ? launch_rdf<NT, R, false, 5, true>(P, smem, grid, s)
                   : launch_rdf<NT, R, false, 2, true>(P, smem, grid, s);
  return am == 5 ? launch_rdf<NT, R, false, 5, false>(P, smem, grid, s)
                 : launch_rdf<NT, R, false, 2, false>(P, smem, grid, s);
}

}  // namespace mdk

using namespace mdk;

extern "C" int mdk_rdf_tile(void) { return TJ; }

extern "C" int mdk_rdf_thresholds(float cutoff, int nbins, float* thr, float* cut2_out) {
  MDK_CHECK_ARG(nbins >= 1 && cutoff > 0.f && thr && cut2_out, "rdf_thresholds: bad argument");
  // tf.histogram_fixed_width CPU functor: step = double(hi - lo) / double(nbins), lo = 0
  const volatile double step = static_cast<double>(cutoff) / static_cast<double>(nbins);
  auto bin_ge = [&](float d, int m) {
    volatile double q = static_cast<double>(d) / step;
    return q >= static_cast<double>(m);  // floor(q) >= m
  };
  // smallest fp32 x >= 0 with sqrtf_rn(x) >= dmin
  auto d2_threshold = [&](float dmin) {
    volatile float x = dmin * dmin;
    for (;;) {
      volatile float s = sqrtf(x);
      if (s >= dmin) break;
      x = nextafterf(x, INFINITY);
    }
    for (;;) {
      if (x <= 0.f) break;
      volatile float xm = nextafterf(x, 0.f);
      volatile float s = sqrtf(xm);
      if (s >= dmin) x = xm; else break;
    }
    return (float)x;
  };
  thr[0] = 0.f;
  for (int m = 1; m < nbins; ++m) {
    volatile float d = static_cast<float>(static_cast<double>(m) * step);
    while (!bin_ge(d, m)) d = nextafterf(d, INFINITY);
    for (;;) {
      volatile float dm = nextafterf(d, 0.f);
      if (dm >= 0.f && bin_ge(dm, m)) d = dm; else break;
    }
    thr[m] = d2_threshold(d);
  }
  thr[nbins] = INFINITY;
  *cut2_out = d2_threshold(cutoff);
  return MDK_OK;
}

extern "C" int mdk_rdf_pack(const float* traj, long long A_total, long long T, long long atom_first,
                            long long atom_count, const int* frames, int n_frames, float* out,
                            long long n_pad, long long dst_first, long long dst_span,
                            mdk_stream_t stream) {
  MDK_CHECK_ARG(traj && frames && out, "rdf_pack: null pointer");
  MDK_CHECK_ARG(atom_first >= 0 && atom_count >= 0 && atom_first + atom_count <= A_total,
                "rdf_pack: atom range [%lld, +%lld) outside %lld", atom_first, atom_count, A_total);
  MDK_CHECK_ARG(dst_span >= atom_count && dst_first >= 0 && dst_first + dst_span <= n_pad,
                "rdf_pack: destination range outside n_pad");
  MDK_CHECK_ARG(n_frames >= 0 && n_frames <= 65535, "rdf_pack: n_frames must be <= 65535 per call");
  if (n_frames == 0 || dst_span == 0) return MDK_OK;
  const int threads = 256;
  long long blocks = (dst_span + threads - 1) / threads;
  if (blocks > 4096) blocks = 4096;
  dim3 grid((unsigned)blocks, (unsigned)n_frames);
  rdf_pack_kernel<<<grid, threads, 0, as_stream(stream)>>>(traj, T, atom_first, atom_count, frames,
                                                           out, n_pad, dst_first, dst_span);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_gather_frames(const float* traj, long long A, long long T, const int* frames,
                                 int n_frames, float* out, mdk_stream_t stream) {
  MDK_CHECK_ARG(A >= 0 && T >= 0 && n_frames >= 0, "gather_frames: negative size");
  if (A == 0 || n_frames == 0) return MDK_OK;
  MDK_CHECK_ARG(traj && frames && out, "gather_frames: null pointer");
  long long blocks = (A * n_frames + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gather_frames_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(traj, T, A, frames,
                                                                         n_frames, out);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_coord_extent(const float* pos_soa, int n_frames, long long n_pad, float* minmax,
                                mdk_stream_t stream) {
  MDK_CHECK_ARG(pos_soa && minmax && n_frames >= 0 && n_pad >= 0, "coord_extent: bad argument");
  const long long total = (long long)n_frames * 3 * n_pad;
  if (total == 0) return MDK_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  coord_extent_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(pos_soa, n_pad, total, minmax);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_rdf_tie_count(const float* pos_frame, long long n_pad, long long n_rows,
                                 const float* box, float cut2, float cutoff, int nbins,
                                 const float* thr, int flags, unsigned long long* out,
                                 mdk_stream_t stream) {
  MDK_CHECK_ARG(pos_frame && box && thr && out, "rdf_tie_count: null pointer");
  MDK_CHECK_ARG(nbins >= 1 && cutoff > 0.f && cut2 > 0.f && n_pad >= 0 && n_rows >= 0 &&
                    n_rows <= n_pad, "rdf_tie_count: bad sizes");
  if (n_rows == 0 || n_pad == 0) return MDK_OK;
  const float inv_step = static_cast<float>(static_cast<double>(nbins) / static_cast<double>(cutoff));
  long long blocks = (n_rows * n_pad + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rdf_tie_count_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      pos_frame, n_pad, n_rows, thr, nbins, cut2, inv_step, box[0], box[1], box[2],
      (flags & MDK_RDF_EXACT_DIV) ? 1 : 0, out);
  MDK_LAUNCH_CHECK();
  return MDK_OK;
}

extern "C" int mdk_rdf_hist(const float* pos_soa, int n_frames, long long n_pad, const int* sp_lo,
                            const int* sp_hi, int n_species, const float* box, float cut2,
                            float cutoff, int nbins, const float* thr, unsigned long long* hist,
                            unsigned int* work_counter, const float* bbox, int flags,
                            mdk_stream_t stream) {
  MDK_CHECK_ARG(pos_soa && sp_lo && sp_hi && box && thr && hist && work_counter,
                "rdf_hist: null pointer");
  MDK_CHECK_ARG(n_species >= 1 && n_species <= MDK_MAX_SPECIES,
                "rdf_hist: n_species %d outside [1, %d]", n_species, MDK_MAX_SPECIES);
  MDK_CHECK_ARG(nbins >= 1 && cutoff > 0.f && cut2 > 0.f, "rdf_hist: bad nbins/cutoff");
  MDK_CHECK_ARG(n_frames >= 0 && n_pad % 4 == 0, "rdf_hist: n_pad must be a multiple of 4");
  for (int s = 0; s < n_species; ++s) {
    MDK_CHECK_ARG(sp_lo[s] % TJ == 0 && sp_hi[s] >= sp_lo[s] && sp_hi[s] <= n_pad,
                  "rdf_hist: species block %d [%d, %d) must start at a multiple of %d inside n_pad",
                  s, sp_lo[s], sp_hi[s], TJ);
    const long long padded_end = ((long long)(sp_hi[s] + TJ - 1) / TJ) * TJ;
    MDK_CHECK_ARG(padded_end <= n_pad && (s + 1 == n_species || padded_end <= sp_lo[s + 1]),
                  "rdf_hist: species block %d is not padded to a multiple of %d", s, TJ);
    MDK_CHECK_ARG(box[s % 3] > 0.f, "rdf_hist: box must be positive");
  }
  if (n_frames == 0) return MDK_OK;
  cudaStream_t s = as_stream(stream);

  // tile configuration: small systems use narrow row tiles.  flags bits 8..11 / 12..15
  // override the (threads, rows-per-thread) configuration / atomic mode (tuning only).
  int max_len = 0;
  for (int q = 0; q < n_species; ++q)
    max_len = max_len > sp_hi[q] - sp_lo[q] ? max_len : sp_hi[q] - sp_lo[q];
  const bool exact = (flags & MDK_RDF_EXACT_DIV) != 0;
  int cfg = (flags >> 8) & 0xf;   // 0 = auto, 1: 128x2, 2: 128x4, 3: 256x2, 4: 256x4
  int am = (flags >> 12) & 0xf;   // 0 = auto, else AM = am - 1
  if (cfg == 0) cfg = max_len <= 8192 ? 1 : 4;
  const bool auto_am = am == 0;
  am = auto_am ? 2 : am - 1;
  MDK_CHECK_ARG(cfg >= 1 && cfg <= 5 && (am == 2 || am == 4 || am == 5 || am == 7 || am == 8),
                "rdf_hist: bad tuning flags (atomic modes: 2, 4, 5, 7, 8)");
  const bool wrapped = !exact && (flags & MDK_RDF_WRAPPED);
  // coordinates verified to span less than one box length: cheaper minimum image
  if (auto_am && wrapped) am = 5;
  if (am == 5 && !wrapped) am = 2;
  // sorted frames with bounding boxes: uniform-image blocks + gated compare (any coordinates)
  if ((auto_am || am == 7) && !exact && bbox && cfg >= 4) am = 7;
  else if (am == 7) am = wrapped ? 5 : 2;
  // wrapped minimum image + clamped gated compare (needs wrapped coordinates; instantiated for
  // the 256 x 4 tile): +4..6 % over AM 5 on unsorted frames
  if (auto_am && am == 5 && cfg >= 4) am = 8;
  if (am == 8 && !wrapped) am = 2;
  if (am == 8 && cfg < 4) am = 5;
  const int NT = cfg == 5 ? 384 : (cfg <= 2) ? 128 : 256;
  const int R = (cfg == 1 || cfg == 3) ? 2 : 4;
  const int TI = NT * R;

  RdfParams P;
  memset(&P, 0, sizeof(P));
  P.pos = pos_soa;
  P.n_pad = n_pad;
  P.n_frames = n_frames;
  P.n_species = n_species;
  for (int q = 0; q < n_species; ++q) {
    P.sp_lo[q] = sp_lo[q];
    P.sp_hi[q] = sp_hi[q];
  }
  for (int d = 0; d < 3; ++d) {
    P.box[d] = box[d];
    P.inv_box[d] = 1.0f / box[d];
  }
  P.cut2 = cut2;
  P.inv_step = static_cast<float>(static_cast<double>(nbins) / static_cast<double>(cutoff));
  P.nbins = nbins;
  {
    const double mid = (static_cast<double>(nbins) + 0.5) * static_cast<double>(cutoff) /
                       static_cast<double>(nbins);
    P.clamp2 = static_cast<float>(mid * mid);
  }
  P.thr = thr;
  P.hist = hist;
  P.counter = reinterpret_cast<unsigned long long*>(work_counter);

  const size_t smem = (size_t)2 * 3 * TJ * sizeof(float) + 4 * sizeof(uint64_t) +
                      6 * sizeof(unsigned long long) +
                      (size_t)((nbins + 1 + 3) & ~3) * sizeof(float) +
                      (size_t)(((nbins + 31) & ~31) + 32) * sizeof(unsigned);
  const bool global_hist = smem > 227 * 1024 || am == 4;
  size_t smem_used = smem;
  if (global_hist) {
    // threshold table + private histogram do not fit: slow path with both in global memory
    am = 4;
    smem_used = (size_t)2 * 3 * TJ * sizeof(float) + 4 * sizeof(uint64_t) +
                6 * sizeof(unsigned long long) + 64;
  }
  // resident CTAs per SM (shared-memory bound) -> persistent grid
  int per_sm = (int)((227 * 1024) / (smem_used + 1024));
  const int max_by_threads = 2048 / NT;
  if (per_sm > max_by_threads) per_sm = max_by_threads;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int grid = sm_count() * per_sm;

  // work items: choose CJ so that there are enough items per CTA where possible
  int np = 0;
  for (int a = 0; a < n_species; ++a)
    for (int b = a; b < n_species; ++b) {
      P.pair_a[np] = a;
      P.pair_b[np] = b;
      P.row_tiles[np] = (sp_hi[a] - sp_lo[a] + TI - 1) / TI;
      P.col_tiles[np] = (sp_hi[b] - sp_lo[b] + TJ - 1) / TJ;
      ++np;
    }
  P.n_pairs = np;
  auto count_items = [&](int CJ) {
    unsigned long long tot = 0;
    for (int p = 0; p < np; ++p) {
      P.chunks_per_row[p] = (P.col_tiles[p] + CJ - 1) / CJ;
      P.item_start[p] = tot;
      tot += (unsigned long long)n_frames * P.row_tiles[p] * P.chunks_per_row[p];
    }
    P.item_start[np] = tot;
    return tot;
  };
  // (half of the items of a same-species pair lie below the diagonal and are empty; ~100 items
  // per CTA keep the tail of the persistent grid short: +2 % at 10^5 atoms x 16 frames)
  int CJ = 64;
  unsigned long long total = count_items(CJ);
  while (CJ > 1 && total < (unsigned long long)grid * 96) {
    CJ /= 2;
    total = count_items(CJ);
  }
  // stress-test overrides (tests/test_gpu_kernels.py::test_rdf_schedule_stress): bits 16..19
  // force the column chunk (CJ = 2^(v-1)), bits 20..23 shrink the persistent grid to v/15 of
  // its size -- the counts must not depend on how the work items are cut or interleaved
  int grid_used = grid;
  if (const int v = (flags >> 16) & 0xf) {
    CJ = 1 << (v - 1);
    total = count_items(CJ);
  }
  if (const int v = (flags >> 20) & 0xf) {
    grid_used = (int)((long long)grid * v / 15);
    if (grid_used < 1) grid_used = 1;
  }
  P.CJ = CJ;
  P.total_items = total;
  // u32 private counters: flush before a bin could overflow (TI * TJ pairs per tile)
  P.flush_tiles = (unsigned)((1ull << 31) / ((unsigned long long)TI * TJ));
  P.one = 1u;
  P.onef = 1.0f;
  {
    const char* e = getenv("MDK_RDF_DUMP_SHARED");
    P.dump_shared = (e && e[0] == '0') ? 0 : 1;  // default: one shared dump word (POPC.INC
                                                 // merges same-address lanes; +0.2..1.7 %)
  }
  P.bbox = exact ? nullptr : bbox;  // culling only on the fast minimum-image path
  P.boxes_per_frame = (int)(n_pad / SUB);
  P.cull2 = cut2 * 1.0001f;
  for (int d = 0; d < 3; ++d) P.cull_eps[d] = 1e-5f * box[d];
  MDK_CHECK_ARG(!bbox || n_pad % TJ == 0, "rdf_hist: bbox needs n_pad to be a multiple of %d", TJ);

  MDK_CUDA(cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), s));
  switch (cfg) {
    case 1: return launch_rdf_cfg<128, 2>(P, smem_used, grid_used, s, exact, am);
    case 2: return launch_rdf_cfg<128, 4>(P, smem_used, grid_used, s, exact, am);
    case 3: return launch_rdf_cfg<256, 2>(P, smem_used, grid_used, s, exact, am);
    case 5:  // tuning: 12 warps per CTA (AM 7 / 8 only)
      if (am == 7 && P.bbox) return launch_rdf<384, 4, false, 7, true>(P, smem_used, grid_used, s);
      if (am == 8 && !P.bbox) return launch_rdf<384, 4, false, 8, false>(P, smem_used, grid_used, s);
      set_error("rdf_hist: tile configuration 5 needs atomic mode 7 (sorted) or 8 (unsorted)");
      return MDK_EINVAL;
    default: return launch_rdf_cfg<256, 4>(P, smem_used, grid_used, s, exact, am);
  }
}
