/*
 * mdk.h -- C ABI of libmdk.so, the B200 (sm_100a) kernels behind the MDSuite
 * calculator API (SamTov/LAMMPS-Analysis).
 *
 * The reference is pure Python on TensorFlow and has no FFI of its own
 * (SURVEY.md section 8b); this is the boundary introduced *below* its calculator
 * classes.  Every entry point names the reference code whose arithmetic it
 * replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - All array pointers are DEVICE pointers owned by the caller unless the
 *     parameter is documented as HOST.
 *   - The library never allocates or frees caller-visible memory, never
 *     synchronises the stream and keeps no global mutable state except a
 *     thread-local last-error string.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Return 0 on success, a negative MDK_E* code otherwise; mdk_last_error()
 *     then describes the failure.
 *   - Accumulating outputs (histograms, sums) are `+=`: zero them first.
 */
#ifndef MDK_H_
#define MDK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MDK_VERSION 100

#define MDK_OK 0
#define MDK_EINVAL -1      /* bad argument */
#define MDK_ECUDA -2       /* CUDA runtime / launch failure */
#define MDK_EUNSUPPORTED -3 /* valid request outside what the kernels support */

#define MDK_MAX_SPECIES 8
#ifndef MDK_RDF_SUBTILE
#define MDK_RDF_SUBTILE 32 /* atoms per bounding box of mdk_rdf_bbox */
#endif

/* flags for mdk_rdf_hist */
#define MDK_RDF_EXACT_DIV 1 /* true fp32 division + separate mul/sub in the minimum image
                               (required when cutoff >= min(box)/2 or coordinates span
                               more than 2.5 box lengths); default is the fast path that
                               is bit-identical under those preconditions */

#define MDK_RDF_WRAPPED 2   /* the caller guarantees that, per dimension, the coordinates span
                               less than one box length (e.g. wrapped into [0, L)): enables
                               the min(|d|, L - |d|) minimum image, bit-identical there */

typedef void* mdk_stream_t;

int mdk_version(void);
const char* mdk_last_error(void);
/* Number of SMs of the current device (used by hosts to size batches). */
int mdk_sm_count(void);

/* ------------------------------------------------------------------------- *
 * RDF  (mdsuite/calculators/radial_distribution_function.py)
 * ------------------------------------------------------------------------- */

/* Rows of one RDF row tile; species blocks in the packed frame array must start
 * at multiples of this and be padded with NaN up to the next multiple. */
int mdk_rdf_tile(void);

/* HOST helper.  Fills thr_host[0..nbins] with the fp32 thresholds on the squared
 * distance that reproduce tf.histogram_fixed_width exactly:
 *   bin(d2) = #{ m in 1..nbins-1 : d2 >= thr[m] },  thr[0] = 0, thr[nbins] = +inf
 * and *cut2_host with the smallest fp32 d2 whose correctly rounded sqrt is
 * >= (float)cutoff, so that `d < cutoff` <=> `d2 < cut2`.
 * Replaces: radial_distribution_function.py:616-645 (bin_minibatch),
 *           utils/linalg.py:125-136 (apply_system_cutoff),
 *           tensorflow/core/kernels/histogram_op.cc (CPU functor). */
int mdk_rdf_thresholds(float cutoff, int nbins, float* thr_host, float* cut2_host);

/* Gather + transpose atom-major trajectory rows into the frame-major SoA layout
 * the pair kernel streams:  out[k][d][dst_first + a] = traj[(atom_first + a)][frames[k]][d]
 * for a < atom_count, and NaN for dst_first + atom_count <= idx < dst_first + dst_span.
 *   traj   : [A_total][T][3] fp32 (MDSuite store layout, simulation_database.py:364-368)
 *   frames : device int32 [n_frames]
 *   out    : [n_frames][3][n_pad] fp32
 * Replaces: data_manager.py:195-201 (frame fancy-index load) and
 *           radial_distribution_function.py:535-563 (_format_data concat). */
int mdk_rdf_pack(const float* traj, long long A_total, long long T, long long atom_first,
                 long long atom_count, const int* frames, int n_frames, float* out,
                 long long n_pad, long long dst_first, long long dst_span, mdk_stream_t stream);

/* Sampled frames of an atom block, kept atom-major:  out[a][k][:] = traj[a][frames[k]][:].
 *   traj   : [A][T][3] fp32, device memory OR page-locked host memory (read in place)
 *   frames : device int32 [n_frames], values in [0, T)
 *   out    : [A][n_frames][3] fp32
 * The send buffer of the multi-rank frame exchange (each rank holds an atom block of the
 * store and the RDF shards frames).  Replaces data_manager.py:195-201 (frame fancy index). */
int mdk_gather_frames(const float* traj, long long A, long long T, const int* frames,
                      int n_frames, float* out, mdk_stream_t stream);

/* Per-dimension min / max of a packed frame array (NaN padding ignored).
 * minmax: device float[6] = {minx,miny,minz,maxx,maxy,maxz}, caller-initialised to
 * {+inf x3, -inf x3}.  Used by the host to decide whether the fast minimum-image
 * path is bit-exact (coordinate extent < 2.5 box lengths). */
int mdk_coord_extent(const float* pos_soa, int n_frames, long long n_pad, float* minmax,
                     mdk_stream_t stream);

/* All-pairs minimum-image distance histogram for every species pair a <= b.
 *   pos_soa : [n_frames][3][n_pad] fp32, species blocks [sp_lo[s], sp_hi[s]) (HOST
 *             arrays, multiples of mdk_rdf_tile()), NaN-padded
 *   box     : HOST float[3]
 *   thr     : device float[nbins+1] from mdk_rdf_thresholds; cut2 likewise
 *   hist    : device u64 [n_pairs][nbins], pair order = combinations_with_replacement
 *             (0,0),(0,1),...,(1,1),... ; accumulated (+=)
 *   work_counter : device scratch, 8 bytes, 8-byte aligned; zeroed by the call
 *   bbox    : NULL, or the tile boxes from mdk_rdf_bbox (enables block culling and, for species
 *             blocks beyond 8192 atoms, the uniform-image fast path: blocks of pairs whose boxes
 *             prove one common periodic image take d + (-n L) instead of the per-pair rint)
 *   flags   : MDK_RDF_* bits; bits 8..11 (tile configuration), 12..15 (binning / atomic mode),
 *             16..19 (column chunk of a work item, 2^(v-1) tiles) and 20..23 (persistent grid
 *             shrunk to v/15) are overrides used by the tests and benchmarks, 0 = automatic
 * Pairs counted: i < j within a species block, all (i, j) across blocks, for each
 * frame: r = p_j - p_i; r -= rint(r / L) * L; d2 = (x*x + y*y) + z*z (fp32, each op
 * rounded); counted iff sqrt(d2) < cutoff.
 * Replaces: utils/linalg.py:84-122 (min image, triu indices),
 *           radial_distribution_function.py:422-524, 616-689, 846-885. */
int mdk_rdf_hist(const float* pos_soa, int n_frames, long long n_pad, const int* sp_lo,
                 const int* sp_hi, int n_species, const float* box, float cut2, float cutoff,
                 int nbins, const float* thr, unsigned long long* hist,
                 unsigned int* work_counter, const float* bbox, int flags, mdk_stream_t stream);

/* Bin-edge tie census of one packed frame ([3][n_pad], as mdk_rdf_hist takes it): for the pairs
 * (i, j > i) of its first n_rows atoms against all atoms,
 *   out[0] += pairs inside the cutoff,
 *   out[1] += pairs whose reference bin (threshold table = tf.histogram_fixed_width's
 *             double-step rule on the correctly rounded fp32 distance) differs from the bin of
 *             a plain fp32 histogram, floor(sqrt_rn(d2) * float(nbins / cutoff)).
 * out: device u64[2], accumulated.  flags: MDK_RDF_EXACT_DIV as for mdk_rdf_hist.  The pair
 * kernel itself always takes the reference bin; this is the "tie count reported" of the parity
 * criterion (radial_distribution_function.py:616-645). */
int mdk_rdf_tie_count(const float* pos_frame, long long n_pad, long long n_rows, const float* box,
                      float cut2, float cutoff, int nbins, const float* thr, int flags,
                      unsigned long long* out, mdk_stream_t stream);

/* Spatially ordered variant of mdk_rdf_pack for ONE frame of one species: the atoms are written
 * in Hilbert-curve order of a 128^3 cell grid (the histogram does not depend on the order of the atoms
 * inside a species block), NaN padded as above.  `workspace` is device scratch of at least
 * mdk_rdf_sort_workspace(atom_count) bytes; box is a HOST float[3].
 *   out_frame : [3][n_pad] fp32 (the frame's slab of the packed array) */
long long mdk_rdf_sort_workspace(int max_atoms);
int mdk_rdf_pack_sorted(const float* traj, long long A_total, long long T, long long atom_first,
                        int atom_count, long long frame, float* out_frame, long long n_pad,
                        long long dst_first, int dst_span, const float* box, void* workspace,
                        long long workspace_bytes, mdk_stream_t stream);

/* The same for a whole launch batch of frames of one species with ONE radix sort (the batch-local
 * frame number rides above the Hilbert index in the key): for systems of ~10^5 atoms the
 * per-frame sorts are launch-bound.  frames: device int[n_frames]; out: the packed array
 * [n_frames][3][n_pad] (slab k receives frames[k]).  mdk_rdf_sort_batch_workspace returns the
 * scratch size, or -1 when the batch is too large for this path (more than 2048 frames or 2^25
 * atom-frames: use the per-frame call). */
long long mdk_rdf_sort_batch_workspace(int max_atoms, int n_frames);
int mdk_rdf_pack_sorted_batch(const float* traj, long long A_total, long long T,
                              long long atom_first, int atom_count, const int* frames,
                              int n_frames, float* out, long long n_pad, long long dst_first,
                              int dst_span, const float* box, void* workspace,
                              long long workspace_bytes, mdk_stream_t stream);

/* Bounding boxes {min xyz, max xyz} of every MDK_RDF_SUBTILE-atom run of a packed frame array
 * (NaN padding ignored): bbox device float [n_frames][n_pad / MDK_RDF_SUBTILE][6].  Passed to
 * mdk_rdf_hist they let the kernel skip (row group, column sub-tile) blocks whose minimum-image
 * box distance exceeds the cutoff -- the counts are unchanged, only provably empty blocks are
 * skipped. */
int mdk_rdf_bbox(const float* pos_soa, int n_frames, long long n_pad, float* bbox,
                 mdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Angular distribution function
 * (mdsuite/calculators/angular_distribution_function.py, utils/neighbour_list.py)
 * ------------------------------------------------------------------------- */

/* Bytes of device scratch mdk_adf_hist needs for a batch (cell lists + cell-ordered copy of
 * the positions); -1 on bad arguments.  box is a HOST float[3]. */
long long mdk_adf_workspace(long long n_atoms, int n_frames, const float* box, float cutoff);

/* Triplet-angle histograms of a batch of frames.
 *   pos      : [n_frames][n_atoms][3] fp32, species concatenated in order (device)
 *   sp_hi    : HOST int[n_species], exclusive end of each species block (sp_hi[last] = n_atoms)
 *   box      : HOST float[3]; coordinates may lie outside the box
 *   cutoff   : neighbour cutoff; the test is the reference's half(|r|) < half(cutoff), |r| != 0
 *   nbins, range_hi : uniform bins on [0, range_hi] (the reference uses 3.15), numpy.histogram's
 *              bin rule on fp32 edges
 *   norm_power : weight of a triple = 1 / (|r_ij| |r_ik|)^norm_power
 *   capacity : neighbours per centre atom held in shared memory; when a centre has more,
 *              *overflow is raised to that count (atomicMax), the centre is skipped, and the
 *              caller repeats the batch with a larger capacity
 *   hist_w   : device fp64 [n_combos][nbins], += sum of weights; hist_c: u64, += triple counts;
 *              combo order = combinations_with_replacement(species, 3) as (centre, j, k)
 *   overflow : device int, caller-zeroed
 * For every centre i and every ORDERED pair of distinct neighbours (j, k) whose species
 * satisfy s_i <= s_j <= s_k: angle = acos(clip(u_ij . u_ik)), u = r / |r|,
 * r_ij = (p_i - p_j) - rint((p_i - p_j) / L) * L in fp32.
 * Replaces: utils/neighbour_list.py:53-177 (all-pairs r_ij, n^3 roll-and-compare triplets),
 *           utils/linalg.py:30-81 (get_angles),
 *           angular_distribution_function.py:302-403 (r_ij matrix, species masks, histogram). */
int mdk_adf_hist(const float* pos, int n_frames, long long n_atoms, const int* sp_hi,
                 int n_species, const float* box, float cutoff, int nbins, double range_hi,
                 double norm_power, int capacity, double* hist_w, unsigned long long* hist_c,
                 int* overflow, void* workspace, long long workspace_bytes, mdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Einstein MSD / Green-Kubo ACF
 * ------------------------------------------------------------------------- */

/* Windowed single-origin mean-square displacement.
 *   traj : [A][T][3] fp32 (atom-major), atoms [a_lo, a_hi), frames [t0, t0 + B)
 *   windows start at t0 + e*ct, e < W;  lags tau[k] (device int32 [n_tau], all < span;
 *   span = data_range, the window length in frames)
 *   msd_sum[k] += sum_w sum_a sum_d (x[a, s_w + tau_k, d] - x[a, s_w, d])^2   (fp64)
 * Replaces: einstein_diffusion_coefficients.py:168-190 (ensemble_operation) and the
 *           window loop :230-244 / data_manager.py:309-339. */
int mdk_msd_windowed(const float* traj, long long A, long long T, long long a_lo, long long a_hi,
                     long long t0, int W, int ct, const int* tau, int n_tau, int span,
                     double* msd_sum, mdk_stream_t stream);

/* Same sums for the dense lag set tau = 0 .. n_lags-1 with correlation_time 1 (the default
 * `tau_values = np.s_[:]`): register-ring kernel, one position load per n_lags/… updates.
 *   msd_sum[k] += sum_{w<W} sum_a sum_d (x[a, t0+w+k, d] - x[a, t0+w, d])^2,  k < n_lags */
int mdk_msd_dense(const float* traj, long long A, long long T, long long a_lo, long long a_hi,
                  long long t0, int W, int n_lags, double* msd_sum, mdk_stream_t stream);

/* Lag products for the windowed unbiased autocorrelation.
 *   P[t - t0][m] += sum_{a in [a_lo,a_hi)} sum_d v[a,t,d] * v[a,t+m,d]
 *   for t0 <= t < t0 + B, 0 <= m < N, t + m < t0 + B          (P: device f64 [B][N])
 * Replaces the per-window FFT of tfp.stats.auto_correlation in
 * green_kubo_self_diffusion_coefficients.py:191-199 and
 * green_kubo_ionic_conductivity.py:201-203 (see mdk_acf_windows). */
int mdk_acf_lagprod(const float* traj, long long A, long long T, long long a_lo, long long a_hi,
                    long long t0, int B, int N, double* P, mdk_stream_t stream);

/* Window sums from the lag products:
 *   S_w[m] = (1/(N-m)) * sum_{t = w*ct}^{w*ct + N-1-m} P[t][m],  w < W
 *   acf_sum[m] += sum_w S_w[m];  if acf_win != NULL: acf_win[w][m] = S_w[m]
 * which equals sum over atoms and dims of tfp.stats.auto_correlation(window,
 * axis=1, normalize=False, center=False).  P is overwritten by its prefix sum in t. */
int mdk_acf_windows(double* P, int B, int N, int W, int ct, double* acf_sum, double* acf_win,
                    mdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Transformations
 * ------------------------------------------------------------------------- */

/* Box-jump unwrapping along time, per (atom, dim):
 *   jump_t = round_half_even((p_t - p_{t-1}) / L)   (fp64, p_{-1} = carry_pos or p_0)
 *   img_t  = carry_img - sum_{u<=t} jump_u
 *   out_t  = (float)((double)p_t + img_t * L)
 *   pos/out : [A][T][3] fp32; box HOST double[3]
 *   carry_pos : device float [A][3] or NULL (first batch); updated to p_{T-1}
 *   carry_img : device double [A][3], in/out (zero it for the first batch)
 * Replaces: transformations/unwrap_coordinates.py:51-81. */
int mdk_unwrap(const float* pos, long long A, long long T, const double* box,
               float* carry_pos, int have_carry, double* carry_img, float* out,
               mdk_stream_t stream);

/* out = (float)((double)pos + (double)img * L).  img: [A][T][3] fp32.
 * Replaces: transformations/unwrap_via_indices.py:49-57. */
int mdk_unwrap_indices(const float* pos, const float* img, long long n_atom_frames,
                       const double* box, float* out, mdk_stream_t stream);

/* Forward-difference velocities of one species: out[a][t] = (pos[a][t + 1] - pos[a][t]) / dt
 * in fp32 for t < T - 1, out[a][T - 1] = out[a][T - 2] (zero when T == 1).
 *   pos, out : [A][T][3] fp32;  dt = float(time_step) * float(sample_rate)
 * Replaces transformations/velocity_from_positions.py:62-77. */
int mdk_velocity_from_positions(const float* pos, long long A, long long T, float dt, float* out,
                                mdk_stream_t stream);

/* J[t][d] += sum_a q_a * v[a][t][d]   (fp64 accumulation)
 *   q_mode 0: scalar *q (HOST double), 1: q device float [A], 2: q device float [A][T]
 * Replaces: transformations/ionic_current.py:48-58. */
int mdk_ionic_current(const float* vel, long long A, long long T, const void* q, int q_mode,
                      double* J, mdk_stream_t stream);

/* J[t][k] += sum over atoms of w(a, t) * x[a][t][comp0 + k], k < 3 (fp64 accumulation):
 *   x  : device fp32 [A][T][ncomp];  w1, w2 : NULL (w = 1) or device fp32 [A][T] (w = w1 + w2,
 *        w2 may be NULL);  J : device double [T][3], accumulated (+=)
 * Replaces MomentumFlux.transform_batch (transformations/momentum_flux.py:45-55: ncomp 6,
 * comp0 3, no weights) and IntegratedHeatCurrent.transform_batch
 * (transformations/integrated_heat_current.py:49-60: unwrapped positions weighted by
 * Kinetic_Energy + Potential_Energy). */
int mdk_flux_sum(const float* x, long long A, long long T, int ncomp, int comp0, const float* w1,
                 const float* w2, double* J, mdk_stream_t stream);

/* J[t][k] += sum over atoms of (KE + PE) v_k - (S v)_k with the symmetric stress tensor S from
 * the six components (xx, yy, zz, xy, xz, yz):  stress [A][T][6], vel [A][T][3], ke / pe [A][T],
 * all device fp32; J device double [T][3], accumulated.
 * Replaces ThermalFlux.transform_batch (transformations/thermal_flux.py:51-92). */
int mdk_thermal_flux(const float* stress, const float* vel, const float* ke, const float* pe,
                     long long A, long long T, double* J, mdk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Ingest (HOST code, no GPU involved): LAMMPS text dump tokenizer
 * ------------------------------------------------------------------------- */

/* Scans the first frame header and counts the lines of a LAMMPS dump.
 *   steps[2] : TIMESTEP of the first two frames (sample rate = steps[1] - steps[0])
 *   box[6]   : xlo xhi ylo yhi zlo zhi;  columns: the names after "ITEM: ATOMS"
 * Replaces: file_io/lammps_trajectory_files.py:100-243 (metadata). */
int mdk_lammps_scan(const char* path, long long* n_atoms, long long* n_frames, long long* steps,
                    double* box, char* columns, int columns_cap);

/* Reads n frames starting at byte *offset (0 first; updated) into out[n][n_atoms][n_cols]
 * float64.  Numbers via strtod (== Python float()), non-numeric tokens -> NaN, rows stably
 * sorted by column id_col per frame unless `sorted`.
 * Replaces: file_io/tabular_text_files.py:122-220 (_read_process_n_configurations). */
int mdk_lammps_read(const char* path, long long n_atoms, int n_cols, int id_col, int sorted,
                    long long n, long long* offset, double* out);

/* ------------------------------------------------------------------------- *
 * Measurement helpers (bench.py / tests only)
 * ------------------------------------------------------------------------- */

/* Runs an FP32 FMA throughput kernel (packed FFMA2 if packed != 0) and returns the
 * measured TFLOP/s in *tflops (2 flop per FMA).  Synchronises the device. */
int mdk_peak_fp32(int packed, int iters, double* tflops);

/* Store `nbytes` of a device array into page-locked host memory that is mapped into the device
 * address space (cudaHostAlloc / cudaHostRegister under unified addressing), with SM stores
 * instead of a copy engine: a result read-back that does not queue behind bulk transfers in
 * flight on the copy engines.  Both pointers 16-byte aligned.  Replaces the `.numpy()` reads of
 * the reference's result tensors (einstein_diffusion_coefficients.py:236-248,
 * green_kubo_self_diffusion_coefficients.py:323-337) on the streamed path. */
int mdk_store_mapped(const void* src, void* dst_host_mapped, long long nbytes, mdk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MDK_H_ */
