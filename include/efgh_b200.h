/* efgh_b200 - C ABI of the B200-native permutohedral-lattice / bilateral-convolution hot path.
 *
 * This is the drop-in boundary.  In the reference the only native FFI on this path is
 * the per-key hash map
 *     void*     khash_int2int_init(void);                          reference lib/khash_int2int.h:8
 *     void      khash_int2int_destroy(void*);                      reference lib/khash_int2int.h:12
 *     khint64_t khash_int2int_get(void*, khint64_t, khint64_t);    reference lib/khash_int2int.h:17
 *     int       khash_int2int_set(void*, khint64_t, khint64_t);    reference lib/khash_int2int.h:24
 * declared for cffi at reference lib/build_khash_cffi.py:6-13 and called one key at a time from numba
 * (reference nets/transforms.py:9-15,125-184).  A GPU cannot be fed through a per-key call, so the
 * boundary moves up one level: each entry point below replaces one whole reference function, cited at
 * its declaration.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; plain pointers + sizes only;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no call synchronises;
 *   - the library never allocates device memory: outputs and workspaces are caller-allocated;
 *   - counts that the reference returns as Python ints (pc1_hash_cnt) live in device memory so that a
 *     whole 5-level scan can be enqueued (or graph-captured) without a host round trip.  Every
 *     `*_dev` count argument may be NULL, in which case the host-side value next to it is exact;
 *     otherwise the host-side value is a CAPACITY and the kernel reads the true count on the device;
 *   - matrices are passed with explicit leading dimensions (element strides) so that tensors narrowed
 *     from over-allocated buffers can be used in place;
 *   - return value: 0 on success, a negative EFGH_E* code otherwise; efgh_last_error() gives the text
 *     (thread-local).  No CPU fallback exists: without a CUDA device every compute entry point fails.
 *   - only lattice dimension d = 3 (d+1 = 4) is implemented - the only value EFGHNet uses
 *     (reference configs/train_rellis.yaml:28).
 */
#ifndef EFGH_B200_H
#define EFGH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFGH_OK 0
#define EFGH_EINVAL (-1)   /* bad argument */
#define EFGH_ECUDA (-2)    /* CUDA runtime error (text in efgh_last_error) */
#define EFGH_ENOMEM (-3)   /* workspace too small */

/* bits of the per-level device status word (efgh_lattice_state.status) */
#define EFGH_ST_KEY_RANGE 1   /* a lattice coordinate fell outside +-2^20: cloud far outside the supported box */
#define EFGH_ST_VERTEX_CAP 2  /* more distinct vertices than the caller's capacity */
#define EFGH_ST_TABLE_FULL 4  /* hash table capacity exhausted (workspace sized for fewer points) */
#define EFGH_ST_ALIASED 8     /* informational, not an error: a neighbour key outside the key box was looked up through the
                                 reference's mixed-radix key2int aliasing (transforms.py:62-78), so blur_neighbors may not
                                 be symmetric (tap f of h is g  <=>  tap mirror(f) of g is h).  Clear = symmetric. */

/* Device-resident per-level record; the caller may copy it back (24 x int32) after the level. */
typedef struct {
  int32_t n;          /* points that entered the level */
  int32_t hash_cnt;   /* distinct lattice vertices = reference pc1_hash_cnt (generate_data.py:139) */
  int32_t status;     /* EFGH_ST_* bits, 0 = ok */
  int32_t table_mask; /* internal */
  int32_t key_min[4]; /* reference key_mins (generate_data.py:136) */
  int32_t key_max[4]; /* reference key_maxs (generate_data.py:135) */
  int32_t tile_counter;
  int32_t reserved[11];
} efgh_lattice_state;

const char *efgh_last_error(void);
int efgh_version(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int efgh_device_sm_count(void);
/* kernels launched by this library in this process so far (bench.py's `gpu_launches` is a difference of two reads;
 * launches replayed from a captured CUDA graph are not re-counted) */
int64_t efgh_launch_count(void);

/* Strided host <-> device copy of a (rows, cols) float32 matrix (leading dimensions in elements);
 * kind 1 = host -> device, 2 = device -> host.  Packs one scan's (3,n) / (C,n) host matrix into its column
 * range of a batch's device matrix (see "Ragged batch" below); the host side should be pinned. */
int efgh_copy_matrix_async(void *dst, int64_t dst_ld, const void *src, int64_t src_ld, int64_t rows, int64_t cols,
                           int kind, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Lattice build.  Together efgh_lattice_points + efgh_lattice_vertices replace one iteration of the
 * loop in GenerateData.__call__ (reference nets/generate_data.py:128-184): get_keys_and_barycentric
 * (:56-112), the key box (:135-136), the unique count (:138-139), numba build_it
 * (reference nets/transforms.py:125-184) with its two khash tables, and the next-level coordinates
 * (:175-178).
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of workspace one level needs for up to n_cap input points (and up to 4*n_cap key slots). */
size_t efgh_lattice_workspace_bytes(int64_t n_cap);

/* Phase 1: per-point elevation / rank / barycentric weights (generate_data.py:56-112), insertion of
 * the 4 simplex vertices of every point into a device hash table that keeps, per key, the smallest
 * position in the point-major / remainder-minor stream (the order build_it walks, transforms.py:152-166),
 * then a single-pass scan over first occurrences -> insertion-ordered vertex indices.
 *   pts (3,n) rows with stride pts_ld, multiplied by `scale` first (generate_data.py:130);
 *   barycentric, el_minus_gr: (4,n) f32 rows with stride out_ld;
 *   state: device record, fully written by this call (state->hash_cnt = number of vertices);
 *   h_cap: capacity the caller will give the vertex-side arrays (status bit if exceeded). */
int efgh_lattice_points(const float *pts, int64_t pts_ld, int64_t n, const int32_t *n_dev, float scale,
                        float *barycentric, float *el_minus_gr, int64_t out_ld, int64_t h_cap,
                        efgh_lattice_state *state, void *workspace, size_t workspace_bytes, void *stream);

/* Phase 2 (needs phase 1's workspace intact): per-point lattice offsets, per-vertex blur neighbours
 * and next-level coordinates.
 *   lattice_offset (4,n) int64 / lattice_offset32 (4,n) int32, rows with stride off_ld (either may be
 *     NULL): pc1_lattice_offset (transforms.py:166);
 *   filter_offsets (F,4) int32 on the device, the reference's traversal order
 *     (reference nets/transforms.py:95-122); F <= 0 skips the neighbour table (radius -1);
 *   blur_neighbors (F,h) int64 / blur_neighbors32 (F,h) int32, rows with stride nbr_ld (either may be
 *     NULL); entry -1 = vertex absent, exactly as khash_int2int_get(..., -1) on the packed key
 *     (transforms.py:168-180), including the mixed-radix aliasing of keys outside the key box;
 *   next_pts (3,h) f32 rows with stride next_ld or NULL: E^T (key / next_divisor)
 *     (generate_data.py:177-178), next_divisor = float(expected_std * scale);
 *   n, h: capacities of the point / vertex side arrays; the true counts are read from `state`. */
int efgh_lattice_vertices(int64_t n, int64_t *lattice_offset, int32_t *lattice_offset32, int64_t off_ld,
                          const int32_t *filter_offsets, int F, int64_t h,
                          int64_t *blur_neighbors, int32_t *blur_neighbors32, int64_t nbr_ld,
                          float *next_pts, int64_t next_ld, float next_divisor,
                          efgh_lattice_state *state, void *workspace, size_t workspace_bytes, void *stream);

/* ---- Ragged batch: B scans per launch sequence (the reference is batch-size-1 only,
 * nets/enet.py:107, nets/bilateralNN.py:162-165).  The point streams of the scans are concatenated: scan b owns
 * points [scan_start[b], scan_start[b+1]) of `pts` / `barycentric` / `el_minus_gr` / `lattice_offset`.  Every scan
 * keeps its own hash table, key box and insertion order, so per scan the outputs equal the single-scan calls';
 * only the vertex numbering is global: scan b's vertices are [vertex_start[b], vertex_start[b+1]), and
 * lattice_offset / blur_neighbors hold global vertex indices (local index + vertex_start[b]; -1 stays -1).
 * The BCL entry points below then treat the whole batch as one lattice, unchanged.
 *   scan_start: (B+1) int32 ON THE DEVICE, ascending, scan_start[0] = 0, no empty scan; for level l+1 pass
 *     level l's vertex_start (= the first B+1 words of its batch_info);
 *   table_entries: hash-table entries reserved for EACH scan, a power of two; efgh_lattice_table_entries(n, h)
 *     gives the size for scans of at most n points / h vertices (2 x min(4n, h), i.e. load factor <= 0.5 - small
 *     tables keep a whole batch's probes in the L2); a scan that outgrows its table sets EFGH_ST_TABLE_FULL;
 *   n_cap_total: capacity of the concatenated arrays (the true total is read from scan_start[B]);
 *   batch_info: efgh_lattice_batch_info_ints(B) int32 on the device, written by efgh_lattice_points_batch
 *     (words [0, B] = vertex_start) and read by efgh_lattice_vertices_batch;
 *   state: totals over the batch (n, hash_cnt, status); its key box is unused. 1 <= B <= 64;
 *   vertex_offsets (efgh_lattice_vertex_offsets_ints(h_cap) int32) / contributions (4*n_cap_total int32), both
 *     optional (NULL): the transposed view of lattice_offset as vertex -> contribution lists for the gather-form
 *     splat (efgh_bcl_splat_gather): contributions[vertex_offsets[h] .. vertex_offsets[h+1]) are the stream
 *     positions 4*point + remainder whose lattice_offset is h, in arbitrary order; behind the h_cap+1 offsets the
 *     array carries the list of vertices with more than 128 contributions.  contributions needs lattice_offset or
 *     lattice_offset32 to be requested too, and vertex_offsets to have been given to the points call of the level;
 *   point_rows (n_cap_total, 8) f32, optional: point-major copy [el_minus_gr[0..3], barycentric[0..3]] per point. */
int64_t efgh_lattice_vertex_offsets_ints(int64_t h_cap);
int64_t efgh_lattice_table_entries(int64_t n_scan, int64_t h_scan);
size_t efgh_lattice_batch_workspace_bytes(int B, int64_t table_entries, int64_t n_cap_total);
int64_t efgh_lattice_batch_info_ints(int B);
int efgh_lattice_points_batch(const float *pts, int64_t pts_ld, int64_t n_cap_total, const int32_t *scan_start, int B,
                              int64_t table_entries, float scale, float *barycentric, float *el_minus_gr, int64_t out_ld,
                              int64_t h_cap, efgh_lattice_state *state, int32_t *batch_info, int32_t *vertex_offsets,
                              float *point_rows, void *workspace, size_t workspace_bytes, void *stream);
int efgh_lattice_vertices_batch(int64_t n_cap_total, const int32_t *scan_start, int B, int64_t table_entries,
                                int64_t *lattice_offset, int32_t *lattice_offset32, int64_t off_ld,
                                const int32_t *filter_offsets, int F, int64_t h,
                                int64_t *blur_neighbors, int32_t *blur_neighbors32, int64_t nbr_ld,
                                float *next_pts, int64_t next_ld, float next_divisor,
                                efgh_lattice_state *state, int32_t *batch_info, int32_t *contributions, void *workspace,
                                size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Bilateral convolution layer pieces (reference nets/bilateralNN.py:148-263).  Lattice-side feature
 * matrices are VERTEX-MAJOR: row h holds the C channels of vertex h (row 0 of a "sink" matrix is the
 * all-zero row that absent neighbours (-1) read, bilateralNN.py:183-184).  Point-side matrices are
 * addressed as base[c*stride_c + n*stride_n] so both the reference's (C,N) layout and a point-major
 * layout work.  idx_bits is 64 (reference int64 tensors) or 32.
 * ---------------------------------------------------------------------------------------------- */

/* efgh_bcl_scatter with E-Net's pointwise stem fused in (reference nets/enet.py:24-28,111,
 * nets/net_utils.py:35-43): the second feature source is COMPUTED per point,
 *   act(W3 act(W2 act(W1 p + b1) + b2) + b3),  act = LeakyReLU(leaky_slope) (0 = ReLU),  p = pts[0..c_in, i],
 * so `conv_in`'s (1, 32, N) output and the torch.cat with el_minus_gr never exist in memory.  The matrix gets
 * C + c3 columns.  stem_weights (device): W1 (c1 x c_in) b1 (c1) W2 (c2 x c1) b2 (c2) W3 (c3 x c2) b3 (c3),
 * row-major = the Conv1d weights' (out, in, 1) layout; efgh_bcl_stem_weight_floats gives the length.
 * pts: (c_in, n) rows with stride pts_ld - the UNSCALED cloud (reference enet.py:109-111).  c_in <= 4, layers <= 32. */
int64_t efgh_bcl_stem_weight_floats(int c_in, int c1, int c2, int c3);
int efgh_bcl_scatter_stem(const float *feat, int64_t stride_c, int64_t stride_n, int C, const float *pts, int64_t pts_ld,
                          int c_in, int c1, int c2, int c3, const float *stem_weights, float leaky_slope, int64_t n,
                          const int32_t *n_dev, const float *w, int64_t w_ld, const void *off, int idx_bits,
                          int64_t off_ld, int row_shift, float *S, int64_t ldS, float *wsum, void *stream);

/* E-Net's pointwise stem as a kernel of its own (reference nets/enet.py:24-28,111, nets/net_utils.py:35-43): writes the
 * (n, c3) POINT-MAJOR feature rows (row stride out_ld) that efgh_bcl_splat_gather reads at level 0; weights packed as
 * for efgh_bcl_scatter_stem. */
int efgh_bcl_stem_rows(const float *pts, int64_t pts_ld, int c_in, int c1, int c2, int c3, const float *stem_weights,
                       float leaky_slope, int64_t n, const int32_t *n_dev, float *out, int64_t out_ld, void *stream);

/* Gather-form splat + density normalisation in one pass (reference nets/bilateralNN.py:176-211, with E-Net's input
 * wiring `cat(el_minus_gr, previous features)`, reference nets/enet.py:113-137): for every vertex h
 *   S[h+1, :] = (sum over its contributions (point i, remainder r) of bary[r,i] * [el_minus_gr[:, i] ; feat2[i, :]]) * inv,
 *   inv = 1 / (sum of bary[r,i] + 1e-5) when `normalize`, else 1;  S[0, :] = 0 (the "-1 neighbour" sink row).
 * Same result as efgh_bcl_zero + efgh_bcl_scatter + efgh_bcl_normalize up to the order of the float additions, with
 * no atomics.  point_rows / vertex_offsets / contributions come from efgh_lattice_points_batch /
 * efgh_lattice_vertices_batch of the same level.
 *   feat2: point-major, C2 (% 4 == 0, <= 512) contiguous floats per point, point stride stride_n2; S has 4 + C2 columns;
 *   rows: vertex CAPACITY - the h_cap given to the lattice calls (it locates the heavy-vertex list behind
 *     vertex_offsets); rows_dev the device-side vertex count; inv_out (rows+1) optional. */
int efgh_bcl_splat_gather(const float *point_rows, const float *feat2, int64_t stride_n2, int C2,
                          const int32_t *vertex_offsets, const int32_t *contributions, int64_t rows,
                          const int32_t *rows_dev, int normalize, float *S, int64_t ldS, float *inv_out, void *stream);

/* One launch zero-fills what a BCL forward accumulates into.  With R = rows_dev ? min(*rows_dev, rows) : rows:
 *   S (R + rows_extra rows x C, leading dimension ldS) and wsum (R + rows_extra floats) - the splat targets,
 *   Y2 (R rows x C2, leading dimension ldY2) - the split-K accumulator of efgh_bcl_conv_tc.
 * Any of the three may be NULL.  Cost is proportional to the vertices a scan actually produced, not to the
 * capacity of the buffers. */
int efgh_bcl_zero(float *S, int64_t ldS, int C, float *wsum, float *Y2, int64_t ldY2, int C2, int64_t rows,
                  const int32_t *rows_dev, int rows_extra, void *stream);

/* Splat (bilateralNN.py:176-191) and its density normaliser (:193-211); also the adjoint of slice.
 *   S[(off[r,n]+row_shift), c] += w[r,n] * X[c,n];  wsum[(off[r,n]+row_shift)] += w[r,n] (if wsum)
 *   X = [feat ; feat2] concatenated along channels on the fly (feat2 may be NULL): E-Net feeds every BCL
 *   torch.cat((el_minus_gr, previous output)) (reference nets/enet.py:113-137), which is never materialised.
 *   S is (rows, C + C2) with leading dimension ldS; the caller zero-fills S and wsum beforehand. */
int efgh_bcl_scatter(const float *feat, int64_t stride_c, int64_t stride_n, int C, const float *feat2,
                     int64_t stride_c2, int64_t stride_n2, int C2, int64_t n, const int32_t *n_dev, const float *w,
                     int64_t w_ld, const void *off, int idx_bits, int64_t off_ld, int row_shift, float *S, int64_t ldS,
                     float *wsum, void *stream);

/* inv[r] = 1 / (wsum[r] + 1e-5)   (bilateralNN.py:210) for r < rows (rows_dev + rows_extra if given) */
int efgh_bcl_inv_norm(const float *wsum, float *inv, int64_t rows, const int32_t *rows_dev, int rows_extra,
                      void *stream);

/* Density normalisation in place (bilateralNN.py:210-211): S[r, :] *= 1 / (wsum[r] + 1e-5); the factor is also
 * written to inv_out[r] when inv_out != NULL (the backward pass needs it). */
int efgh_bcl_normalize(float *S, int64_t ldS, int C, const float *wsum, float *inv_out, int64_t rows,
                       const int32_t *rows_dev, int rows_extra, void *stream);

/* Slice (bilateralNN.py:251-261) and the adjoint of splat:
 *   out[c,n] = sum_r w[r,n] * Z[off[r,n]+row_shift, c] * (row_scale ? row_scale[row] : 1) + (bias ? bias[c] : 0) */
int efgh_bcl_gather(const float *Z, int64_t ldZ, int C, const float *row_scale, int64_t n, const int32_t *n_dev,
                    const float *w, int64_t w_ld, const void *off, int idx_bits, int64_t off_ld, int row_shift,
                    const float *bias, float *out, int64_t stride_c, int64_t stride_n, void *stream);

/* Lattice convolution (bilateralNN.py:240-244): gather the F neighbours of each vertex and contract.
 *   Y[h, m] = act( bias[m] + sum_{f,c} Wt[(f*C + c), m] * X[nbr[f,h]+1, c] * (row_scale ? row_scale[row] : 1) )
 *   X (rows, C) vertex-major with leading dimension ldX, row 0 = sink; nbr (F,h) rows with stride nbr_ld.
 *   nbr == NULL means F = 1 and row = h (a 1x1 convolution over a matrix WITHOUT sink row).
 *   Wt is the reference weight (M, C, F, 1) re-laid as (F*C, M) row-major (done once per weight update
 *   by the host module).  act: 0 none, 1 ReLU, 2 LeakyReLU(0.1).  Y (h, M) leading dimension ldY.
 *   precision: 0 = fp32 FFMA (CUDA cores); other values are reserved for the tensor-core paths. */
int efgh_bcl_conv(const float *X, int64_t ldX, int C, const float *row_scale, const void *nbr, int idx_bits,
                  int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev, const float *Wt, const float *bias,
                  int M, int act, float *Y, int64_t ldY, int precision, void *stream);

/* Tensor-core variant of efgh_bcl_conv (tcgen05 / TMEM, kind::tf32, fp32 accumulate).
 *   nsplit = 3: 3xTF32 operand splitting, fp32-equivalent accuracy (parity tests hold 1e-5);
 *   nsplit = 1: one TF32 pass (what cuDNN does for the reference under torch's default allow_tf32).
 * Weights are passed PACKED: efgh_bcl_pack_weights turns the (F*C, M) row-major matrix that efgh_bcl_conv
 * takes into the swizzled shared-memory image the kernel streams in with bulk copies (pack once per weight
 * update; efgh_bcl_packed_weight_bytes gives the buffer size, 16-byte aligned).
 * efgh_bcl_conv_tc_supported: 1 if the shape can run here (C % 4 == 0, M in {32,64,...,256}), else use
 * efgh_bcl_conv.  X, Y 16-byte aligned, ldX and ldY multiples of 4.  Unlike efgh_bcl_conv there is no
 * row_scale argument: normalise the splat matrix first (efgh_bcl_normalize) - rows are copied from L2 into
 * the tensor core's shared-memory tiles asynchronously, without passing through registers.
 *
 * Long contractions are cut into chains of 256 terms (tensor-core accumulation rounds toward zero, so a chain's error
 * grows with its length) whose partial sums are added with round-to-nearest adds:
 *   M <= 128: on chip (shared-memory running sum in the epilogue); efgh_bcl_conv_tc_groups(K, M) == 1 and
 *             accumulate = 0 writes Y = act(bias + sum) once;
 *   M  > 128: efgh_bcl_conv_tc_groups(K, M) K groups added in L2 (also the split-K that fills the SMs on small
 *             lattices): call with accumulate = 1 on a zero-filled Y - the kernel then adds raw partial sums and the
 *             caller applies bias / activation afterwards (efgh_bcl_bias_act, or `in_bias` / `in_act` of the next
 *             efgh_bcl_conv_tc, which applies act(x + in_bias[c]) to X while loading it).
 * accumulate = 0 is only valid when efgh_bcl_conv_tc_groups(F*C, M) == 1. */
int efgh_bcl_conv_tc_supported(int C, int F, int M, int nsplit);
int efgh_bcl_conv_tc_groups(int K, int M);
size_t efgh_bcl_packed_weight_bytes(int K, int M, int nsplit);
int efgh_bcl_pack_weights(const float *Wt, int K, int M, int nsplit, float *Wimg, void *stream);
int efgh_bcl_conv_tc(const float *X, int64_t ldX, int C, const float *in_bias, int in_act,
                     const void *nbr, int idx_bits, int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev,
                     const float *Wimg, const float *bias, int M, int act, float *Y, int64_t ldY, int nsplit,
                     int accumulate, void *stream);
/* Y[h, m] = act(Y[h, m] + bias[m]) in place, h < (h_dev ? *h_dev : h). */
int efgh_bcl_bias_act(float *Y, int64_t ldY, int M, int64_t h, const int32_t *h_dev, const float *bias, int act,
                      void *stream);

/* Backward of efgh_bcl_conv w.r.t. its input: dX[nbr[f,h]+1, c] += sum_m dY[h,m] * Wt[(f*C+c), m]
 * (absent neighbours are skipped; nbr == NULL: dX[h, c] = ..., plain store).  If act_out != NULL the
 * incoming gradient is first masked by the forward activation: dY *= act'(act_out) where act_out is
 * the forward OUTPUT (ReLU: >0 ? 1 : 0, Leaky: >0 ? 1 : 0.1).  The caller zero-fills dX when nbr != NULL. */
int efgh_bcl_conv_dgrad(const float *dY, int64_t ldY, const float *act_out, int64_t ldA, int act, int M,
                        const void *nbr, int idx_bits, int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev,
                        const float *Wt, int C, float *dX, int64_t ldX, void *stream);

/* Backward of efgh_bcl_conv w.r.t. weights and bias (both accumulate into zero-filled outputs):
 *   dWt[(f*C+c), m] += sum_h X[nbr[f,h]+1, c] * row_scale[row] * dYm[h, m];  dbias[m] += sum_h dYm[h,m]
 *   with dYm = dY masked by the activation as above. */
int efgh_bcl_conv_wgrad(const float *X, int64_t ldX, int C, const float *row_scale, const void *nbr, int idx_bits,
                        int64_t nbr_ld, int F, int64_t h, const int32_t *h_dev, const float *dY, int64_t ldY,
                        const float *act_out, int64_t ldA, int act, int M, float *dWt, float *dbias, void *stream);

/* Tensor-core variant of efgh_bcl_conv_wgrad (tcgen05 / TMEM, 3xTF32, both operands MN-major in shared memory):
 *   dWt[(f*C+c), m] += sum_h X[nbr[f,h]+1, c] * dY[h, m];   dbias[m] += sum_h dY[h, m]
 * dY must already carry the activation mask (efgh_bcl_act_bwd); there is no row_scale.  dWt (and dbias) are accumulated
 * into: zero-fill first.  efgh_bcl_conv_wgrad_tc_supported: C % 4 == 0 and M in {32, 64, 96, 128, 256}; X, dY, dWt
 * 16-byte aligned, ldX and ldY multiples of 4. */
int efgh_bcl_conv_wgrad_tc_supported(int C, int F, int M);
int efgh_bcl_conv_wgrad_tc(const float *X, int64_t ldX, int C, const void *nbr, int idx_bits, int64_t nbr_ld, int F,
                           int64_t h, const int32_t *h_dev, const float *dY, int64_t ldY, int M, float *dWt, float *dbias,
                           void *stream);

/* ------------------------------------------------------------------------------------------------
 * Training helpers of the batched pipeline (reference iterater.py:35-43: forward -> loss -> backward).
 *
 * efgh_bcl_act_bwd: dX[h, :] *= act'(act_out[h, :]) in place over rows [0, rows) (rows_dev: device-side count) -
 *   the backward of the ReLU that reference nets/bilateralNN.py:111-115 puts between the two convolutions
 *   (autograd's threshold_backward); act as in efgh_bcl_conv (1 ReLU, 2 LeakyReLU 0.1).
 * efgh_bcl_loss_half_mean_square: loss = (1/B) sum_b 0.5 * mean(Z_b^2) over the B scans of a batch, Z_b = rows
 *   [scan_start[b], scan_start[b+1]) of the vertex-major (rows, C) matrix; writes the loss (one float) and its
 *   gradient dZ in one pass.  A stand-in objective for benchmarks and tests: the reference's losses
 *   (losses/efghloss.py) act on the E-Net head, which is outside the lattice path. */
int efgh_bcl_act_bwd(float *dX, int64_t ldX, const float *act_out, int64_t ldA, int C, int act, int64_t rows,
                     const int32_t *rows_dev, void *stream);
int efgh_bcl_loss_half_mean_square(const float *Z, int64_t ldZ, int C, const int32_t *scan_start, int B, float *dZ,
                                   int64_t ldD, float *loss, int64_t rows_cap, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Point -> image scatter projections and cloud pre-processing (row f4).
 *
 * efgh_project_range_image   replaces reference common/torch_utils.py:11-59 (range_img_from_cartesian_pc_torch):
 *   pc: (batch, 3, n) float32 (row stride pc_ld, batch stride in elements); img: (batch, 4, height, width) = x, y, z,
 *   range at pixel (u, v) = (long)((fov_up - asin(z/r)) / (fov_up - fov_down) * (height-1)),
 *   (long)((-atan2(y, x) + pi) / (2 pi) * (width-1)) for points with fov_down < pitch < fov_up, 0 elsewhere.
 *   fov_up / fov_down in radians (the reference's lidar_fov_rad[i] * pi, as Python floats).
 * efgh_project_depth_image   replaces reference common/torch_utils.py:61-103 (depth_img_from_cartesian_pc_torch):
 *   [x' y' w] = cam_T_velo (batch, 3, 4) @ [pc; 1]; pixel ((long)(y'/w), (long)(x'/w)) inside the image and w > 0
 *   receives (x, y, z, w).
 * Duplicate pixels: the point with the LARGEST index wins (the sequential semantics of the reference's
 * `img[u.tolist(), v.tolist()] = values`); `winner` is a (batch, height, width) int32 scratch that ends up holding
 * that point index (-1 = empty pixel).
 *
 * efgh_preproc_cloud         replaces reference data_loader/loader_utils.py:163-202 (preproc_pcd, without its
 *   reduce_lidar_line branch): xyzi = the (n, 4) float32 records of a `.bin` scan (loader_utils.py:59-61) in device
 *   memory; crop to -radius <= x, y < radius keeping the order (use_radius = 0: radius None); if more than num_points
 *   survive, take points sample[0..num_points) of the CROPPED cloud (the caller draws `sample` - numpy's RNG stays on
 *   the host: np.random.choice(m, num_points, replace=False)), else zero-pad; then the 4x4 float64 rigid transform.
 *   out64: (4, num_points) float64 = the reference's return value; out32: (3, num_points) float32 = what the network
 *   consumes after `.float()` (either may be NULL).  kept_count (device int32) receives the cropped size m - the
 *   caller needs it to size `sample` (two-phase use: call once with num_points >= n to learn m, or over-provision
 *   sample).  A sample index outside [0, m) sets bit 0 of the int32 that follows the workspace's index arrays. */
int efgh_project_range_image(const float *pc, int64_t pc_ld, int64_t batch_stride, int64_t n, int batch, int height,
                             int width, double fov_up, double fov_down, int32_t *winner, float *img, void *stream);
int efgh_project_depth_image(const float *pc, int64_t pc_ld, int64_t batch_stride, int64_t n, int batch,
                             const float *cam_T_velo, int height, int width, int32_t *winner, float *img, void *stream);
size_t efgh_preproc_workspace_bytes(int64_t n);
int efgh_preproc_cloud(const float *xyzi, int64_t n, int use_radius, float radius, const int64_t *sample,
                       int64_t n_sample, int64_t num_points, const double *transform, double *out64, float *out32,
                       int32_t *kept_count, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EFGH_B200_H */
