/* gnssacq.h — C ABI of libgnssacq.so, the B200 (sm_100a) FFT acquisition engine.
 *
 * The reference (pmonta/GNSS-DSP-tools) is pure Python and has no FFI: its hot path is the
 * local search() function of each acquire-*.py script (acquire-gps-l1.py:18-40 and the
 * variants listed in SURVEY.md §8a), fanned out over PRNs by worker()/mp.Pool
 * (acquire-gps-l1.py:100-108). This header is the boundary that replaces that fan-out:
 * one handle per GPU, one batched call for all replicas x Doppler bins. The ctypes binding a
 * maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success or a negative GNSSACQ_E* code;
 * gnssacq_last_error() gives the message for the calling thread. Pointers are caller-owned
 * host memory unless the name says "device". A handle is bound to one CUDA device and one
 * stream; it is thread-compatible (one thread at a time per handle), with no global state.
 */
#ifndef GNSSACQ_H
#define GNSSACQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gnssacq gnssacq_t;

enum {
  GNSSACQ_OK = 0,
  GNSSACQ_EINVAL = -1,   /* bad argument (sizes, NULLs, unsupported FFT length) */
  GNSSACQ_ECUDA = -2,    /* CUDA runtime error; message carries cudaGetErrorString */
  GNSSACQ_ESTATE = -3,   /* call order: signal / replicas not set */
  GNSSACQ_ENOMEM = -4
};

/* Per-replica result as produced on the device (16 bytes; the multi-GPU path all-gathers
 * arrays of these). dbin indexes the nco_freq list of the call, -1 when no metric was > 0
 * (the reference then returns (0,0,0), acquire-gps-l1.py:25,36-40). */
typedef struct {
  float metric;
  int32_t lag;
  int32_t dbin;
  int32_t pad;
} gnssacq_record_t;

const char* gnssacq_last_error(void);

/* Host buffers and asynchrony. Calls that take host INPUT pointers (gnssacq_set_signal,
 * gnssacq_set_replicas[_i8], gnssacq_set_replicas_from_chips, gnssacq_preprocess, the nco_freq list
 * of the search calls) enqueue their host-to-device copies on the handle's stream and may return
 * before the copy has run (gnssacq_set_signal copies on an internal stream, behind everything the
 * handle's stream holds at that moment, so that the transfer overlaps the replica set-up that usually
 * follows; every call that reads or rewrites the capture, and every call that synchronises, waits
 * for it). Pageable memory is staged by CUDA before the call returns, so ordinary malloc'd / numpy
 * buffers may be reused at once; PINNED (page-locked) buffers are read by DMA
 * later and must stay valid and unmodified until gnssacq_synchronize() or the next call that
 * returns results to the host (gnssacq_search, gnssacq_search_grouped, gnssacq_search_sharded,
 * gnssacq_mix, gnssacq_correlate_bank), all of which synchronise the stream. */

/* Replaces the mp.Pool set-up of acquire-gps-l1.py:105-108: one engine per GPU. */
int gnssacq_create(int device, gnssacq_t** out);
int gnssacq_destroy(gnssacq_t* h);

/* Run all work of this handle on an existing CUDA stream (cudaStream_t passed as void*),
 * e.g. torch.cuda.current_stream().cuda_stream. NULL restores the handle's own stream. */
int gnssacq_set_stream(gnssacq_t* h, void* cuda_stream);

/* The 1024-entry complex128 phase table nco_table of gnsstools/nco.py:3-4, interleaved
 * re,im. Optional: by default the library computes cos/sin(2*pi*k/1024) itself; the Python
 * wrapper passes numpy's table so nco.mix is bit-identical to the reference. */
int gnssacq_set_nco_table(gnssacq_t* h, const double* table_c128);

/* The capture `x` that search(x, ...) receives (acquire-gps-l1.py:18): complex64,
 * interleaved re,im, n_samples samples. Copied to the device. */
int gnssacq_set_signal(gnssacq_t* h, const float* iq_c64, int64_t n_samples);
/* Same, but borrow a device buffer (no copy); it must outlive the searches. */
int gnssacq_set_signal_device(gnssacq_t* h, const void* device_iq_c64, int64_t n_samples);

/* Time-domain replicas, R rows of N float32 (+-1, 0 in a zero-padded half; already
 * BOC-modulated): what the reference feeds to fft.fft() at acquire-gps-l1.py:23-24,
 * acquire-gps-l5i.py:23-24, acquire-galileo-e1b.py:24-26. N is the FFT length = number of
 * lags searched. The replica spectra are computed on the device. */
int gnssacq_set_replicas(gnssacq_t* h, const float* replicas, int32_t R, int32_t N);

/* Same, from R*N float32 already resident on the device (no copy; read before return of
 * the stream work, so keep it alive until the stream has drained). */
int gnssacq_set_replicas_device(gnssacq_t* h, const void* device_replicas, int32_t R, int32_t N);

/* Replicas as int8 (+1, -1, 0 — what code()/boc11()/zero padding produce), host or device: a
 * quarter of the bytes of the float32 form; converted on the device, same result. */
int gnssacq_set_replicas_i8(gnssacq_t* h, const int8_t* replicas, int32_t R, int32_t N);
int gnssacq_set_replicas_i8_device(gnssacq_t* h, const void* device_replicas_i8, int32_t R, int32_t N);

/* The Doppler x block loops of search() for every replica at once.
 *   nco_freq[D]   wipe-off frequency per Doppler bin in cycles/sample, i.e. the float64
 *                 value -doppler/fs (or -(562500*chan+doppler)/fs, acquire-glonass-l1.py:28)
 *                 exactly as the reference computes it
 *   block_stride  samples between the starts of consecutive non-coherent blocks
 *                 (n for both circular and zero-padded variants; block length is N)
 *   n_blocks      non-coherent sums ("ms", "ms//10", ... per script)
 *   normalize     1: metric = q[idx]/mean(q) (acquire-gps-l1.py:35); 0: raw q[idx]
 *   n_lags        argmax over lags [0, n_lags); pass N (or <= 0) for the reference behaviour
 * Outputs, each R long: metric, lag (= idx of acquire-gps-l1.py:34), dbin (index into
 * nco_freq, -1 if nothing exceeded 0). q_dump, if not NULL, receives the full R x D x N
 * float32 grid of q (testing aid). Synchronous. */
int gnssacq_search(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride,
                   int32_t n_blocks, int32_t normalize, int32_t n_lags,
                   float* metric, int32_t* lag, int32_t* dbin, float* q_dump);

/* Same search, asynchronous on the handle's stream, writing R gnssacq_record_t to a device
 * buffer (e.g. a torch tensor that is then all-gathered). nco_freq is read before return. */
int gnssacq_search_device(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride,
                          int32_t n_blocks, int32_t normalize, int32_t n_lags,
                          void* device_records);

/* One call for a list of Doppler GROUPS: nco_freq holds D / group_len consecutive groups of
 * group_len entries and the per-replica best is taken inside each group — the channel loop of the
 * FDMA scripts (acquire-glonass-l1.py:26-39, 60-69: one search() per channel with
 * w = nco(-(562500*chan+doppler)/fs), the replica being channel-independent) as a single batch.
 * Outputs: R * (D / group_len) entries, entry g*R + r for group g and replica r; dbin counts from the
 * start of the group (-1: nothing exceeded 0). */
int gnssacq_search_grouped(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t group_len, int32_t block_stride,
                           int32_t n_blocks, int32_t normalize, int32_t n_lags, float* metric, int32_t* lag, int32_t* dbin);

/* Multi-GPU, one process (and one handle) per GPU — the reference's only parallelism is
 * mp.Pool over PRNs on one host (acquire-gps-l1.py:105-108); here the Doppler list is split into
 * contiguous ascending shards and the per-replica records are exchanged with ONE all-gather.
 * NCCL is bound at run time (dlopen "libnccl.so.2"), the library does not link against it.
 *   gnssacq_nccl_unique_id   128-byte ncclUniqueId; create it on one rank and hand it to the others
 *                            by any means (file, socket, MPI, torch.distributed).
 *   gnssacq_nccl_init        collective over all ranks: joins this handle's device to the communicator.
 *   gnssacq_search_sharded   collective: rank k searches bins [k*D/world ...) of the SAME nco_freq
 *                            list every rank passes, all-gather of R 16-byte records on the handle's
 *                            stream, then a device merge in rank order (strict '>': ties go to the
 *                            lowest bin, as the reference's ascending scan) — every rank returns the
 *                            single-GPU answer, dbin indexing the full list. Without a communicator
 *                            it is gnssacq_search. */
int gnssacq_nccl_unique_id(void* id128);
int gnssacq_nccl_init(gnssacq_t* h, const void* id128, int32_t rank, int32_t world);
int gnssacq_search_sharded(gnssacq_t* h, const double* nco_freq, int32_t D, int32_t block_stride, int32_t n_blocks,
                           int32_t normalize, int32_t n_lags, float* metric, int32_t* lag, int32_t* dbin);

/* nco.mix(x, f, p) of gnsstools/nco.py:30-41 on the GPU, in place on a host complex64
 * buffer (copy in, mix, copy out). Bit-identical to the reference. */
int gnssacq_mix(gnssacq_t* h, float* iq_c64, int64_t n_samples, double f, double p);

/* The capture front end of every acquire-*.py (acquire-gps-l1.py:80-96) on the GPU: raw
 * interleaved int8 I/Q (what io.get_samples_complex reads, gnsstools/io.py:3-12) ->
 * nco.mix(x, mix_f, mix_p) -> scipy.signal.filtfilt(fir, [1], x) (default odd padding of
 * 3*ntaps) -> np.interp at positions step*t, t < n_out. fir is the caller's firwin() result,
 * step is the reference's (1/fsr). The complex64 result becomes the engine's capture (as if
 * passed to gnssacq_set_signal) without leaving the device; if out_c128 is not NULL the
 * complex128 values (equal to the scipy/numpy pipeline to ~1e-15 relative) are also copied back,
 * interleaved re,im, n_out entries. */
int gnssacq_preprocess(gnssacq_t* h, const int8_t* iq_int8, int64_t n_samples, double mix_f, double mix_p,
                       const double* fir, int32_t ntaps, double step, int64_t n_out, double* out_c128);

/* Introspection for tests and the benchmark. */
/* Replica set-up on the device: what `<sig>.code(prn,0,0,incr,n)`
 * (gnsstools/gps/ca.py:106-112), optionally times nco.boc11(0,0,incr,n) (gnsstools/nco.py:12-19),
 * optionally followed by n zeros, hands to fft.fft() in every search()
 * (acquire-gps-l1.py:22-24, acquire-gps-l1cd.py:22-26, acquire-gps-l5i.py:22-24) — built from
 * the 0/1 chip tables instead of being resampled on the host and uploaded sample by sample.
 * chips01: R tables of L chips (values 0/1). Replica r, sample i < n takes chip
 * floor(base + incr*i) mod L with base = (chips % L) + frac formed by the caller in float64
 * (0.0 for the acquisition scripts); boc != 0 multiplies by the BOC(1,1) square wave with
 * base2 = (chips % 2) + frac; samples n..N-1 are zero. Then as gnssacq_set_replicas. */
int gnssacq_set_replicas_from_chips(gnssacq_t* h, const int8_t* chips01, int32_t R, int32_t L, int32_t n, int32_t N,
                                    double base, double incr, int32_t boc, double base2);

/* Time-domain correlator bank for the serial long-code acquisitions:
 * acquire-gps-l2cl.py:18-33, acquire-glonass-l1-p.py:14-32, acquire-glonass-l2-p.py:14-32.
 * For every hypothesis h < H and block b < n_blocks:
 *   out[h][b] = sum_{i<n} x[b*block_stride + i] * nco(nco_freq, 0, n)[i] * (1 - 2*chips01[idx])
 *   idx = floor(base[h*n_blocks + b] + incr*i) mod L,   base = (chips % L) + frac (caller, float64)
 * on the capture given to gnssacq_set_signal. out_c128: H*n_blocks complex128, interleaved
 * re,im (the reference then forms q[h] = sum_b |out[h][b]| and keeps the first maximum). */
int gnssacq_correlate_bank(gnssacq_t* h, const int8_t* chips01, int32_t L, double nco_freq, int32_t n,
                           int32_t n_blocks, int32_t block_stride, const double* base, int32_t H, double incr,
                           double* out_c128);

/* Batched tracking correlators (second half): `<sig>.correlate(x, prn, chips,
 * frac, incr, c[, boc11])` of the tracking scripts (track-gps-l1.py:51-53: early / prompt / late) for
 * H hypotheses (channel x tap) in one call:
 *   out[h] = sum_{i<n} x[xsel[h]*n + i] * (1 - 2*chips01[csel[h]*L + int(cp_i)]) * sub_i
 *   cp_0 = start[h] mod L (start = chips + frac, formed by the caller in float64),
 *   cp_{i+1} = (cp_i + incr[h]) mod L — the reference's float64 recurrence, evaluated as such, so every
 *   sample sees the chip the reference's loop sees.
 * mode 0: sub_i = 1                                      gnsstools/gps/ca.py:120-128
 * mode 1: sub_i = sub[int(bp_i)], bp_0 = (2*start) mod 2, bp += 2*incr (mod 2)   BOC(1,1) / L2C RZ slots
 * mode 2: sub_i = a1*sub[int(bp_i)] + a6*sub[int(bp6_i)], bp6_0 = (12*start) mod 2, bp6 += 12*incr   gnsstools/galileo/e1b.py:45-58
 * mode 3: sub_i = pattern[int(cp_i) mod 33] ? sub[int(bp6_i)] : sub[int(bp_i)]          gnsstools/gps/l1cp.py:210-228
 * params (modes 1-3): {sub[0], sub[1], a1, a6, pattern[0..32]} (37 doubles for mode 3, 4 otherwise).
 * x_c64: nx blocks of n complex64 samples (already carrier-wiped, as track() does with nco.mix);
 * chips01: ncodes tables of L chips; out_c128: H complex128, interleaved re,im. Data-parallel across
 * hypotheses, sequential in time inside each (one CTA per hypothesis). */
int gnssacq_correlate_epl(gnssacq_t* h, const float* x_c64, int32_t nx, int32_t n, const int8_t* chips01, int32_t ncodes, int32_t L,
                          int32_t mode, const double* params, int32_t H, const int32_t* xsel, const int32_t* csel,
                          const double* start, const double* incr, double* out_c128);

int gnssacq_plan_info(gnssacq_t* h, int32_t* N, int32_t* N1, int32_t* N2, int32_t* large);
int64_t gnssacq_launch_count(gnssacq_t* h);   /* kernels launched by this handle so far */
/* Which correlate kernels the current plan runs: bit 0 = specialised rows kernel, bit 1 =
 * specialised columns kernel; 4 = the 16x16x16 single-CTA kernels (N = 4096); 0 = generic
 * runtime-planned kernels; bit 3 / bit 4 = the length-N1 / length-N2 tile transform runs in its
 * twiddle-free prime-factor form (coprime radix schedule); bit 5 = the four-step split itself
 * is coprime (Good-Thomas: no twiddle pass between the two transforms). Negative on error. */
int gnssacq_kernel_variant(gnssacq_t* h);
int gnssacq_synchronize(gnssacq_t* h);
/* Tuning switches, for tests and A/B measurements. "specialized_kernels" (default 1): use the
 * plan-specialised correlate kernels when the FFT length has one; 0 forces the generic
 * runtime-planned kernels. Both produce the same results to rounding.
 * "overlap_chunks" (default 1): large plans alternate their unit chunks over two internal
 * streams so the rows kernel of one chunk overlaps the columns kernel of the other; 0 runs
 * every kernel back to back on the handle's stream (what per-kernel stage times need).
 * "small_ctas" (default 3): bit 0 = 8-row / 128-160-thread rows kernel, bit 1 = 128-thread
 * columns kernel with the radix-31 butterfly in one thread (4-7 CTAs per SM instead of 2-3;
 * bit-identical results); 0 = the 256-thread kernels.
 * "gt_split" (default 1): lengths with a coprime split N = N1*N2 (163680 = 341 x 480,
 * 61380 = 279 x 220) run the four-step in Good-Thomas form, without the twiddle pass; 0 keeps
 * the Cooley-Tukey split. Replicas must be set again afterwards.
 * "v3" (default 1): those lengths run the copy-engine-fed correlate pair (bulk / tensor-map
 * copies + mbarriers); "v3_rows" / "v3_cols" select among its instantiated tile shapes and
 * protocols (0 = the measured best: for 480-point rows the two-role kernel, warp 0 multiply +
 * first stage, warp 1 second stage + bulk store; the others are A/B variants, registry.cu),
 * "v3_rc" / "v3_g" the replicas x Doppler bins of one launch pair, "lanes" the number of
 * internal streams, "fused" (default 0) the single-launch persistent form, "fwd_v6" (default 1) the
 * copy-engine-fed rows pass of the forward transforms where one is instantiated. All bit-identical. */
int gnssacq_set_option(gnssacq_t* h, const char* name, int32_t value);
/* Tuning: force the stage radices (forward order) of the length-N1 (which = 1) or length-N2
 * (which = 2) tile transform; ignored when their product does not match; n = 0 restores the
 * planner's choice. Replicas must be set again afterwards. */
int gnssacq_set_schedule(gnssacq_t* h, int32_t which, const int32_t* radices, int32_t n);
/* Per-stage device time: when profiling is on, CUDA events bracket the launches of each stage
 * on the handle's stream. Stages: 0 wipe-off+forward FFT, 1 correlate rows kernel (large
 * plans only), 2 correlate kernel (mid plans) / correlate columns kernel (large plans),
 * 3 finalize. gnssacq_get_stage_times synchronises, then returns accumulated milliseconds
 * and launch counts (4 entries each) and optionally resets them. */
int gnssacq_set_profiling(gnssacq_t* h, int32_t on);
int gnssacq_get_stage_times(gnssacq_t* h, double* ms4, int64_t* launches4, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* GNSSACQ_H */
