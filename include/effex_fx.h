/* effex_fx.h -- C ABI of the B200-native FX-correlator hot path (libeffex_fx.so).
 *
 * Drop-in boundary for the DSP region of evanmayer/effex.  The reference has
 * no FFI layer of its own: its hot path is five private methods of
 * `Correlator` that call cupy/cuSignal (effex/effex.py:476-627).  Each entry
 * point below names the reference lines it replaces; `INTEGRATION.md` shows
 * the ctypes stub a maintainer would add to effex.py.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no C++/torch types.
 *   - every call returns 0 (FX_OK) or a negative fx_status; the message for
 *     the last failure on a handle is fx_last_error(h) (NULL handle: the last
 *     failure of fx_create on this thread).
 *   - pointers named d_* are DEVICE pointers on the handle's device and stay
 *     owned by the caller; h_* are HOST pointers.  The handle owns taps,
 *     twiddles, workspaces and its stream.
 *   - calls are asynchronous on the handle's stream unless stated; fx_sync()
 *     waits.  A handle is not re-entrant (the reference calls this path from
 *     one thread, effex/effex.py:399-410).
 *   - there is NO CPU fallback: without a CUDA device fx_create fails.
 *
 * Data layout
 *   raw IQ    uint8, interleaved I,Q, one array per channel (RTL-SDR format,
 *             what pyrtlsdr's packed_bytes_to_iq consumes: x=(b/127.5-1)).
 *             A "block" is num_samp complex samples (2*num_samp bytes); blocks
 *             of one call are contiguous.
 *   spectra   complex64 as float pairs (re,im).  Cross-spectra rows are in
 *             the order effex writes them (fftshifted, effex.py:521).
 */
#ifndef EFFEX_FX_H
#define EFFEX_FX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FX_ABI_VERSION 2

typedef enum {
    FX_OK = 0,
    FX_ERR_INVALID = -1,      /* bad argument (Python wrapper raises ValueError)      */
    FX_ERR_CUDA = -2,         /* CUDA runtime failure                                  */
    FX_ERR_UNSUPPORTED = -3,  /* shape outside the supported set (e.g. ntaps > 32)     */
    FX_ERR_STATE = -4,        /* call order (e.g. fx_process before fx_set_taps)       */
    FX_ERR_COMM = -5          /* cross-GPU reduce: no peer access, or a rank never arrived */
} fx_status;

typedef struct fx_handle fx_handle;

typedef struct {
    int32_t device;       /* CUDA device ordinal                                           */
    int32_t ntaps;        /* T, PFB taps per branch (effex.py:115 uses 4; cuSignal cap 32) */
    int32_t nbins;        /* N = --resolution, 8..65536; any integer (effex.py:734 takes any: cuFFT);
                             powers of two run the fused kernels, others the unfused ones + Bluestein */
    int32_t dc_remove;    /* 1: per-block DC removal of effex.py:394-395; 0: none          */
    int64_t num_samp;     /* S complex samples per block (effex.py:87, --num_samp)         */
    int32_t max_blocks;   /* capacity: blocks per fx_process / fx_integrate call           */
    int32_t flags;        /* FX_FLAG_* bit mask                                            */
} fx_config;

#define FX_FLAG_FORCE_GENERIC 1   /* never take the fused kernels (used to cross-check them) */
#define FX_FLAG_LOCKSTEP_KERNEL 2 /* fused path: use the simpler lock-step kernel instead of the staggered one */
#define FX_FLAG_CROSS_ONLY 4      /* accumulators (fx_integrate*, fx_process_acc, fx_process_reduce) carry the cross-spectrum
                                     and the frame count only -- what the reference outputs (effex.py:520-521); the auto-power
                                     parts d_acc_a0/a1 are left untouched and the kernels skip |F|^2 (4 % fewer FP32 ops)       */

/* ---- lifetime --------------------------------------------------------- */
int fx_abi_version(void);
int fx_device_count(void);
/* Replaces Correlator.__init__'s GPU set-up (effex.py:109-127). */
int fx_create(const fx_config *cfg, fx_handle **out);
int fx_destroy(fx_handle *h);
const char *fx_last_error(const fx_handle *h);
int fx_sync(fx_handle *h);
/* 1 if fx_process on this handle runs the fused unpack->PFB->FFT->X kernel: ntaps == 4, nbins a power
 * of two in [256, 4096], num_samp a multiple of 8 and 16-byte aligned inputs.  ntaps == 4 with nbins in
 * [8192, 65536] runs two kernels around one intermediate (fx_bigfft.cuh; returns 0 here).  Every other
 * shape (ntaps up to 32, nbins 8..128, ragged num_samp, unaligned pointers) runs the unfused kernels. */
int fx_uses_fused(const fx_handle *h);

/* ---- parameters --------------------------------------------------------
 * fx_set_taps: prototype filter h[T*N] in the reference's order, float64, as
 * built on the host by `get_window("hamming",T*N)*firwin(T*N,1/N,'rectangular')`
 * (effex.py:126-127).  Synchronous.                                         */
int fx_set_taps(fx_handle *h, const double *h_taps, size_t n);
/* fx_set_rot: rot[c], c in natural FFT order, as (re,im) float64 pairs,
 * = exp(+2j*pi*(fftfreq(N,1/bw)[c]+fc)*tau) (effex.py:516-519); built in
 * float64 on the host because the phase spans ~1e4 cycles.  NULL = all ones. */
int fx_set_rot(fx_handle *h, const double *h_rot_re_im, size_t nbins);

/* ---- the hot path ------------------------------------------------------
 * fx_process: for each of n_blocks block pairs, what one RUN trip of the loop
 * does (effex.py:391-395 DC removal, :490-527 _run_task/_pfb_xcorr):
 *   d_xspec[b][j]  (complex64, [n_blocks][N])  = fftshift(mean_i F0*conj(F1*rot))
 *   d_auto0/1[b][j] (float32, [n_blocks][N], may be NULL) = fftshift(mean_i |Fk|^2)
 * Input: d_iq0/d_iq1 uint8[n_blocks][2*num_samp].                           */
int fx_process(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
               float *d_xspec, float *d_auto0, float *d_auto1);

/* fx_integrate: same per-block arithmetic, but ADDS the un-normalised sums
 * over all frames of all n_blocks blocks into float64 accumulators in
 * natural bin order, without rot:
 *   d_acc_x[c] (re,im) += sum F0*conj(F1); d_acc_a0[c] += sum|F0|^2; d_acc_a1 likewise
 *   *d_frames += number of frames added.
 * These are the small per-integration accumulators that are reduced across
 * GPUs (one NCCL reduce) before the host applies 1/frames, rot and fftshift. */
int fx_integrate(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                 double *d_acc_x, double *d_acc_a0, double *d_acc_a1, double *d_frames);

/* fx_process_acc: fx_process and fx_integrate in one pass over the input (rows
 * AND accumulators from the same kernel run; d_xspec may be NULL).            */
int fx_process_acc(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                   float *d_xspec, float *d_auto0, float *d_auto1,
                   double *d_acc_x, double *d_acc_a0, double *d_acc_a1, double *d_frames);

/* ---- streaming-history variant (north_star: time shards with PFB halos) ---------------------------
 * The reference restarts the PFB with zero history at every block; the streaming variant
 * treats the n_blocks blocks of the call as ONE contiguous span of a longer recording: PFB history carries
 * across the block boundaries, d_halo0/1 hold the raw bytes of the (ntaps-1)*nbins complex samples that
 * precede the span (NULL = start of the recording: zero history), and the DC mean is h_sums/total_samp when
 * h_sums != NULL (recording-wide byte sums {I0,Q0,I1,Q1}, e.g. all-reduced over ranks), else the span's own.
 * It equals the reference's arithmetic applied to the whole recording as one giant block.  Adds the
 * un-normalised sums of the span's frames into the float64 accumulators like fx_integrate.
 * fx_span_sums returns the exact byte sums of the span (synchronous).                                  */
int fx_span_sums(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks, uint64_t h_sums[4]);
int fx_integrate_stream(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                        const uint8_t *d_halo0, const uint8_t *d_halo1, const uint64_t *h_sums, int64_t total_samp,
                        double *d_acc_x, double *d_acc_a0, double *d_acc_a1, double *d_frames);

/* ---- multi-GPU: the one collective of the path (SURVEY 8(b) `fx_reduce`, 8(e)) ---------------------------
 * The reference is single-GPU; north_star shards a recording by contiguous time blocks over the GPUs of a
 * box and combines the small per-integration accumulators (effex.py:520-521's frame mean, taken over all
 * ranks' frames) with ONE reduce.  The library owns that reduce: every handle has a MAILBOX in its HBM that
 * its peers map over NVLink (cudaIpc), a rank's contribution is written straight into the root's mailbox by
 * the kernel that folds the call's partial sums, and the root adds the world contributions in rank order
 * (deterministic float64) one collective call later, when they have long landed -- no library collective, no
 * extra launch on non-root ranks, no rank ever waiting for another inside a step.
 *
 *   fx_comm_export   every rank: create this handle's mailbox for `world` ranks with slots of at least
 *                    slot_bytes (0 = the accumulators only; the lag search needs 8*M bytes, M = 2^ceil(log2 2n))
 *                    and write its FX_COMM_TOKEN_BYTES token to h_token.  Synchronous.
 *   (the host exchanges tokens with whatever it has: torch.distributed, MPI, a file)
 *   fx_comm_attach   every rank: h_tokens = the world tokens in rank order.  Maps the peers' mailboxes.
 *   fx_process_reduce / fx_integrate_stream_reduce
 *                    fx_process_acc / fx_integrate_stream whose sums are ADDED to the root's accumulators
 *                    d_acc_flat = double[4N+1] = [acc_x (2N) | acc_a0 (N) | acc_a1 (N) | frames (1)]
 *                    (ignored on the other ranks).  COLLECTIVE: every rank of the world makes the same
 *                    sequence of reduce calls with the same root.  Asynchronous; the root folds an epoch at
 *                    its NEXT collective call, so its accumulators are complete after fx_sync(), or for later
 *                    work on fx_stream() after fx_comm_fence() (both issue the outstanding fold).
 *   fx_reduce_f64/f32  in-place sum of d_buf[n] over the ranks into the root's d_buf (collective; the lag
 *                    search reduces its accumulated 2n-point cross-spectrum with it); folded at once on
 *                    fx_stream().  (One host thread driving several ranks must call the root last.)
 * A rank that waits ~30 s for a peer gives up and the next fx_sync() returns FX_ERR_COMM. All ranks must
 * fx_sync() (and the host must barrier) before any of them destroys its handle.                           */
#define FX_COMM_TOKEN_BYTES 128
int fx_comm_export(fx_handle *h, int world, size_t slot_bytes, void *h_token);
int fx_comm_attach(fx_handle *h, int rank, int world, const void *h_tokens);
int fx_comm_fence(fx_handle *h);
int fx_process_reduce(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                      float *d_xspec, float *d_auto0, float *d_auto1, int root, double *d_acc_flat);
int fx_integrate_stream_reduce(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                               const uint8_t *d_halo0, const uint8_t *d_halo1, const uint64_t *h_sums,
                               int64_t total_samp, int root, double *d_acc_flat);
int fx_reduce_f64(fx_handle *h, double *d_buf, size_t n, int root);
int fx_reduce_f32(fx_handle *h, float *d_buf, size_t n, int root);

/* fx_process_host: fx_process with HOST buffers (pinned or pageable): H2D of
 * the raw bytes, compute, D2H of the rows, pipelined in chunks on two
 * streams.  Synchronous.  This is the call the reference-facing wrapper
 * makes per batch of dequeued blocks (effex.py:391-410 + :693 asnumpy).     */
int fx_process_host(fx_handle *h, const uint8_t *h_iq0, const uint8_t *h_iq1, int64_t n_blocks,
                    float *h_xspec, float *h_auto0, float *h_auto1);

/* fx_copy_probe: the copies of fx_process_host and nothing else (same chunks, streams and staging buffers,
 * no kernels; h_xspec receives stale staging contents).  bench.py times it beside fx_process_host: it is the
 * PCIe/host-memory roof of the host-buffer path on the box at hand.                                         */
int fx_copy_probe(fx_handle *h, const uint8_t *h_iq0, const uint8_t *h_iq1, int64_t n_blocks, float *h_xspec);

/* ---- pieces exposed for the reference's own tests -----------------------
 * fx_pfb_c64: Correlator._spectrometer_poly(x, ntaps, n_branches, window)
 * (effex.py:530-555) on a complex64 device array of num_samp samples:
 * d_frames[i][c], [P][N] complex64, P = num_samp / N, INCLUDING the
 * exp(-2j*pi*c/N) factor so it equals cuSignal's channelize_poly(...).T.     */
int fx_pfb_c64(fx_handle *h, const float *d_x, float *d_frames);
/* fx_pfb_u8: same from raw bytes of one block (unpack + DC removal first).  */
int fx_pfb_u8(fx_handle *h, const uint8_t *d_iq, float *d_frames);

/* ---- delay calibration --------------------------------------------------
 * Correlator._estimate_delay_gaussian (effex.py:583-622): zero-pad to 2n,
 * xc = fftshift(ifft(fft(a)*conj(fft(b)))), imax = argmax|xc| (first max),
 * and the three magnitudes |xc[imax-1]|,|xc[imax]|,|xc[imax+1]|.  The host
 * does the float64 log-parabola (:623-627).  n = num_samp of the handle.
 * With n_blocks > 1 the 2n-point cross-spectrum is accumulated over the
 * blocks before the single inverse FFT (BASELINE config 2).  Synchronous.
 * Integer lag = n - imax.  nbhd[k] = -1 marks an out-of-range neighbour
 * (the reference raises IndexError there, effex.py:619 TODO).
 * The chain is ~10 short launches; a call that repeats the previous call's
 * device buffers and n_blocks (re-calibration on a reused staging area) is
 * captured into a CUDA graph on its second occurrence and replayed from the
 * third on -- the replay reads the buffers' current contents
 * (EFFEX_FX_LAG_GRAPH=0 disables this).                                      */
int fx_lag_c64(fx_handle *h, const float *d_x0, const float *d_x1, int64_t n_blocks,
               int64_t *imax, float nbhd[3]);
int fx_lag_u8(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
              int64_t *imax, float nbhd[3]);

/* The same search in two halves, for accumulations that span calls or GPUs ("reduce the 2n-point
 * accumulated cross-spectrum and IFFT on rank 0"):
 *   fx_lag_fft_len         M = 2^ceil(log2(2*num_samp)), the length of the zero-padded transforms
 *   fx_lag_accumulate_*    d_xacc[M] (complex64, natural order) (=|+=) sum_b FFT(a_b)*conj(FFT(b_b)); first != 0
 *                          overwrites.  Asynchronous on fx_stream.
 *   fx_lag_finish          inverse transform of d_xacc (left intact) + argmax as in fx_lag_*; synchronous.
 *   fx_lag_finish_async    the same with the results left on the device (int64 d_imax[1], float d_nbhd[3]).    */
int64_t fx_lag_fft_len(const fx_handle *h);
int fx_lag_accumulate_u8(fx_handle *h, const uint8_t *d_iq0, const uint8_t *d_iq1, int64_t n_blocks,
                         float *d_xacc, int first);
int fx_lag_accumulate_c64(fx_handle *h, const float *d_x0, const float *d_x1, int64_t n_blocks,
                          float *d_xacc, int first);
int fx_lag_finish(fx_handle *h, const float *d_xacc, int64_t *imax, float nbhd[3]);
int fx_lag_finish_async(fx_handle *h, const float *d_xacc, int64_t *d_imax, float *d_nbhd);

/* ---- CSV rows (host side; replaces np.savetxt in Correlator._write_data, effex.py:687-696) ----
 * Formats n_rows rows of nbins complex64 values exactly as `np.savetxt(fh, [row], delimiter=',')`
 * writes them for complex128 (" (%.18e%+.18ej)" per element, ',' between, '\n' after each row),
 * in parallel over rows.  h_out must hold fx_csv_rows_bound(n_rows, nbins) bytes; *out_len gets the
 * number of bytes written (no terminating NUL).  n_threads < 1 = all hardware threads.            */
size_t fx_csv_rows_bound(int64_t n_rows, int64_t nbins);
int fx_csv_format_rows(const float *h_rows, int64_t n_rows, int64_t nbins, int n_threads,
                       char *h_out, size_t out_cap, size_t *out_len);
/* "%.18e" (plus_sign == 0) or "%+.18e" of one double exactly as printf writes it, from the formatter
 * fx_csv_format_rows uses (exact integer arithmetic, no snprintf); returns the character count, no NUL
 * (h_out must hold 32 bytes).                                                                         */
int fx_csv_format_double(double v, int plus_sign, char *h_out);

/* ---- memory helpers (so a non-torch host can drive the library) --------- */
int fx_dev_alloc(fx_handle *h, size_t bytes, void **d_ptr);
int fx_dev_free(fx_handle *h, void *d_ptr);
int fx_host_alloc_pinned(size_t bytes, void **h_ptr);
int fx_host_free_pinned(void *h_ptr);
int fx_memcpy_h2d(fx_handle *h, void *d_dst, const void *h_src, size_t bytes);
int fx_memcpy_d2h(fx_handle *h, void *h_dst, const void *d_src, size_t bytes);
int fx_memset(fx_handle *h, void *d_ptr, int value, size_t bytes);

/* ---- measurement --------------------------------------------------------
 * Counters since the last fx_reset_counters: kernels this library launched,
 * and device time of the dominant (fused / FFT) kernel measured with CUDA
 * events on the handle's stream when timing is enabled.                     */
int fx_reset_counters(fx_handle *h);
int64_t fx_kernel_launches(const fx_handle *h);
int fx_enable_timing(fx_handle *h, int on);
/* ms spent in, and launches of, the dominant kernel while timing was on. */
int fx_dominant_kernel_time(fx_handle *h, double *ms_total, int64_t *launches);
/* raw cudaStream_t of the handle (so a torch host can wait/record on it).  The byte-sum
 * pre-pass (DC means) of every call runs on a second stream, fx_stream_aux, so that it overlaps the
 * previous call's fused kernel: INPUT buffers must be complete before the call, or be ordered
 * before BOTH streams; outputs are ordered on fx_stream only.                                       */
void *fx_stream(fx_handle *h);
void *fx_stream_aux(fx_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* EFFEX_FX_H */
