/* maf.h -- C ABI of the B200-native audio front-end (spectral-gating denoise, BS.1770 loudness, Whisper log-mel).
 *
 * The reference (AndersonGACFilho/Modular-Audio-Pipeline) is pure Python and has no FFI for this path: its
 * "plugin interface" is the PreprocessorProtocol / TranscriberProtocol seam (audio_pipeline/protocols.py:68-90,
 * :119-133).  These entry points are what a ctypes binding behind that seam calls; every one cites the reference
 * code it replaces.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions: every function returns 0 on success or a negative MAF_ERR_* code (text via maf_last_error()).
 * All `*_dev` pointers are caller-owned DEVICE pointers (e.g. torch.Tensor.data_ptr()); `stream` is a cudaStream_t
 * passed as void*.  Nothing allocates device memory after maf_plan_create(); the scratch `workspace_dev` is
 * caller-allocated with maf_plan_workspace_bytes() bytes (256-byte aligned).  All calls are asynchronous on `stream`
 * unless noted.  A ctx / plan may be used from one thread at a time.
 */
#ifndef MAF_H_
#define MAF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAF_OK 0
#define MAF_ERR_INVALID (-1)      /* bad argument */
#define MAF_ERR_CUDA (-2)         /* CUDA runtime error */
#define MAF_ERR_UNSUPPORTED (-3)  /* geometry the kernels do not cover */
#define MAF_ERR_TOO_SHORT (-4)    /* audio shorter than one 400 ms loudness block (pyloudnorm raises ValueError) */

typedef struct maf_ctx maf_ctx;    /* per-device constant tables (windows, twiddles, mel banks) */
typedef struct maf_plan maf_plan;  /* geometry of one batch: item / chunk tables, workspace layout */

/* One recording of a batch -- or one time shard of a long recording (whole 600 000-sample denoise chunks). */
typedef struct maf_item_desc {
    int64_t buf_len;   /* samples in the item's local PCM buffer (own range plus supplied halo)               */
    int64_t origin;    /* local index of the first own sample (0 for a whole recording)                      */
    int64_t own_len;   /* samples this item produces output for                                              */
    int64_t g0;        /* index of the first own sample in the full recording (0 for a whole recording)      */
    int64_t n_global;  /* length of the full recording                                                       */
    int32_t lo_avail;  /* denoised halo samples the caller will provide before / after the own range         */
    int32_t hi_avail;  /* (sharded mode only; 0 for a whole recording)                                       */
} maf_item_desc;

typedef struct maf_params {
    int32_t sample_rate;             /* config.audio.sample_rate (config.py:52-57)                            */
    int32_t stationary;              /* 1 = stationary gate (what preprocessor.py:200-206 calls), 0 = non-stationary */
    float prop_decrease;             /* 0.8 (preprocessor.py:204)                                            */
    int32_t auto_detect_noise;       /* NoiseReductionConfig.auto_detect_noise (config.py:76-82)             */
    float noise_sample_duration_s;   /* NoiseReductionConfig.noise_sample_duration_s                          */
    int32_t external_noise;          /* 1 = thresholds come from maf_set_noise_clips() (noise_sample_path)    */
    int32_t peak_normalize;          /* 1 = pydub normalize between denoise and loudness (pipeline.py:346)    */
    float target_lufs;               /* -16.0 (preprocessor.py:244)                                           */
    int32_t n_mels;                  /* 80 / 128; 0 = no log-mel                                              */
    int32_t mel_padding;             /* zero samples appended: 480000 whisper.transcribe, 160 faster-whisper, 0 bare */
} maf_params;

typedef struct maf_layout {           /* per-item offsets into the caller's arrays (filled by maf_plan_layout) */
    int64_t pcm_off;                  /* int16 samples: input buffer of the item                              */
    int64_t out_off;                  /* int16 samples: output (own range) of the item                        */
    int64_t mel_off;                  /* floats: [n_mels, mel_frames] block of the item                       */
    int32_t mel_frames;
    int32_t n_chunks;
} maf_layout;

int maf_version(void);
const char* maf_last_error(void);

int maf_create(int device, maf_ctx** out);
void maf_destroy(maf_ctx* ctx);

/* Geometry of a batch.  Replaces the per-file Python set-up of AudioPreprocessor.__init__ / from_config
 * (preprocessor.py:42-69) plus noisereduce's chunk bookkeeping (get_traces).  Synchronous (uploads small tables). */
int maf_plan_create(maf_ctx* ctx, const maf_params* params, const maf_item_desc* items, int32_t n_items, maf_plan** out);
void maf_plan_destroy(maf_plan* plan);
size_t maf_plan_workspace_bytes(const maf_plan* plan);
int maf_plan_layout(const maf_plan* plan, int32_t item, maf_layout* out);
int64_t maf_plan_total_pcm(const maf_plan* plan);   /* int16 samples the pcm array must hold   */
int64_t maf_plan_total_out(const maf_plan* plan);   /* int16 samples the output array must hold */
int64_t maf_plan_total_mel(const maf_plan* plan);   /* floats the mel array must hold           */
int64_t maf_plan_launches(const maf_plan* plan);    /* kernels launched by this library through the plan so far */

/* Per-stage device timing for bench.py's roofline: after maf_plan_profile(plan, S) every maf_process call records CUDA
 * events between its stages into a ring of S sets; maf_plan_stage_ms reads stage 1..9 (noise+S0, stft, ns_mask,
 * smooth_mask, istft, kweight, gate, logmel, mel_finalize) of ring slot `step` once the stream is synchronised. */
int maf_plan_profile(maf_plan* plan, int32_t max_steps);
int maf_plan_stage_ms(maf_plan* plan, int32_t step, int32_t stage, float* ms);

/* Explicit noise clips (reduce_stationary_noise(input_wav, noise_sample_path), preprocessor.py:174-178):
 * clip i is noise_pcm_dev[offsets[i] .. offsets[i] + lengths[i]).  Host arrays, copied synchronously. */
int maf_set_noise_clips(maf_plan* plan, const int64_t* offsets, const int32_t* lengths);

/* ---- the fused hot path: denoise -> (peak normalise) -> loudness normalise -> log-mel ------------------------
 * pipeline.py:331-347 followed by the transcriber front-end (transcriber.py:229-238).
 * out_pcm_dev : final int16 PCM (what `_loudnorm.wav` holds)            [maf_plan_total_out]
 * mel_dev     : float32 [n_mels, frames] per item                        [maf_plan_total_mel]  (NULL iff n_mels = 0)
 * lufs_dev    : float64 per item, loudness measured BEFORE the gain      (may be NULL)
 * clip_dev    : int32 [n_items][2] detected noise clip (start, end) in samples (may be NULL)
 * noise_pcm_dev: only when params.external_noise                                                              */
int maf_process(maf_plan* plan, const int16_t* pcm_dev, const int16_t* noise_pcm_dev, void* workspace_dev,
                int16_t* out_pcm_dev, float* mel_dev, double* lufs_dev, int32_t* clip_dev, void* stream);

/* Same call with HOST buffers: copies in, runs maf_process, copies out, and returns when the results are in host
 * memory.  Device staging buffers are allocated on first use and cached in the plan. */
int maf_process_host(maf_plan* plan, const int16_t* pcm_host, const int16_t* noise_pcm_host, int64_t noise_len,
                     int16_t* out_pcm_host, float* mel_host, double* lufs_host, int32_t* clip_host);

/* ---- stage entry points (the Protocol methods call these one at a time) ----------------------------------------
 * maf_denoise        AudioPreprocessor.reduce_stationary_noise   preprocessor.py:149-214 (incl. _detect_noise_segments :92-147)
 * maf_peak_normalize AudioPreprocessor.normalize_audio           preprocessor.py:216-242 (16-bit mono same-rate case)
 * maf_loudness       AudioPreprocessor.normalize_loudness        preprocessor.py:244-288
 * maf_logmel         whisper.audio.log_mel_spectrogram           transcriber.py:229-238 / :421-427
 * in/out int16 arrays use the plan layout (pcm_off for inputs, out_off for outputs).                              */
int maf_denoise(maf_plan* plan, const int16_t* pcm_dev, const int16_t* noise_pcm_dev, void* workspace_dev,
                int16_t* out_pcm_dev, float* out_f32_dev /* may be NULL */, int32_t* clip_dev /* may be NULL */, void* stream);
int maf_peak_normalize(maf_plan* plan, const int16_t* pcm_dev, void* workspace_dev, int16_t* out_pcm_dev, void* stream);
int maf_loudness(maf_plan* plan, const int16_t* pcm_dev, void* workspace_dev, int16_t* out_pcm_dev, double* lufs_dev,
                 void* stream);
int maf_logmel(maf_plan* plan, const int16_t* pcm_dev, void* workspace_dev, float* mel_dev, void* stream);

/* ---- pieces for time-sharded recordings (one rank per GPU; the caller runs the NCCL collectives between them) ---
 * maf_frame_stats      per-frame energy (float32) / zero-crossing count (int32) of the own range -> all-gather
 * maf_select_clip      selection over the gathered statistics of ONE recording -> seg_dev[2] (global sample range)
 * maf_sharded_denoise  thresholds from the broadcast clip + K1..K3 -> workspace den16 (own range) and local peak
 * maf_sharded_energy   after the caller all-reduced the peak and exchanged den16 halos: sub-block energies
 * maf_gate             gating over the gathered sub-block energies of ONE recording -> gate_dev (opaque 24 bytes)
 * maf_sharded_logmel   gain + log-mel of the own frames (un-clamped), local max; maf_mel_finalize after all-reduce   */
int maf_frame_stats(maf_plan* plan, const int16_t* pcm_dev, float* energy_dev, int32_t* zc_dev, void* stream);
int maf_select_clip(maf_ctx* ctx, const float* energy_dev, const int32_t* zc_dev, int64_t n_frames, int64_t n_global,
                    int32_t sample_rate, float noise_sample_duration_s, int32_t* seg_dev, void* stream);
int maf_sharded_denoise(maf_plan* plan, const int16_t* pcm_dev, const int16_t* noise_pcm_dev, void* workspace_dev, void* stream);
int maf_sharded_energy(maf_plan* plan, void* workspace_dev, void* stream);
int maf_gate(maf_ctx* ctx, const double* sub_energy_dev, int64_t n_tiles, int64_t n_blocks, int32_t sample_rate,
             float target_lufs, const int32_t* peak_norm_dev, void* gate_dev, void* stream);
int maf_sharded_logmel(maf_plan* plan, void* workspace_dev, const void* gate_dev, int16_t* out_pcm_dev, float* mel_dev, void* stream);
int maf_mel_finalize(maf_plan* plan, void* workspace_dev, float* mel_dev, void* stream);
/* byte offset / size of a named workspace buffer, so the host can view it as a tensor for a collective:
 * "peak" int32[n_items], "peak_norm" int32[n_items], "mel_max" int32[n_items] (ordered-int float),
 * "sub_energy" float64[tiles], "den16" int16 (own range preceded / followed by 2048 halo samples per item),
 * "gates", "T2", "seg" int32[n_items][2] (noise-clip range), "mel_tmin" int32[n_items][max tiles] (ordered-int float: minimum
 * log10 mel of every 32-frame tile -- the tiles below max - 8 are the ones maf_mel_finalize rewrites)             */
int maf_plan_buffer(const maf_plan* plan, const char* name, int64_t* offset_bytes, int64_t* size_bytes);
/* (offset, size) of EVERY buffer carved out of the workspace, in order; returns their number (fills at most `cap` entries).
 * With MAF_DEBUG_REDZONE=<bytes> in the environment at plan creation the buffers are that far apart: the bounds-check run of
 * the parity suite verifies that the gaps are never written and never influence a result.                              */
int maf_plan_ws_layout(const maf_plan* plan, int64_t* offsets_bytes, int64_t* sizes_bytes, int32_t cap);
int64_t maf_plan_den_offset(const maf_plan* plan, int32_t item);   /* int16 index of own sample 0 inside "den16" */

/* Sharded noise clip without a host round trip (seg_dev = the {start, end} maf_select_clip wrote, in global samples):
 * maf_sharded_clip_contrib fills clip_dev[600000] (int16) with the samples of the range this shard owns and zeros
 * elsewhere -- every sample is owned by exactly one rank, so the all-reduce(SUM) of the buffer viewed as 300000 int32
 * words (pairs of samples; no carries) is the clip; maf_set_noise_clips_dev is maf_set_noise_clips for the
 * one item of a sharded plan with the length taken from seg_dev on the device (clip at offset 0 of noise_pcm_dev). */
int maf_sharded_clip_contrib(maf_plan* plan, const int16_t* pcm_dev, const int32_t* seg_dev, int16_t* clip_dev, void* stream);
int maf_set_noise_clips_dev(maf_plan* plan, const int32_t* seg_dev, void* stream);

/* Fused exchanges of the time-sharded mode (one all-gather per barrier instead of one collective per statistic; every buffer
 * is caller-owned device memory, nothing here synchronises, so a whole sharded step can be captured in a CUDA graph):
 *   maf_gather_compact      dst = concatenation over ranks r of src[r * stride + offset : ... + counts_dev[r]], elements of 4 or
 *                           8 bytes -- un-pads an all-gather of per-rank arrays of different lengths (frame statistics)
 *   maf_sharded_pack_edges  out[maf_sharded_edge_bytes()] = first 2048 own denoised samples | last 2048 | local peak
 *   maf_sharded_unpack_edges  gathered[world][edge_bytes] -> global peak into the workspace, the left neighbour's tail / right
 *                           neighbour's head into the 2048-sample halos of the denoised signal (K-weighting warm-up, STFT-400 halo)
 *   maf_sharded_pack_energy out[1 + max_tiles] (float64) = local normalised peak | 100 ms sub-block energies, zero padded
 *   maf_sharded_unpack_energy gathered[world][1 + max_tiles] -> global normalised peak into the workspace, sub_all_dev = the
 *                           energies of the whole recording (counts_dev[r] tiles of rank r) for maf_gate                       */
int maf_gather_compact(maf_ctx* ctx, const void* src_dev, int32_t elem_bytes, int32_t world, int64_t stride, int64_t offset,
                       const int32_t* counts_dev, void* dst_dev, void* stream);
int maf_sharded_edge_bytes(void);
int maf_sharded_pack_edges(maf_plan* plan, void* workspace_dev, void* out_dev, void* stream);
int maf_sharded_unpack_edges(maf_plan* plan, void* workspace_dev, const void* gathered_dev, int32_t world, int32_t rank, void* stream);
int maf_sharded_pack_energy(maf_plan* plan, void* workspace_dev, double* out_dev, int32_t max_tiles, void* stream);
int maf_sharded_unpack_energy(maf_plan* plan, void* workspace_dev, const double* gathered_dev, int32_t world, int32_t max_tiles,
                              const int32_t* counts_dev, double* sub_all_dev, void* stream);

/* ---- normalize_loudness on MULTI-CHANNEL 16-bit input (preprocessor.py:256-261: the reference averages the channels in float32,
 * `samples.reshape(-1, channels).mean(axis=1)`, meters the mean and writes the gained MONO signal, :271-282).  pcm_dev holds
 * n_frames interleaved frames of `channels` int16 samples; out_dev receives n_frames mono samples -- untouched when the
 * recording is too quiet (*lufs_dev non-finite or < -70: the reference returns its input path, :267-269).
 * workspace_dev: maf_loudness_interleaved_ws() bytes. */
int64_t maf_loudness_interleaved_ws(int64_t n_frames, int32_t sample_rate);
int maf_loudness_interleaved(maf_ctx* ctx, const int16_t* pcm_dev, int64_t n_frames, int32_t channels, int32_t sample_rate,
                             float target_lufs, void* workspace_dev, int16_t* out_dev, double* lufs_dev, void* stream);

/* ---- silence detection / removal for AudioPreprocessor.remove_silence and detect_silence_segments (preprocessor.py:290-405 ->
 * pydub.silence.detect_silence / detect_nonsilent over audioop.rms, then the margin / cross-fade / position loop :330-358 over
 * AudioSegment.__getitem__ / append).
 * maf_ms_energy: exact integer sum of squares per millisecond bucket [int(i r), int((i+1) r)), r = rate / 1000, and of the
 * whole recording (seg.rms -> seg.dBFS; the host turns that ONE scalar into the threshold with the reference's float64
 * expression, `db_to_float(seg.dBFS - offset) * 32768`).
 * maf_silence: everything else on the device -- window RMS from prefix sums, pydub's merge of silent starts into ranges,
 * detect_nonsilent's complement, the reference's margin / min(20, dur // 4) ms cross-fade / running-position bookkeeping,
 * and (out_dev != NULL) the compaction of the kept chunks with pydub's cross-fade into out_dev (at most out_cap frames).
 * The workspace starts with the result block the host copies back (maf_silence_layout: result_bytes):
 *   int32 status (0 ok, 1 no non-silent segment -> the reference returns its input, 2 pydub ValueError "crossfade is longer than
 *   the original AudioSegment", 3 pydub TooManyMissingFrames), int32 n_silent, int32 n_chunks, int32 j_base, int64 n_out, int64 n_ms,
 *   then at off_silent int32[max_ranges][2] silent ranges in ms, at off_map int32[max_ranges + 1][4] per kept chunk
 *   (processed_start_ms, duration_ms, original_start_ms, original_end_ms) -> TimestampMapping (protocols.py:35-41). */
int maf_ms_energy(maf_ctx* ctx, const int16_t* pcm_dev, int64_t n, int32_t sample_rate, int64_t n_ms, uint64_t* bucket_dev,
                  uint64_t* total_dev, void* stream);
int maf_silence_layout(int64_t n_ms, int32_t window_ms, int64_t* ws_bytes, int64_t* result_bytes, int64_t* max_ranges,
                       int64_t* off_silent, int64_t* off_map);
int maf_silence(maf_ctx* ctx, const int16_t* pcm_dev, int64_t n, int32_t sample_rate, const uint64_t* bucket_dev, int64_t n_ms,
                int32_t window_ms, double thresh, int32_t margin_ms, void* workspace_dev, int16_t* out_dev, int64_t out_cap, void* stream);

/* ---- front door of AudioPreprocessor.normalize_audio for 16-bit input that is not yet mono at the target rate
 * (preprocessor.py:227-233: pydub normalize(seg).set_frame_rate(sr).set_channels(1) -> audioop.max / mul / ratecv /
 * tomono, in this order).  pcm_dev holds n_frames_in interleaved frames of `channels` (1 or 2) int16 samples;
 * out_dev receives maf_front_door_out_len() mono samples at out_rate.  peak_normalize = 0 skips the normalise. */
int64_t maf_front_door_out_len(int64_t n_frames_in, int32_t in_rate, int32_t out_rate);
int maf_front_door(maf_ctx* ctx, const int16_t* pcm_dev, int64_t n_frames_in, int32_t channels, int32_t in_rate, int32_t out_rate,
                   int32_t peak_normalize, int16_t* out_dev, void* stream);

/* ---- 30 s encoder windows, the hand-off into a Whisper encoder on the same GPU (transcriber.py:229-238 ->
 * whisper/transcribe.py seek loop: mel_segment = mel[:, seek : seek + segment_size]; pad_or_trim(mel_segment, N_FRAMES)
 * .to(dtype); faster-whisper slices features[:, seek : seek + nb_max_frames] the same way).  mel_dev is one recording's
 * [n_mels, frames] float32 log-mel; out_dev receives [n_windows, n_mels, window_frames] float32 (out_fp16 = 0) or IEEE
 * half, round to nearest even (out_fp16 = 1): out[w, m, f] = mel[m, seeks[w] + f] for f < lens[w] (lens_dev NULL:
 * window_frames) and seeks[w] + f < frames, else 0.  seeks_dev / lens_dev are int32 device arrays of n_windows entries. */
int maf_mel_windows(maf_ctx* ctx, const float* mel_dev, int32_t n_mels, int64_t frames, const int32_t* seeks_dev,
                    const int32_t* lens_dev, int32_t n_windows, int32_t window_frames, int32_t out_fp16, void* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAF_H_ */
